// krotov_warp_kernel instances: 128 threads (rows) per trajectory for 64 < d <= 128 with narrow rows.
#include "kernel_table.h"
namespace kr {
void add_warp_instances_wide128(KernelMap &t) {
    KR_INSTW(4, 0, 512, 128); KR_INSTW(6, 0, 512, 128); KR_INSTW(8, 0, 512, 128); KR_INSTW(12, 0, 512, 128);
    KR_INSTW(16, 0, 512, 128); KR_INSTW(24, 0, 512, 128);
    KR_INSTW(4, 1, 256, 128); KR_INSTW(6, 1, 256, 128); KR_INSTW(8, 1, 256, 128); KR_INSTW(4, 2, 256, 128);
    KR_INSTW(6, 2, 256, 128); KR_INSTW(8, 2, 256, 128);
}
}  // namespace kr
