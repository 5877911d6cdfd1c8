// krotov_warp_kernel instances: 64 threads (rows) per trajectory for 32 < d <= 64 with narrow rows.
#include "kernel_table.h"
namespace kr {
void add_warp_instances_wide64(KernelMap &t) {
    KR_INSTW(4, 0, 512, 64); KR_INSTW(6, 0, 512, 64); KR_INSTW(8, 0, 512, 64); KR_INSTW(10, 0, 512, 64);
    KR_INSTW(12, 0, 512, 64); KR_INSTW(16, 0, 512, 64); KR_INSTW(20, 0, 512, 64); KR_INSTW(24, 0, 512, 64);
    KR_INSTW(4, 1, 256, 64); KR_INSTW(6, 1, 256, 64); KR_INSTW(8, 1, 256, 64); KR_INSTW(10, 1, 256, 64);
    KR_INSTW(4, 2, 256, 64); KR_INSTW(6, 2, 256, 64); KR_INSTW(8, 2, 256, 64);
}
}  // namespace kr
