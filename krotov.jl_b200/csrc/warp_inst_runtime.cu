// krotov_warp_kernel instances: runtime L (rows reloaded from L1/L2 per use), one warp per trajectory.
#include "kernel_table.h"
namespace kr {
void add_warp_instances_runtime(KernelMap &t) {
    KR_INST(1, 0, 512); KR_INST(2, 0, 512); KR_INST(3, 0, 512); KR_INST(4, 0, 512); KR_INST(5, 0, 512);
    KR_INST(6, 0, 512); KR_INST(7, 0, 512); KR_INST(8, 0, 512); KR_INST(10, 0, 512); KR_INST(12, 0, 512);
    // wide rows: 255 registers per thread (at most 7 trajectory warps per CTA) -- at 512 threads per CTA these
    // instances spilled 316 ... 773 bytes per thread into local memory
    KR_INST(16, 0, 256); KR_INST(20, 0, 256); KR_INST(24, 0, 256); KR_INST(31, 0, 256);
}
}  // namespace kr
