// krotov_warp2_kernel instances: two trajectories of one generator per warp, register-resident rows only.
#include "kernel_table.h"
#include "warp2_kernel.cuh"
namespace kr {
void add_warp2_instances(KernelMap &t) {
    KR_INST2(2, 1); KR_INST2(3, 1); KR_INST2(4, 1); KR_INST2(5, 1); KR_INST2(6, 1); KR_INST2(7, 1); KR_INST2(8, 1);
    KR_INST2(2, 2); KR_INST2(3, 2); KR_INST2(4, 2); KR_INST2(5, 2); KR_INST2(6, 2); KR_INST2(7, 2); KR_INST2(8, 2);
    KR_INST2(4, 3); KR_INST2(6, 3);
}
}  // namespace kr
