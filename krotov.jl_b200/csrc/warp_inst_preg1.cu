// krotov_warp_kernel instances: register-resident term rows ((1+L)(W+1) <= 36 double2), L = 1 and L = 3.
#include "kernel_table.h"
namespace kr {
void add_warp_instances_preg1(KernelMap &t) {
    KR_INST(1, 1, 256); KR_INST(2, 1, 256); KR_INST(3, 1, 256); KR_INST(4, 1, 256); KR_INST(5, 1, 256);
    KR_INST(6, 1, 256); KR_INST(7, 1, 256); KR_INST(8, 1, 256); KR_INST(10, 1, 256); KR_INST(12, 1, 256);
    KR_INST(2, 3, 256); KR_INST(4, 3, 256); KR_INST(6, 3, 256);
}
}  // namespace kr
