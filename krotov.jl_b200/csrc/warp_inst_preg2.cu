// krotov_warp_kernel instances: register-resident term rows, L = 2 (C3 / C4 run on <6, 2>).
#include "kernel_table.h"
namespace kr {
void add_warp_instances_preg2(KernelMap &t) {
    KR_INST(1, 2, 256); KR_INST(2, 2, 256); KR_INST(3, 2, 256); KR_INST(4, 2, 256); KR_INST(5, 2, 256);
    KR_INST(6, 2, 256); KR_INST(7, 2, 256); KR_INST(8, 2, 256); KR_INST(10, 2, 256);
}
}  // namespace kr
