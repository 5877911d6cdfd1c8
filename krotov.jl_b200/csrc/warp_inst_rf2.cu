// krotov_warp_kernel instances with the replicated forward sweep compiled in: one and three controls (register-resident
// rows) and runtime L (rows reloaded per use).
#include "kernel_table.h"
namespace kr {
void add_warp_instances_rf2(KernelMap &t) {
    KR_INSTR(1, 1, 256); KR_INSTR(2, 1, 256); KR_INSTR(3, 1, 256); KR_INSTR(4, 1, 256); KR_INSTR(5, 1, 256);
    KR_INSTR(6, 1, 256); KR_INSTR(7, 1, 256); KR_INSTR(8, 1, 256); KR_INSTR(10, 1, 256); KR_INSTR(12, 1, 256);
    KR_INSTR(2, 3, 256); KR_INSTR(4, 3, 256); KR_INSTR(6, 3, 256);
    KR_INSTR(1, 0, 512); KR_INSTR(2, 0, 512); KR_INSTR(3, 0, 512); KR_INSTR(4, 0, 512); KR_INSTR(5, 0, 512);
    KR_INSTR(6, 0, 512); KR_INSTR(7, 0, 512); KR_INSTR(8, 0, 512); KR_INSTR(10, 0, 512); KR_INSTR(12, 0, 512);
    KR_INSTR(16, 0, 256); KR_INSTR(20, 0, 256); KR_INSTR(24, 0, 256); KR_INSTR(31, 0, 256);
}
}  // namespace kr
