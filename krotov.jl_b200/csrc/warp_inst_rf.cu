// krotov_warp_kernel instances with the replicated forward sweep of several ranks compiled in (template parameter RF):
// one warp per trajectory, register-resident term rows for one to three controls and the row-reloading variant.
#include "kernel_table.h"
namespace kr {
void add_warp_instances_rf(KernelMap &t) {
    KR_INSTR(1, 2, 256); KR_INSTR(2, 2, 256); KR_INSTR(3, 2, 256); KR_INSTR(4, 2, 256); KR_INSTR(5, 2, 256);
    KR_INSTR(6, 2, 256); KR_INSTR(7, 2, 256); KR_INSTR(8, 2, 256); KR_INSTR(10, 2, 256);
}
void add_warp_instances_rf_emul(KernelMap &t) {
    KR_INSTRE(6, 2, 256);  // two coupled transmons (C3 / C4), register-resident rows
    KR_INSTRE(6, 0, 512);  // the same pattern with rows reloaded per use
    KR_INSTRE(2, 1, 256);  // single transmon / spin-1 chains, one control
}
}  // namespace kr
