// krotov_warp_kernel instances that serve SEVERAL emulated ranks in one cooperative launch (krotov_group_iterate:
// the multi-rank exchange protocols on a single device, for tests and diagnostics).
#include "kernel_table.h"
namespace kr {
void add_warp_instances_emul(KernelMap &t) {
    KR_INSTE(6, 2, 256);  // two coupled transmons (C3 / C4), register-resident rows
    KR_INSTE(6, 0, 512);  // the same pattern with rows reloaded per use
    KR_INSTE(2, 1, 256);  // single transmon / spin-1 chains, one control
}
}  // namespace kr
