"""Ad-hoc first GPU check (not a pytest file): parity of the CUDA path against the oracle on C1-C3 and a cut-down C4."""
import sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
from oracle import krotov_oracle as O, c_oracle as C

def compare(w, iters, use_c=False):
    p = W.to_oracle(w)
    t = time.time(); ref = (C.optimize_krotov_c(p, iters) if use_c else O.optimize_krotov(p, iters)); t_or = time.time() - t
    t = time.time(); got = run_product(w, iters); t_gpu = time.time() - t
    r, a = rel_abs(got["J_T"], ref["J_T"])
    print(f"== {w.name}: oracle {t_or:.2f}s product {t_gpu:.2f}s info={got['info']}")
    print("   J_T oracle ", [f"{x:.12e}" for x in ref["J_T"]])
    print("   J_T product", [f"{x:.12e}" for x in got["J_T"]])
    print("   rel", [f"{x:.1e}" for x in r], "abs", [f"{x:.1e}" for x in a])
    print("   pulses max-abs diff", np.abs(got["pulses"] - ref["pulses"]).max(), " g_a diff", np.abs(np.array(got["g_a_int"]) - np.array(ref["g_a_int"])).max())
    sys.stdout.flush()

compare(W.c1_tls(), 5)
compare(W.c2_transmon_x(), 4)
compare(W.dummy_dense(d=10, n_traj=2, n_controls=2), 3)
compare(W.dummy_dense(d=32, n_traj=5, n_controls=3, hermitian=False), 2)
compare(W.c3_two_transmon(), 2, use_c=True)
compare(W.c4_ensemble(n_samples=8), 2, use_c=True)
compare(W.c4_ensemble(n_samples=64), 2, use_c=True)
