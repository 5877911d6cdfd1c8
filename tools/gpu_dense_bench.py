"""Ad-hoc: throughput of the dense (DMMA) path on a C5-shaped problem with a short time grid."""
import sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
import argparse
ap = argparse.ArgumentParser(); ap.add_argument("--d", type=int, default=4096); ap.add_argument("--n", type=int, default=64)
ap.add_argument("--grid", type=int, default=11); ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
t = time.time(); w = W.c5_dense(d=a.d, n_traj=a.n, n_grid=a.grid); print("workload built", time.time() - t, flush=True)
out = []
def cb(wrk, it, *args):
    info = wrk.engine.info(); out.append((it, wrk.result.J_T, info["ms_last"], info["launches_last"], info["m_fw"], info["m_bw"]))
    print(out[-1], flush=True)
t = time.time(); res = K.optimize(to_problem(w, iter_stop=a.iters, callback=cb), method=K.Krotov); print("optimize wall", time.time() - t)
N_T, L = w.N_T, w.L
dp = (a.d + 31) // 32 * 32
for it, jt, ms, nl, mf, mb in out[1:]:
    gemms = N_T * (mb - 1) + N_T * (mf - 1) + N_T * L
    flops = gemms * 8.0 * dp * dp * ((a.n + 7) // 8 * 8)
    print(f"iter {it}: {ms:.1f} ms, {nl} launches, {gemms} GEMMs, {flops / ms / 1e9:.2f} TFLOP/s (complex GEMM flops), {ms / (2 * N_T):.3f} ms per step-direction")
print("norms", np.abs(np.linalg.norm(np.array(res.states), axis=1) - 1).max())
