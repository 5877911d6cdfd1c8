"""Ad-hoc: dense generators of moderate size through the DMMA GEMM path (launch per term, CUDA graph once settled) and
through the persistent sweep of the sparse path with full-width rows (W = d)."""
import sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
for d in (48, 100, 200, 400, 512):
    for n_traj in (16, 64):
        w = W.dummy_dense(d=d, n_traj=n_traj, n_controls=2, n_grid=41, seed=3)
        line = f"d={d:4d} N={n_traj:3d}:"
        for fp in (2, 3):
            out = []
            def cb(wrk, it, *args):
                i = wrk.engine.info(); out.append((i["ms_last"], i["launches_last"], i["m_fw"], wrk.result.J_T))
            K.optimize(to_problem(w, iter_stop=5, callback=cb, force_path=fp), method=K.Krotov)
            ms, nl, m, jt = out[-1]
            line += f"  path {fp}: {ms:7.2f} ms/iteration ({ms / (2 * w.N_T) * 1e3:6.1f} us per step-direction, m={m}, {nl} launches, J_T={jt:.10f})"
        print(line, flush=True)
