"""Ad-hoc: dense generators of moderate size -- the launch-per-term DMMA stream (KROTOV_NO_DSWEEP=1), the persistent
cluster sweep (dense_sweep.cuh, default for d <= 288) and the sparse sweep with full-width rows (force_path=3)."""
import os, sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
for d in (48, 100, 200, 288, 400):
    for n_traj in (16, 64):
        w = W.dummy_dense(d=d, n_traj=n_traj, n_controls=2, n_grid=41, seed=3)
        line = f"d={d:4d} N={n_traj:3d}:"
        for name, fp, env in (("stream", 2, {"KROTOV_NO_DSWEEP": "1"}), ("cluster sweep", 2, {}), ("ELL sweep", 3, {})):
            os.environ.pop("KROTOV_NO_DSWEEP", None)
            os.environ.update(env)
            out = []
            def cb(wrk, it, *args):
                i = wrk.engine.info(); out.append((i["ms_last"], i["launches_last"], i["m_fw"], wrk.result.J_T, i["grid_blocks"]))
            K.optimize(to_problem(w, iter_stop=5, callback=cb, force_path=fp), method=K.Krotov)
            ms, nl, m, jt, grid = out[-1]
            line += f"  {name}: {ms:7.2f} ms/it ({ms / (2 * w.N_T) * 1e3:6.1f} us/step-dir, m={m}, {nl} launches, grid {grid}, J_T={jt:.10f})"
        print(line, flush=True)
