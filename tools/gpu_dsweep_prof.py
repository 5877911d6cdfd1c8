"""Ad-hoc: in-kernel cycle counters of the dense cluster sweep (KROTOV_PROF=1), d = 100 / 200, 64 trajectories."""
import os, sys
os.environ["KROTOV_PROF"] = "1"
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
for d in (100, 200):
    w = W.dummy_dense(d=d, n_traj=64, n_controls=2, n_grid=41, seed=3)
    out = []
    def cb(wrk, it, *args):
        i = wrk.engine.info(); out.append((i["ms_last"], i["launches_last"], i["m_fw"], i["grid_blocks"]))
    K.optimize(to_problem(w, iter_stop=3, callback=cb), method=K.Krotov)
    ms, nl, m, grid = out[-1]
    print(f"d={d} N=64: {ms:.2f} ms/it ({ms / (2 * w.N_T) * 1e3:.1f} us/step-dir, m={m}, {nl} launches, grid {grid})", flush=True)
