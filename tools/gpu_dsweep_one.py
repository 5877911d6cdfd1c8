"""Ad-hoc: two Krotov iterations of a dense d = 200 generator, 64 trajectories, through the cluster sweep (what the ncu
capture of dense_cluster_sweep_kernel runs)."""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
d = int(sys.argv[1]) if len(sys.argv) > 1 else 200
w = W.dummy_dense(d=d, n_traj=64, n_controls=2, n_grid=21, seed=3)
out = run_product(w, 2)
print("J_T", out["J_T"], "ms", out["info"]["ms_last"], "launches", out["info"]["launches_last"], "grid", out["info"]["grid_blocks"])
