"""Ad-hoc: two Krotov iterations of a spin chain on the sparse path (what the ncu captures of the sweep kernel run)."""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
spins = int(sys.argv[1]) if len(sys.argv) > 1 else 10
w = W.spin_chain(n_spins=spins, n_traj=64, n_grid=21)
out = run_product(w, 2)
print("J_T", out["J_T"], "ms", out["info"]["ms_last"], "launches", out["info"]["launches_last"], "grid", out["info"]["grid_blocks"])
