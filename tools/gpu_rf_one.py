"""Ad-hoc: two Krotov iterations with 2 EMULATED ranks in replicate mode (what the ncu capture of the RF backward kernel runs).
    KROTOV_WPC=7 python tools/gpu_rf_one.py [samples] [n_grid]"""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ng = int(sys.argv[2]) if len(sys.argv) > 2 else 501
w = W.c4_ensemble(n_samples=ns, n_grid=ng)
out = run_product(w, 2, emulate_ranks=2, multi_gpu="replicate")
print("J_T", out["J_T"], "ms", out["info"]["ms_last"], "exchange", out["info"]["exchange"], "grid", out["info"]["grid_blocks"])
