"""Ad-hoc: warps per CTA against ensemble size -- is it better to spread few trajectories over many SMs
(one warp per SM: no crossbar sharing, but a grid sum per time step) or to pack them (KROTOV_WPC)?"""
import os, sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
for ns in [1, 2, 8, 16, 37, 64, 128]:
    w = W.c4_ensemble(n_samples=ns)
    row = []
    for wpc in [0, 1, 2, 3, 4, 7]:
        if wpc: os.environ["KROTOV_WPC"] = str(wpc)
        else: os.environ.pop("KROTOV_WPC", None)
        ms = []
        def cb(wrk, it, *a):
            if it >= 1: ms.append((wrk.engine.info()["ms_last"], wrk.engine.info()["grid_blocks"]))
        try:
            K.optimize(to_problem(w, iter_stop=4, callback=cb), method=K.Krotov)
            row.append(f"wpc={wpc or 'auto'}: {min(m for m, _ in ms[1:]):6.2f} ms ({ms[-1][1]} CTAs)")
        except Exception as e:
            row.append(f"wpc={wpc}: {type(e).__name__}")
    print(f"samples={ns:4d} N={w.N:4d} | " + " | ".join(row), flush=True)
