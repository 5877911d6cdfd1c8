"""Ad-hoc: single-GPU throughput of the warp kernel vs ensemble size (N = 4 x samples trajectories)."""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
for ns in [64, 256, 512, 1024, 2048]:
    w = W.c4_ensemble(n_samples=ns, n_grid=501)
    out = {}
    def cb(wrk, it, a, b):
        if it >= 1: out["info"] = wrk.engine.info()
    K.optimize(to_problem(w, iter_stop=3, callback=cb), method=K.Krotov)
    i = out["info"]
    st = 2.0 * w.N * w.N_T / (i["ms_last"] * 1e-3)
    flops = 2.0 * w.N * w.N_T * 16 * 8 * 137 + w.N * w.N_T * 2 * 8 * (137 + 25)
    print(f"N={w.N:5d} grid={i['grid_blocks']}x{i['block_threads']} ms={i['ms_last']:.2f} state-timesteps/s={st/1e6:.1f}M  fp64={flops/i['ms_last']/1e9:.2f} TFLOP/s", flush=True)
