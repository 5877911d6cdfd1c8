"""Ad-hoc: in-kernel cycle counters of one iteration for several ensemble sizes (KROTOV_PROF=1)."""
import os, sys
os.environ["KROTOV_PROF"] = "1"
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
for ns in [1, 8, 64, 256]:
    w = W.c4_ensemble(n_samples=ns)
    out = {}
    def cb(wrk, it, a, b):
        if it >= 1:
            out["prof"] = wrk.engine.profile(); out["info"] = wrk.engine.info()
            n = out["info"]["grid_blocks"]
            out["all"] = [wrk.engine.profile(c) for c in range(n)]
    K.optimize(to_problem(w, iter_stop=3, callback=cb), method=K.Krotov)
    pr, info = out["prof"], out["info"]
    ghz = 1.965
    print(f"samples={ns:4d} grid={info['grid_blocks']}x{info['block_threads']} ms={info['ms_last']:.2f} | " +
          " ".join(f"{k}={v/ghz/1e3/w.N_T:.3f}us/step" for k, v in pr.items()))
    al = out["all"]
    f = lambda v: v / ghz / 1e3 / w.N_T
    for key in ["backward", "forward", "overlap", "wait_pulse", "fw_step_total", "comm_gather"]:
        vals = np.array([f(a[key]) for a in al])
        print(f"      {key:12s} cta0={vals[0]:.3f} others: min={vals[1:].min() if len(vals)>1 else 0:.3f} mean={vals[1:].mean() if len(vals)>1 else 0:.3f} max={vals[1:].max() if len(vals)>1 else 0:.3f}")
