"""Ad-hoc: warp-kernel variants for ensembles beyond one trajectory per warp (C4's generator, 501-point grid)."""
import os, sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
cases = [(512, "pair kernel", {}), (512, "sequential, rows in registers", {"KROTOV_SEQ_PREG": "1"}),
         (1024, "sequential, rows in registers", {}), (1024, "rows re-read (previous)", {"KROTOV_NO_SEQ_PREG": "1"}),
         (768, "sequential, rows in registers", {}), (768, "rows re-read (previous)", {"KROTOV_NO_SEQ_PREG": "1"})]
ws = {}
for ns, name, env in cases:
    for k in ("KROTOV_SEQ_PREG", "KROTOV_NO_SEQ_PREG"):
        os.environ.pop(k, None)
    os.environ.update(env)
    if ns not in ws:
        ws[ns] = W.c4_ensemble(n_samples=ns, n_grid=501)
    w = ws[ns]
    out = {}
    def cb(wrk, it, a, b):
        if it >= 1: out["info"] = wrk.engine.info(); out["J_T"] = wrk.result.J_T; out["p"] = float(np.abs(np.array(a)).max())
    K.optimize(to_problem(w, iter_stop=3, callback=cb), method=K.Krotov)
    i = out["info"]
    st = 2.0 * w.N * w.N_T / (i["ms_last"] * 1e-3)
    print(f"N={w.N:5d} {name:32s} grid={i['grid_blocks']}x{i['block_threads']} ms={i['ms_last']:.2f} state-timesteps/s={st/1e6:.1f}M  J_T={out['J_T']:.14f} max|eps|={out['p']:.14f}", flush=True)
