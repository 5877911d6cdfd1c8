"""Ad-hoc: block path (sparse / dense) on launch-bound problems, direct launches vs the whole-iteration CUDA graph."""
import os, sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
cases = [("two transmons, 8 levels (d=64), sparse block path", lambda: W.c3_two_transmon(levels=8), 3),
         ("two transmons, 8 levels (d=64), dense block path", lambda: W.c3_two_transmon(levels=8), 2),
         ("dummy dense d=100, 20 trajectories, N_T=200", lambda: W.dummy_dense(d=100, n_traj=20, n_controls=1, n_grid=201, seed=3), 0)]
for name, make, fp in cases:
    for graph in (0, 1):
        if graph: os.environ.pop("KROTOV_NO_GRAPH", None)
        else: os.environ["KROTOV_NO_GRAPH"] = "1"
        w = make()
        rows = []
        def cb(wrk, it, *a):
            if it >= 1: rows.append((wrk.engine.info()["ms_last"], wrk.engine.info()["launches_last"], time.perf_counter(), wrk.result.J_T))
        kw = dict(force_path=fp) if fp else {}
        K.optimize(to_problem(w, iter_stop=6, callback=cb, lambda_a=1e3 * w.lambda_a, **kw), method=K.Krotov)
        walls = [1e3 * (b[2] - a[2]) for a, b in zip(rows, rows[1:])]
        print(f"{name}: graph={graph} device ms per iteration {[round(r[0], 1) for r in rows]} wall {[round(x, 1) for x in walls]} launches {rows[-1][1]} J_T {rows[-1][3]:.12f}", flush=True)
