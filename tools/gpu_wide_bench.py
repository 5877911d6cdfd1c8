"""Ad-hoc: persistent kernel with wide groups (d in (32,128]) vs the SpMM block path on two transmons with more levels."""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
for levels in (5, 6, 8, 11):
    w = W.c3_two_transmon(levels=levels)
    for fp in (0, 3):
        if levels == 5 and fp == 3:
            continue
        out = {}
        def cb(wrk, it, *a):
            if it >= 1: out["i"] = wrk.engine.info()
        K.optimize(to_problem(w, iter_stop=2, callback=cb, force_path=fp), method=K.Krotov)
        i = out["i"]
        print(f"two transmons, {levels} levels (d={w.d}), 4 trajectories, N_T=2000: path={i['path']} {i['ms_last']:.1f} ms per iteration, {i['launches_last']} launches, grid {i['grid_blocks']}x{i['block_threads']}", flush=True)
