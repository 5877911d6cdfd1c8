"""Ad-hoc: cost of one spectral-range update (host) on C4."""
import sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
w = W.c4_ensemble()
def cb(wrk, it, *a):
    if it == 2:
        for name, st, d in (("fw", wrk.fw_settings, 0), ("bw", wrk.bw_settings, 1)):
            for rep in range(2):
                t0 = time.perf_counter(); st._derive(); t1 = time.perf_counter(); st.push(wrk.engine, d); t2 = time.perf_counter()
                print(f"{name}: derive {1e3*(t1-t0):.1f} ms (tabulate incl.), push {1e3*(t2-t1):.1f} ms")
        import cProfile, pstats
        pr = cProfile.Profile(); pr.enable(); wrk.fw_settings._derive(); wrk.fw_settings.push(wrk.engine, 0); pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(8)
K.optimize(to_problem(w, iter_stop=2, callback=cb), method=K.Krotov)
