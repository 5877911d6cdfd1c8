import os, sys, time
os.environ["KROTOV_TRACE"] = "1"
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
w = W.c4_ensemble()
def cb(wrk, it, *a):
    if it == 1:
        for _ in range(2):
            t = time.perf_counter(); wrk.fw_settings.push(wrk.engine, 0); print("push ms", 1e3 * (time.perf_counter() - t))
            t = time.perf_counter(); wrk.fw_settings._derive(); print("derive ms", 1e3 * (time.perf_counter() - t))
K.optimize(to_problem(w, iter_stop=1, callback=cb), method=K.Krotov)
