"""Ad-hoc: per-iteration host phases of optimize() on C4 (which phase makes an 'event' iteration slow?).

    python tools/gpu_e2e_phases.py [iterations]      # KROTOV_TRACE=1 adds the library's own laps on stderr
"""
import sys
import time

sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
from krotov_jl_b200 import cheby as CH
from krotov_jl_b200 import engine as EN

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 14
acc = {}


def timed(cls, name):
    orig = getattr(cls, name)

    def wrapper(*a, **k):
        t = time.perf_counter()
        try:
            return orig(*a, **k)
        finally:
            acc[name] = acc.get(name, 0.0) + 1e3 * (time.perf_counter() - t)

    setattr(cls, name, wrapper)


for n in ("_derive", "_tabulate", "push", "reinit"):
    timed(CH.ChebyDirection, n)
for n in ("iterate", "tau", "set_cheby"):
    timed(EN.KrotovCuda, n)

w = W.c4_ensemble()
last = [time.perf_counter()]


def cb(wrk, it, *a):
    now = time.perf_counter()
    dev = wrk.engine.info()["ms_last"]
    print(f"iter {it}: wall {1e3 * (now - last[0]):7.2f} ms  device {dev:6.2f}  m_fw {int(wrk.fw_settings.coeff_count.max())} "
          f"m_bw {int(wrk.bw_settings.coeff_count.max())}  " + "  ".join(f"{k} {v:.2f}" for k, v in sorted(acc.items())), flush=True)
    acc.clear()
    last[0] = time.perf_counter()


K.optimize(to_problem(w, iter_stop=iters, callback=cb), method=K.Krotov)
