"""Cost of the second-order update (sigma) on C4 at full size: device ms and host wall ms per iteration, first order vs
second order with a per-iteration re-estimated sigma (NumericalSigma).   python tools/gpu_second_order.py [samples]"""
import sys
import time

sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np  # noqa: E402
from util import K, W, run_product  # noqa: E402

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 256
w = W.c4_ensemble(n_samples=ns)
warm, steps = 3, 8


def run(label, **kw):
    wall, dev = [], []

    def cb(wrk, it, a, b):
        wall.append(time.perf_counter())
        if it >= 1:
            dev.append(wrk.engine.info()["ms_last"])

    h = run_product(w, warm + steps, callback=cb, **kw)
    wall_ms = 1e3 * np.diff(wall)[warm:]
    print(f"{label}: device {np.mean(dev[warm:]):.2f} ms/iteration, wall {np.mean(wall_ms):.2f} ms/iteration "
          f"(min {wall_ms.min():.2f}, max {wall_ms.max():.2f}); J_T {h['J_T'][0]:.6f} -> {h['J_T'][-1]:.6f}")
    return h


print(f"C4: {w.N} trajectories, d = {w.d}, N_T = {w.N_T}")
a = run("first order ")
sig = K.NumericalSigma(0.0, 1e-4)
b = run("second order", sigma=sig)
print(f"sigma after the run: {sig(0.0):.3e} (A = {sig.A:.3e}); max |pulse difference| to first order: "
      f"{np.abs(a['pulses'] - b['pulses']).max():.3e}")
