"""Ad-hoc: time per time step of the tiny kernel against the grid length (is a 0.4 ms launch clock-ramp bound?)."""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
for n_grid in [501, 5001, 50001]:
    w = W.c1_tls(n_grid=n_grid)
    w.tlist = np.linspace(0.0, 5.0 * (n_grid - 1) / 500, n_grid)  # same dt as C1, longer pulse
    w.controls = [lambda t: 0.2]
    w.update_shape = lambda t: 1.0
    ms = []
    def cb(wrk, it, *a):
        if it >= 1: ms.append((wrk.engine.info()["ms_last"], wrk.engine.info()["m_fw"]))
    K.optimize(to_problem(w, iter_stop=6, callback=cb), method=K.Krotov)
    best = min(m for m, _ in ms[1:])
    print(f"N_T={n_grid-1}: {best:.3f} ms per iteration, {1e3*best/(n_grid-1):.3f} us per time step (both sweeps), m={ms[-1][1]}", flush=True)
