"""Ad-hoc: cProfile of the host side of optimize() on C4 (where do the e2e milliseconds go?)."""
import cProfile, pstats, sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
w = W.c4_ensemble()
t = time.time(); problem = to_problem(w, iter_stop=8); print("build problem", time.time() - t)
pr = cProfile.Profile()
marks = []
def cb(wrk, it, *a):
    marks.append((time.perf_counter(), wrk.engine.info()["ms_last"]))
problem.kwargs["callback"] = cb
pr.enable(); res = K.optimize(problem, method=K.Krotov); pr.disable()
for i in range(1, len(marks)):
    print(f"iter {i}: wall {1e3*(marks[i][0]-marks[i-1][0]):.1f} ms, device {marks[i][1]:.1f} ms")
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
