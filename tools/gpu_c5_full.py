"""BASELINE config C5 at FULL size on one B200: d = 4096, 64 trajectories, N_T = 10000, dense GUE-like generator
(FP64 DMMA path, chi trajectory 42 GB in HBM).  One initial forward sweep + `iters` Krotov iterations."""
import json, sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n_grid = int(sys.argv[2]) if len(sys.argv) > 2 else 10001
n_traj = int(sys.argv[3]) if len(sys.argv) > 3 else 64  # fewer columns = the per-GPU shard of a multi-GPU run
t0 = time.time()
w = W.c5_dense(d=4096, n_traj=n_traj, n_grid=n_grid)
print("workload built %.1f s" % (time.time() - t0), flush=True)
rows = []
def cb(wrk, it, *a):
    i = wrk.engine.info()
    rows.append(dict(iteration=it, J_T=wrk.result.J_T, device_ms=i["ms_last"], launches=i["launches_last"], m=i["m_fw"],
                     hbm_bytes_state=i["hbm_bytes_state"], wall_s=time.time() - t0))
    print(json.dumps(rows[-1]), flush=True)
res = K.optimize(to_problem(w, iter_stop=iters, callback=cb), method=K.Krotov)
print("message:", res.message)
it = [r for r in rows if r["iteration"] >= 1]
if it:
    ms = it[-1]["device_ms"]; m = it[-1]["m"]; N_T = w.N_T
    gemms = N_T * (2 * (m - 1) + w.L)
    print(json.dumps(dict(config="C5 full size", N_T=N_T, s_per_iteration=ms * 1e-3, state_timesteps_per_s=2.0 * n_traj * N_T / (ms * 1e-3), n_traj=n_traj,
                          gemm_tflops=gemms * 8.0 * 4096 * 4096 * n_traj / (ms * 1e-3) / 1e12, generator_GBps=gemms * 16.0 * 4096 * 4096 / (ms * 1e-3) / 1e9, monotonic=all(b["J_T"] <= a["J_T"] + 1e-12 for a, b in zip(rows, rows[1:])))))
