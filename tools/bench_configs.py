"""Throughput of all five BASELINE configs on one B200, with the CPU restatement timed beside each.

    python tools/bench_configs.py > profiles/r1_configs.json

C1-C4 run at full size.  C5 (d=4096, 64 trajectories, 10000 steps: ~150 s per iteration on the GPU, hours on
the CPU) runs at full WIDTH on a short grid and is scaled linearly in N_T -- every time step costs the same --
and is marked "extrapolated".  The CPU column is oracle/krotov_oracle.c with OpenMP (Julia is not installed);
for C4/C5 it runs on a bounded sample and is scaled per state-timestep."""
import json, os, sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
from util import *  # noqa
from oracle import c_oracle as C

def gpu(w, iters, warm=2, **kw):
    ms, info = [], {}
    def cb(wrk, it, *a):
        if it >= 1:
            i = wrk.engine.info(); ms.append(i["ms_last"]); info.update(i)
    t = time.perf_counter()
    res = K.optimize(to_problem(w, iter_stop=warm + iters, callback=cb, **kw), method=K.Krotov)
    wall = time.perf_counter() - t
    m = float(np.mean(ms[warm:]))
    return m, info, res

out = {}
for name, make, iters, cpu_make, cpu_iters in [
    ("C1 TLS d=2 N=1 N_T=500", W.c1_tls, 10, W.c1_tls, 5),
    ("C2 transmon X d=3 N=2 N_T=500", W.c2_transmon_x, 10, W.c2_transmon_x, 5),
    ("C3 two-transmon d=25 N=4 N_T=2000", W.c3_two_transmon, 6, W.c3_two_transmon, 3),
    ("C4 ensemble d=25 N=1024 N_T=2000", W.c4_ensemble, 6, lambda: W.c4_ensemble(n_samples=64), 2),
]:
    w = make()
    ms, info, res = gpu(w, iters)
    wc = cpu_make(); c = C.optimize_krotov_c(W.to_oracle(wc), cpu_iters)
    cpu_st = 2.0 * wc.N * wc.N_T * cpu_iters / c["secs"]
    st = 2.0 * w.N * w.N_T / (ms * 1e-3)
    out[name] = dict(gpu_ms_per_iteration=ms, gpu_iterations_per_s=1e3 / ms, gpu_state_timesteps_per_s=st,
                     us_per_time_step_both_sweeps=1e3 * ms / w.N_T, grid=[info["grid_blocks"], info["block_threads"]],
                     m=info["m_fw"], J_T=res.J_T, cpu_state_timesteps_per_s=cpu_st, cpu_threads=c["threads"],
                     cpu_sample=f"{wc.N} trajectories, {cpu_iters} iterations, {c['secs']:.2f} s",
                     speedup_per_state_timestep=st / cpu_st)
    print(name, json.dumps(out[name]), file=sys.stderr, flush=True)

# C5: full width, short grid
n_grid = 11
w = W.c5_dense(d=4096, n_traj=64, n_grid=n_grid)
ms, info, res = gpu(w, 2, warm=1)
N_T = w.N_T
m = info["m_fw"]
gemms = N_T * (2 * (m - 1) + w.L)
flops = gemms * 8.0 * 4096 * 4096 * 64
wc = W.c5_dense(d=512, n_traj=16, n_grid=5)
c = C.optimize_krotov_c(W.to_oracle(wc), 1)
cpu_flops_per_s = (wc.N_T * (2 * (c["m"][0] - 1) * 2 + wc.L)) * 8.0 * 512 * 512 * 16 / c["secs"]  # 2 terms (H0, H1) per matvec
out["C5 dense d=4096 N=64 N_T=10000 (extrapolated from N_T=%d)" % N_T] = dict(
    gpu_ms_per_time_step_both_sweeps=ms / N_T, gpu_s_per_iteration_extrapolated=ms / N_T * 10000 * 1e-3,
    gpu_state_timesteps_per_s=2.0 * 64 * N_T / (ms * 1e-3), gemm_tflops=flops / (ms * 1e-3) / 1e12,
    dmma_peak_tflops=37.1, frac_of_dmma_peak=flops / (ms * 1e-3) / 1e12 / 37.1, m=m,
    cpu_gflops_sample=cpu_flops_per_s / 1e9, cpu_threads=c["threads"],
    cpu_sample=f"d=512, 16 trajectories, {wc.N_T} steps, {c['secs']:.2f} s (naive CSR loops; not extrapolated to d=4096)")
print(json.dumps(out, indent=1))
