"""Ad-hoc: sparse path (persistent sweep, launch-per-term stream) vs dense (DMMA) path on a spin chain."""
import os, sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from util import *  # noqa
import argparse
ap = argparse.ArgumentParser(); ap.add_argument("--spins", type=int, default=12); ap.add_argument("--n", type=int, default=64)
ap.add_argument("--grid", type=int, default=21); ap.add_argument("--no-dense", action="store_true"); ap.add_argument("--variants", action="store_true"); ap.add_argument("--sweep-only", action="store_true")
a = ap.parse_args()
t = time.time(); w = W.spin_chain(n_spins=a.spins, n_traj=a.n, n_grid=a.grid); print("workload built", round(time.time() - t, 1), "s", flush=True)
variants = [("sweep", {}), ("sweep cg-sync", {"KROTOV_SWEEP_CGSYNC": "1"}), ("sweep R=4", {"KROTOV_SWEEP_ROWS": "4"}),
            ("sweep R=2", {"KROTOV_SWEEP_ROWS": "2"}), ("sweep R=1", {"KROTOV_SWEEP_ROWS": "1"}),
            ("launch per term", {"KROTOV_NO_SWEEP": "1"})]
if not a.variants:
    variants = [variants[0]] if a.sweep_only else [variants[0], variants[-1]]
runs = [(0, n, e) for n, e in variants] + ([] if a.no_dense else [(2, "dense", {})])
for fp, name, env in runs:
    for k in ("KROTOV_NO_SWEEP", "KROTOV_SWEEP_CGSYNC", "KROTOV_SWEEP_ROWS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    out = []
    def cb(wrk, it, *args):
        i = wrk.engine.info(); out.append((it, wrk.result.J_T, i["ms_last"], i["launches_last"], i["m_fw"], i["path"], i["ell_width"], i["nnz_union"], i["grid_blocks"]))
    t = time.time(); res = K.optimize(to_problem(w, iter_stop=3, callback=cb, force_path=fp), method=K.Krotov)
    it, jt, ms, nl, m, path, W_, nnz, grid = out[-1]
    steps = 2 * w.N_T
    d = w.d
    vec_bytes = 16.0 * d * ((a.n + 7) // 8 * 8)
    # per term: gather reads W rows of the state block, + V_{j-2}, OUT read, V_j, OUT written
    l2_bytes = (W_ + 4) * vec_bytes * (m - 1) * steps if path == 3 else 0
    print(f"{name:16s} path={path} W={W_} nnz={nnz} m={m} grid={grid}: {ms:.2f} ms per iteration ({ms/steps*1e3:.1f} us per step-direction, {nl} launches), "
          f"J_T={jt:.12f}" + (f", L2->SM {l2_bytes/ms/1e6:.0f} GB/s" if path == 3 else f", {(m-1)*steps*8.0*d*d*64/ms/1e9:.1f} TFLOP/s"), flush=True)
