"""Ad-hoc: in-kernel cycle counters of the multi-rank exchange variants (run under torchrun, KROTOV_PROF=1).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 tools/gpu_mg_prof.py [samples]
"""
import os, sys
os.environ["KROTOV_PROF"] = "1"
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch
import torch.distributed as dist
from util import *  # noqa
from krotov_jl_b200.distributed import Comm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
comm = Comm(device=local)
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 256
w = W.c4_ensemble(n_samples=ns)
ghz = 1.965
mode = os.environ.get("MG_MODE", "shard")  # shard (per-step exchange) | replicate (replicated forward sweep)
variants = [("hierarchical", {"KROTOV_XCHG": "hier"}), ("hierarchical, stores", {"KROTOV_XCHG": "hierst"}),
            ("one-hop", {"KROTOV_XCHG": "onehop"}), ("mailbox", {"KROTOV_XCHG": "mbox"})]
if mode == "replicate":
    variants = [("replicated forward sweep", {})]
if os.environ.get("MG_VARIANTS") and mode != "replicate":
    variants = [v for v in variants if v[1]["KROTOV_XCHG"] in os.environ["MG_VARIANTS"].split(",")]
if os.environ.get("MG_WPC"):
    variants = [(n + " wpc=" + w, dict(e, KROTOV_WPC=w)) for w in os.environ["MG_WPC"].split(",") for n, e in variants]
for name, env in variants:
    for k in ("KROTOV_XACC_STRIDE", "KROTOV_NO_XACC", "KROTOV_XCHG", "KROTOV_WPC"):
        os.environ.pop(k, None)
    os.environ.update(env)
    out = {"ms": []}

    def cb(wrk, it, a, b):
        if it >= 1:
            info = wrk.engine.info()
            out["ms"].append(info["ms_last"])
            out["info"] = info
            out["all"] = [wrk.engine.profile(c) for c in range(info["grid_blocks"])]

    K.optimize(to_problem(w, iter_stop=4, callback=cb, device=local, multi_gpu=mode), method=K.Krotov, comm=comm)
    comm.barrier()
    f = lambda v: v / ghz / 1e3 / w.N_T
    if rank != 0 and not os.environ.get("MG_ALL_RANKS"):
        for r in range(world):
            comm.barrier()
        continue
    lines = [f"[rank {rank}] {name}: exchange={out['info'].get('exchange')} grid={out['info']['grid_blocks']}x{out['info']['block_threads']} ms={['%.2f' % m for m in out['ms']]} fallback={out['info'].get('fallback_steps')} rank_barrier_wait_ms={out['info'].get('ms_rank_wait', 0.0):.3f}"]
    for key in ["backward", "forward", "overlap", "wait_pulse", "fw_step_total", "comm_wait_partials", "comm_reduce", "comm_gather"]:
        vals = np.array([f(a[key]) for a in out["all"]])
        lines.append(f"      {key:18s} cta0={vals[0]:.3f} all: min={vals.min():.3f} mean={vals.mean():.3f} max={vals.max():.3f} us/step")
    for r in range(world):
        if r == rank:
            print("\n".join(lines), flush=True)
        comm.barrier()
dist.destroy_process_group()
