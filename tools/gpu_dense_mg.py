"""Ad-hoc: dense (DMMA) path on 1..N GPUs under torchrun: C5 width, short grid."""
import os, sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch, torch.distributed as dist
from util import *  # noqa
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
comm = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    from krotov_jl_b200.distributed import Comm
    comm = Comm(device=lr)
w = W.c5_dense(d=4096, n_traj=64, n_grid=7)
ms = []
def cb(wrk, it, *a):
    if it >= 1: ms.append(wrk.engine.info()["ms_last"])
res = K.optimize(to_problem(w, iter_stop=2, callback=cb, device=lr), method=K.Krotov, comm=comm)
t = torch.tensor([ms[-1]], dtype=torch.float64, device="cuda")
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    gemms = w.N_T * (2 * 23 + 1)
    print(f"world={world}: {float(t[0]):.1f} ms per iteration of {w.N_T} steps, J_T={res.J_T:.12f}, aggregate {gemms*8.0*4096*4096*64/float(t[0])/1e9:.1f} TFLOP/s", flush=True)
if world > 1: dist.destroy_process_group()
