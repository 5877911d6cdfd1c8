"""Two ranks on two GPUs: trajectories sharded, per-step overlap sums exchanged in-kernel over NVLink mailboxes.
Needs >= 2 CUDA devices (skipped on the single-GPU box); run with `gpurun --gpus 2`."""
import os

import numpy as np
import pytest

import workloads as W

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n_samples, n_grid, iters, q):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import krotov_jl_b200 as K
        from krotov_jl_b200.distributed import Comm
        from util import to_problem

        comm = Comm(device=rank)
        w = W.c4_ensemble(n_samples=n_samples, n_grid=n_grid)
        hist = {"J_T": [], "shard": None}

        def cb(wrk, it, eps_new, eps_old):
            hist["J_T"].append(wrk.result.J_T)
            hist["pulses"] = np.array([np.array(e) for e in eps_new])
            hist["shard"] = wrk._shard
            hist["ga"] = np.array(wrk.g_a_int)

        res = K.optimize(to_problem(w, iter_stop=iters, callback=cb, device=rank), method=K.Krotov, comm=comm)
        q.put((rank, hist["J_T"], hist["pulses"], hist["shard"], res.message, np.array(res.states), hist["ga"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_samples,n_grid", [(8, 201), (64, 101)])
def test_two_ranks_match_single_gpu(n_samples, n_grid):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from util import run_product

    iters = 2
    single = run_product(W.c4_ensemble(n_samples=n_samples, n_grid=n_grid), iters)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_samples, n_grid, iters, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, J0, P0, s0, m0, st0, ga0), (r1, J1, P1, s1, m1, st1, ga1) = out
    assert m0 == m1 == "Reached maximum number of iterations"
    assert s0 == (0, 2 * n_samples) and s1 == (2 * n_samples, 4 * n_samples)
    # replicas must hold bit-identical pulses (every rank applies the same rank-ordered sum)
    assert np.array_equal(P0, P1) and J0 == J1 and np.array_equal(ga0, ga1)
    # and agree with the single-GPU run up to the summation order of the overlap sums
    assert np.abs(np.array(J0) - np.array(single["J_T"])).max() < 1e-12
    assert np.abs(P0 - single["pulses"]).max() < 1e-12
    assert np.abs(st0 - np.array(single["result"].states)).max() < 1e-11
