"""Multi-GPU plumbing: one process per GPU, trajectories sharded across ranks.

The data path has exactly one cross-GPU dependency per time step -- the L-vector of overlap sums
(``src/optimize.jl:340-349``) -- and it is exchanged INSIDE the persistent kernel through peer-mapped
mailboxes over NVLink (``krotov_comm_export`` / ``krotov_comm_connect``).  ``torch.distributed`` is
used only for plumbing: gathering the IPC descriptors once, and gathering tau / states per iteration
(N complex numbers).  With the ``gloo`` backend the same host logic runs on CPU in the tests."""
from __future__ import annotations

import numpy as np

__all__ = ["shard_bounds", "Comm"]


def shard_bounds(gen_of_traj, rank, world):
    """Contiguous block [lo, hi) of trajectories for ``rank``.  Blocks are cut at generator boundaries
    when the trajectories are grouped by generator (ensemble samples stay on one GPU, so each
    generator is stored once); otherwise at equal counts."""
    gen_of_traj = np.asarray(gen_of_traj)
    N = len(gen_of_traj)
    if world <= 1:
        return 0, N
    # boundaries where the generator index changes
    cuts = [0] + [k for k in range(1, N) if gen_of_traj[k] != gen_of_traj[k - 1]] + [N]
    grouped = len(set(gen_of_traj.tolist())) == len(cuts) - 1  # each generator is one contiguous run
    if grouped and len(cuts) - 1 >= world:
        # choose the cut closest to the ideal equal split
        ideal = [round(r * N / world) for r in range(world + 1)]
        bounds = [min(cuts, key=lambda c: abs(c - x)) for x in ideal]
        bounds[0], bounds[-1] = 0, N
        for r in range(1, world + 1):  # keep strictly increasing
            if bounds[r] <= bounds[r - 1]:
                bigger = [c for c in cuts if c > bounds[r - 1]]
                bounds[r] = bigger[0] if bigger else N
        bounds[-1] = N
        if any(bounds[r + 1] <= bounds[r] for r in range(world)):  # degenerate cut-aligned split: equal counts
            bounds = [(r * N) // world for r in range(world + 1)]
    else:
        bounds = [(r * N) // world for r in range(world + 1)]
    if any(bounds[r + 1] <= bounds[r] for r in range(world)):
        # every rank evaluates the same bounds, so every rank raises: nobody is left waiting in a collective
        raise ValueError(f"cannot shard {N} trajectories over {world} ranks: some rank would hold none")
    return int(bounds[rank]), int(bounds[rank + 1])


class Comm:
    """Thin wrapper over an initialised ``torch.distributed`` process group."""

    def __init__(self, device=None):
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.backend = dist.get_backend()
        self.device = self.rank if device is None else device
        # host-side object collectives (IPC descriptors, tau, states) go through a gloo group: they are small
        # CPU buffers, and NCCL would stage them through device memory
        self.cpu_group = dist.new_group(backend="gloo") if self.backend != "gloo" else None

    def all_gather_object(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.cpu_group)
        return out

    def connect(self, engine):
        """Exchange the mailbox IPC descriptors and map the peers' mailboxes.  Ranks must sit on distinct GPUs:
        their persistent kernels wait for one another and two of them on one device may never be co-resident."""
        import socket

        where = self.all_gather_object((socket.gethostname(), int(self.device)))
        if self.backend != "gloo" and len(set(where)) != len(where):
            raise RuntimeError(f"ranks share a GPU: (host, device) per rank = {where}")
        descs = self.all_gather_object(engine.comm_export())
        engine.comm_connect(self.rank, self.world, descs)
        self.barrier()

    def all_gather_rows(self, local, n_total):
        """Concatenate per-rank blocks of rows (tau or states) in rank order.  COLLECTIVE: every rank must call it
        (touching `wrk.result.states` in a callback on one rank only would leave that rank waiting)."""
        import torch

        local = np.ascontiguousarray(local)
        if not hasattr(self, "_row_counts") or self._row_counts[0] != n_total:
            counts = self.all_gather_object(int(local.shape[0]))  # once per problem: the shards never change
            self._row_counts = (n_total, counts)
        counts = self._row_counts[1]
        assert sum(counts) == n_total
        # array collective on the host group (no pickling: C5 moves 4 MB of states per call); shards are padded to the
        # largest one because all_gather wants equal shapes
        is_cplx = np.iscomplexobj(local)
        flat = local.astype(np.complex128 if is_cplx else np.float64, copy=False).reshape(local.shape[0], -1)
        width = flat.shape[1]
        rows = max(counts)
        buf = np.zeros((rows, width), flat.dtype)
        buf[: flat.shape[0]] = flat
        t = torch.from_numpy(buf.view(np.float64) if is_cplx else buf)
        outs = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(outs, t, group=self.cpu_group)
        parts = []
        for r, o in enumerate(outs):
            a = o.numpy()
            a = a.view(np.complex128) if is_cplx else a
            parts.append(a[: counts[r]])
        out = np.concatenate(parts, axis=0).reshape((n_total,) + local.shape[1:])
        return out

    def barrier(self):
        self.dist.barrier(group=self.cpu_group)
