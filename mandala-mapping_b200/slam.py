"""Multi-scan 6D SLAM sweep (gpu6DSLAM::registerAll, src/gpu6DSLAM.cpp:424-597) sharded over the GPUs of one box.

One process per GPU (torchrun).  Within a Jacobi sweep every (i, j) pair reads only the OLD poses and adds a
28-double block (21 upper-triangular AtPA, 6 AtPl, observation count) to scan i's normal equations, so pairs are the
unit of work:

    pairs  = gate_pairs(poses, 10 m)                       # gpu6DSLAM.cpp:454-469
    mine   = partition_pairs(pairs, scan sizes, world)[rank]
    neq    = zeros(n_scans, 28) on the device
    ctx.sweep_accumulate(mine, poses, params, neq)         # grid of scan i built once per i, NN, fused reduction
    dist.all_reduce(neq)                                   # NCCL over NVLink/NVSwitch: the ONLY exchange step
    poses  = ctx.sweep_solve(neq, poses, params)           # every rank solves all scans (6x6 Cholesky each): no gather

The PRODUCT path is :class:`DeviceSweep`: one call of the C entry point ``m3dreg_slam_sweep`` per sweep — gate,
partition, accumulation, the NCCL all-reduce (NCCL resolved by the library itself, communicator created from an id that
rank 0 broadcasts) and the solve all happen inside ``libm3dreg.so`` on the context's stream; Python only hands the
128-byte id around.  :func:`gate_pairs` / :func:`partition_pairs` / :class:`SweepDriver` restate the same plan in
Python with an injected back-end so the sharding logic can be exercised on CPU (gloo) in the tests, and the C++ plan
(``m3dreg_slam_plan``) is checked against them.
"""
from __future__ import annotations

import numpy as np


def gate_pairs(poses: np.ndarray, distance_threshold: float = 10.0):
    """All ordered pairs i != j whose pose translations are closer than the threshold (float arithmetic as upstream)."""
    p = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, 4, 4)
    t = p[:, :3, 3]
    d = t[:, None, :] - t[None, :, :]
    dist = np.sqrt((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]).astype(np.float32))
    ii, jj = np.nonzero((dist < np.float32(distance_threshold)) & ~np.eye(len(p), dtype=bool))
    return ii.astype(np.int32), jj.astype(np.int32)


def partition_pairs(pair_i, pair_j, sizes, world: int):
    """Assign pairs to ranks: whole groups of equal i first (the grid of scan i is then built once), heaviest group to
    the least loaded rank; groups heavier than the ideal share are split.  Returns a list of index arrays."""
    pair_i = np.asarray(pair_i)
    pair_j = np.asarray(pair_j)
    sizes = np.asarray(sizes, dtype=np.float64)
    cost = sizes[pair_i] + sizes[pair_j]
    total = float(cost.sum() + sizes[np.unique(pair_i)].sum()) if len(pair_i) else 0.0
    ideal = total / max(world, 1)
    groups = []
    for i in np.unique(pair_i):
        idx = np.nonzero(pair_i == i)[0]
        c = float(cost[idx].sum() + sizes[i])
        if world > 1 and c > 1.25 * ideal and len(idx) > 1:
            parts = int(min(len(idx), np.ceil(c / ideal)))
            for chunk in np.array_split(idx, parts):
                groups.append((float(cost[chunk].sum() + sizes[i]), chunk))
        else:
            groups.append((c, idx))
    groups.sort(key=lambda g: -g[0])
    load = np.zeros(world)
    out = [[] for _ in range(world)]
    for c, idx in groups:
        r = int(np.argmin(load))
        load[r] += c
        out[r].append(idx)
    return [np.sort(np.concatenate(o)) if o else np.zeros(0, dtype=np.int64) for o in out]


class DeviceSweep:
    """registerAll over the GPUs of one box through ``m3dreg_slam_sweep`` (one context per rank, scans uploaded to
    slots 0..n-1 of every context).  With torch.distributed initialised (any backend) the NCCL communicator of the
    library is created from an id broadcast from rank 0; without it the sweep runs on one GPU."""

    def __init__(self, ctx, params, distance_threshold: float = 10.0, first_optimised: int = 0):
        self.ctx, self.params = ctx, params
        self.threshold, self.first_optimised = distance_threshold, first_optimised
        self.rank, self.world = 0, 1
        self.last_stats = None
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.rank, self.world = dist.get_rank(), dist.get_world_size()
        except Exception:
            pass
        if self.world > 1:
            import importlib
            import torch
            import torch.distributed as dist
            pkg = importlib.import_module(__package__)
            uid = torch.zeros(128, dtype=torch.uint8)
            if self.rank == 0:
                uid = torch.from_numpy(pkg.nccl_unique_id().copy())
            dev = torch.device("cuda", ctx.device) if dist.get_backend() == "nccl" else torch.device("cpu")
            uid = uid.to(dev)
            dist.broadcast(uid, src=0)
            ctx.nccl_init(uid.cpu().numpy(), self.rank, self.world)

    def sweep(self, poses: np.ndarray):
        """One Jacobi sweep; returns (new_poses [n,4,4] float32, status [n] int32)."""
        new_poses, status, st = self.ctx.slam_sweep(poses, self.params, self.threshold, self.first_optimised)
        self.last_stats = st
        return new_poses, status


class SweepDriver:
    """The same sweep with the compute back-end injected (tests: a CPU oracle back-end over gloo).  The back-end provides
    new_neq(n), accumulate(pair_i, pair_j, poses, neq) and solve(neq, poses)."""

    def __init__(self, backend, sizes, distance_threshold: float = 10.0, process_group=None):
        self.backend = backend
        self.sizes = np.asarray(sizes)
        self.threshold = distance_threshold
        self.group = process_group
        self.rank, self.world = 0, 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.rank, self.world = dist.get_rank(process_group), dist.get_world_size(process_group)
        except Exception:
            pass
        self.last_pairs = 0
        self.last_points = 0

    def sweep(self, poses: np.ndarray):
        """One Jacobi sweep; returns (new_poses [n,4,4] float32, status [n] int32)."""
        poses = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, 4, 4)
        n = len(poses)
        pi, pj = gate_pairs(poses, self.threshold)
        mine = partition_pairs(pi, pj, self.sizes, self.world)[self.rank]
        neq = self.backend.new_neq(n)
        self.backend.accumulate(pi[mine], pj[mine], poses, neq)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(neq, group=self.group)
        self.last_pairs = len(pi)
        self.last_my_pairs = len(mine)
        self.last_points = int((self.sizes[pi] + self.sizes[pj]).sum())
        self.last_neq = neq          # all-reduced per-scan normal equations (n_scans x 28), kept for checks
        return self.backend.solve(neq, poses)
