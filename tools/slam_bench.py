#!/usr/bin/env python
"""Multi-scan 6D SLAM sweep benchmark (BASELINE configs C4/C5 shape): registerAll over synthetic scans along a loop,
pairs sharded over the ranks, ONE NCCL all-reduce of the n_scans x 28 normal-equation blocks per sweep.

    python tools/slam_bench.py [--scans 100] [--kind hdl32|sick] [--sweeps 5] [--bucket 1.0] [--check]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/slam_bench.py ...

Prints one JSON line on rank 0: scans/s and points/s per sweep (device time, max over ranks), the all-reduce share,
and with --check the largest relative deviation of the distributed normal equations from a single-rank sweep.
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _claim_stdout():
    """Rank 0 prints exactly ONE JSON line on stdout: keep a private handle on the real stdout and send everything any
    library prints there (NCCL's version banner is a bare printf) to stderr."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=100)
    ap.add_argument("--kind", default="hdl32")
    ap.add_argument("--sweeps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--bucket", type=float, default=1.0)
    ap.add_argument("--dof", type=int, default=4)       # the reference's live call sites use the 4-DOF solver (gpu6DSLAM.cpp:406,575)
    ap.add_argument("--mode", default="icp", choices=["icp", "ndt"])
    ap.add_argument("--points", type=int, default=0, help="override points per scan (n_azimuth*32 for hdl32, n*n for sick)")
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("mandala-mapping_b200")
    slam = importlib.import_module("mandala-mapping_b200.slam")
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    kw = {}
    if args.points:
        if args.kind == "hdl32":
            kw["n_azimuth"] = max(8, args.points // 32)
        else:
            side = int(round(args.points ** 0.5)); kw["n_beams"] = side; kw["n_profiles"] = side
    t0 = time.perf_counter()
    scans, truth, init = pkg.synth.slam_scans(args.scans, kind=args.kind, seed=42, spacing=1.0, **kw)
    gen_s = time.perf_counter() - t0
    sizes = [len(s) for s in scans]
    prm = pkg.default_params(args.bucket, dof=args.dof, mode=pkg.MODE_NDT if args.mode == "ndt" else pkg.MODE_ICP)
    ctx = pkg.Context(local)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    for k, s in enumerate(scans):            # scans are replicated: only the 28-double blocks ever cross NVLink
        ctx.scan_upload(k, s)
    drv = slam.DeviceSweep(ctx, prm, 10.0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    poses = init.copy()
    for _ in range(args.warmup):
        drv.sweep(poses)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    status = None
    for _ in range(args.sweeps):
        poses, status = drv.sweep(poses)
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_sweep = float(ms.item()) / args.sweeps

    check = None
    if args.check:
        # distributed normal equations of ONE sweep from the initial poses vs the same sweep computed by this rank alone
        drv.sweep(init)
        a = ctx.slam_neq(len(init))
        pi, pj = slam.gate_pairs(init, 10.0)
        neq_solo = torch.zeros(len(init) * 28, dtype=torch.float64, device="cuda")
        ctx.sweep_zero(neq_solo, len(init))
        ctx.sweep_accumulate(pi, pj, np.ascontiguousarray(init, dtype=np.float32).reshape(-1, 4, 4), prm, neq_solo)
        ctx.synchronize()
        b = neq_solo.cpu().numpy().reshape(-1, 28)
        scale = np.abs(b[:, :27]).max(axis=1, keepdims=True) + 1e-300
        check = {"max_rel_dev_normal_equations": float((np.abs(a[:, :27] - b[:, :27]) / scale).max()),
                 "counts_identical": bool(np.array_equal(a[:, 27], b[:, 27]))}

    if rank == 0:
        err0 = float(np.abs(init[:, :3, 3] - truth[:, :3, 3]).max())
        err1 = float(np.abs(poses[:, :3, 3] - truth[:, :3, 3]).max())
        line = {"metric": "6DSLAM scans/sec (one Jacobi registerAll sweep)", "value": args.scans / (ms_sweep * 1e-3), "unit": "scans/s",
                "n_gpus": world, "sweeps": args.sweeps, "ms_per_sweep": ms_sweep, "scaling": "strong",
                "points_per_s": int(drv.last_stats.points_all) / (ms_sweep * 1e-3),
                "config": {"workload": f"{args.scans} synthetic {args.kind} scans x {sizes[0]} points along a loop, {args.bucket} m buckets, 10 m pair gate",
                           "pairs": int(drv.last_stats.n_pairs), "pairs_rank0": int(drv.last_stats.n_pairs_mine), "mode": args.mode, "dof": args.dof,
                           "collective": f"one all_reduce of {args.scans}x28 float64 per sweep ({args.scans * 28 * 8} bytes)"},
                "solved_scans": int((status == 0).sum()), "max_translation_error_m": {"initial": err0, "after": err1},
                "check": check, "scan_generation_s": gen_s}
        print(json.dumps(line), file=out, flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
