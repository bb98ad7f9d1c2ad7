#!/usr/bin/env python
"""bench.py — points/s per ICP iteration of the device-resident registration hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c1|c3] [--mode icp|ndt]

A "step" is one complete registration iteration over one scan pair: transform of the first scan by the current
pose + bounding box, bucket keys, radix sort, bucket table, semantic NN over the 27 neighbouring buckets,
normal-equation reduction, 6-DOF Cholesky solve and pose update — all on the device (m3dreg_icp_step).
N=1 runs BASELINE config C2 (1M-point rotating-SICK pair, 1.0 m buckets).  N>1 (under torchrun, one rank per
GPU) is the pair-sharded multi-scan case: every rank registers its own C2-sized pair and the 28-double normal
equation blocks are all-reduced with NCCL every step (weak scaling).

One JSON line is printed by rank 0; see the task contract for the keys.  `--impl reference` times the UNMODIFIED
reference kernels (oracle/_ref, its own malloc/copy/sync pattern, cudaWrapper.cpp:344-424,516-581) plus the host
glue the reference runs on the CPU each iteration; if that library is unavailable it times the CPU oracle port.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (generator kind, kwargs, bucket/radius, description)
    "c1": ("hdl32", {}, 0.5, "C1 two-scan ICP, 65 536-point HDL-32E-like pair, 0.5 m grid"),
    "c2": ("sick", {}, 1.0, "C2 two-scan registration, 1 048 576-point rotating-SICK pair, 1.0 m buckets"),
    "c3": ("sick", {"n_beams": 2048, "n_profiles": 2048}, 0.25, "C3 dense 4 194 304-point pair, 0.25 m grid"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region: NVML polled from a thread every millisecond (the timed region
    of a default run is a few milliseconds — a 100 ms nvidia-smi poll cannot see it), plus one reading right before and
    right after.  nvidia-smi is only the fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None
        self.nvml = None
        self.handle = None
        self.samples = []          # (sm MHz, reasons bitmask)
        self.edge = []             # readings right before / right after
        self.stop_flag = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            uuid = None
            try:
                import torch
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
            except Exception:
                uuid = None
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if isinstance(uuid, str) else uuid) if uuid else None
            except Exception:
                self.handle = None
            if self.handle is None:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nvml = None

    def _read(self):
        n = self.nvml
        mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            reasons = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            try:
                reasons = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            except Exception:
                reasons = 0
        return float(mhz), int(reasons)

    def start(self):
        if self.nvml:
            try:
                self.edge.append(self._read())
            except Exception:
                self.nvml = None
        if self.nvml:
            def poll():
                while not self.stop_flag:
                    try:
                        self.samples.append(self._read())
                    except Exception:
                        break
                    time.sleep(0.001)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        if self.nvml:
            self.stop_flag = True
            try:
                self.edge.append(self._read())
            except Exception:
                pass
            if self.thread:
                self.thread.join(timeout=1.0)
            n = self.nvml
            bits = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                    "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                    "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                    "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4))}
            allr = self.samples + self.edge
            reasons = sorted(k for k, b in bits.items() if any(r & b for _, r in allr))
            try:
                mx = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
            except Exception:
                mx = None
            sm = [m for m, _ in (self.samples or self.edge)]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(self.samples),
                    "sm_mhz_before_after": [m for m, _ in self.edge], "source": "NVML polled every ms during the timed region", "reasons": reasons}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["neither NVML nor nvidia-smi available"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, n in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "source": "nvidia-smi -lms 100", "reasons": sorted(reasons)}


def ncu_traffic(kernel, workload):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/roofline_traffic.json), else None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        e = json.load(open(p))[kernel]
        return int(e["dram_bytes_per_launch"]) if e.get("workload") == workload else None
    except Exception:
        return None


def alg_bytes(n1, n2, nb, nc):
    """Algorithmic bytes per ICP iteration: B_alg = 104*N1 + 36*N2 + 24*B + 40*Nc."""
    return 104 * n1 + 36 * n2 + 24 * nb + 40 * nc


def nn_alg_bytes(n1, n2, nb):
    """NN stage: table 8*N1 + first cloud 28*N1 + bucket table 12*B + queries 28*N2 + nn 4*N2."""
    return 36 * n1 + 32 * n2 + 12 * nb


def make_pair(pkg, workload, seed):
    kind, kw, res, _ = WORKLOADS[workload]
    first, second, pose_init, pose2, pose_true = pkg.synth.scan_pair(kind, seed=seed, **kw)
    return first, second, pose_init, pose2, pose_true, res


def cpu_baseline(first, second, pose_init, pose2, res, dof, budget_s=20.0, max_iters=4):
    """The oracle port (CPU restatement) timed on this box's host cores on a bounded sample of the same workload."""
    import oracle
    prm = oracle.default_params(res, dof=dof)
    sg = oracle.transform_cloud(second, oracle.euler_to_matrix(*oracle.matrix4_to_euler(pose2)))
    pose = pose_init.copy()
    t0 = time.perf_counter()
    iters = 0
    while iters < max_iters and (time.perf_counter() - t0) < budget_s:
        _, pose, _, _, _ = oracle.icp_iteration(first, sg, pose, prm)
        iters += 1
    dt = time.perf_counter() - t0
    pts = (len(first) + len(second)) * iters
    cores = oracle.lib().orc_num_threads()
    out = {"value": pts / dt, "unit": "points/s", "cores": cores, "kind": "port",
           "sample": f"{iters} full ICP iterations of the same pair (transform+grid+NN+obs+solve), {dt:.1f} s wall"}
    # also a single-thread figure (one iteration, bounded)
    try:
        oracle.lib().orc_set_num_threads(1)
        t1 = time.perf_counter()
        oracle.icp_iteration(first, sg, pose_init.copy(), prm)
        out["single_thread"] = {"value": (len(first) + len(second)) / (time.perf_counter() - t1), "unit": "points/s", "cores": 1,
                                "sample": "1 full ICP iteration of the same pair"}
    finally:
        oracle.lib().orc_set_num_threads(cores)
    return out


def run_reference(args, rank, world, out):
    """Reference arm: the reference's own kernels driven the way cudaWrapper.cpp / gpu6DSLAM.cpp drive them.

    The reference is single-threaded (one ROS spin loop, main.cpp:431-436), so its host glue — the per-iteration CPU
    transform of both clouds (gpu6DSLAM.cpp:635-663), the label counting and observation assembly (gpu6DSLAM.cpp:323-398)
    — runs on ONE thread here, whatever the box or the launcher.  The line carries a per-stage breakdown: host stages by
    wall clock, the reference's device calls by wall clock around each call (every one of them ends in
    cudaDeviceSynchronize or a blocking copy), so the kernel-versus-kernel ratios can be read off next to the whole-step
    ratio."""
    if rank != 0:
        return
    pkg = importlib.import_module("mandala-mapping_b200")
    import ctypes as C
    import oracle
    first, second, pose_init, pose2, _, res = make_pair(pkg, args.workload, args.seed)
    n1, n2 = len(first), len(second)
    n_pts = n1 + n2
    dof = args.dof
    use_ref = oracle.ref_available()
    kind = "reference"
    if use_ref:
        try:
            from tests import refwrap
            if oracle.ref().ref_device_count() <= 0:
                use_ref = False
            else:
                oracle.ref().ref_warm_up(0)
        except Exception:
            use_ref = False
    if not use_ref:
        kind = "port"
    threads_before = oracle.lib().orc_num_threads()
    oracle.lib().orc_set_num_threads(1)
    prm = oracle.default_params(res, dof=dof)
    pose = pose_init.copy()
    weights = (10.0, 1.0, 10.0, 10.0)
    host_ms = {"euler round trips + CPU transform of both clouds (gpu6DSLAM.cpp:276-307,635-663)": 0.0,
               "label counts + observation assembly (gpu6DSLAM.cpp:323-398)": 0.0}
    hk = list(host_ms)
    nb_last = [0]
    nc_last = [0]

    def step(pose, timed):
        # gpu6DSLAM.cpp:276-307: Euler round trips + CPU transform of BOTH clouds every iteration
        t0 = time.perf_counter()
        o1, t1 = oracle.matrix4_to_euler(pose)
        p1 = oracle.euler_to_matrix(o1, t1)
        fg = oracle.transform_cloud(first, p1)
        sg = oracle.transform_cloud(second, oracle.euler_to_matrix(*oracle.matrix4_to_euler(pose2)))
        t1c = time.perf_counter()
        if use_ref:
            nn, gp, *_ = refwrap.nn_search_host(fg, sg, res, res, 1.0, 100, 100, export=False)     # cudaWrapper.cpp:344-424
        else:
            nn, gp, *_ = oracle.semantic_nn(fg, sg, res, res, 1.0, 100, 100)
        nb_last[0] = int(gp["number_of_buckets"][0])
        t2 = time.perf_counter()
        obs = oracle.build_observations(fg, first, sg, nn, weights)                              # gpu6DSLAM.cpp:323-398
        nc_last[0] = len(obs)
        t3 = time.perf_counter()
        if timed:
            host_ms[hk[0]] += 1e3 * (t1c - t0)
            host_ms[hk[1]] += 1e3 * (t3 - t2)
        pose6 = [t1[0], t1[1], t1[2], o1[0], o1[1], o1[2]]
        if len(obs) > 100:
            if use_ref:
                st, p6, _ = refwrap.register_ls_host(obs, pose6, dof)                           # cudaWrapper.cpp:516-648
            else:
                st, p6, _ = oracle.register_ls(obs, pose6, dof)
            if st == 0:
                pose = oracle.euler_to_matrix(np.float32(p6[3:]), np.float32(p6[:3]))
        return pose

    # bounded: the CPU port needs seconds per step on 1M points
    steps = args.steps if use_ref else min(args.steps, 3)
    warm = args.warmup if use_ref else min(args.warmup, 1)
    for _ in range(warm):
        pose = step(pose, False)
    if use_ref:
        oracle.ref().ref_get_stage_ms(None, None, C.c_int(1))
    t0 = time.perf_counter()
    for _ in range(steps):
        pose = step(pose, True)
    dt = time.perf_counter() - t0
    stage_ms = {k: v / steps for k, v in host_ms.items()}
    if use_ref:
        ms = (C.c_double * 8)()
        calls = (C.c_int * 2)()
        oracle.ref().ref_get_stage_ms(ms, calls, C.c_int(1))
        names = ["cudaMalloc + H2D of both 40-B clouds (cudaWrapper.cpp:360-372)", "cudaCalculateGridParams (lesson_16.cu:23-106)",
                 "cudaMalloc x3 + cudaCalculateGrid (lesson_16.cu:200-243)", "cudaSemanticNearestNeighborSearch (lesson_16.cu:531-738)",
                 "D2H of nn + cudaFree x5 (cudaWrapper.cpp:406-420)", "cudaMalloc x4 + H2D of the observations (cudaWrapper.cpp:523-536)",
                 "fill_A_l_cuda (lesson_16.cu:355-439)", "Solve_ATPA_ATPl_x: AtP + 2x DGEMM + potrf/potrs + handles + frees (CCUDAAXBSolverWrapper.cpp:407-539)"]
        for k in range(8):
            stage_ms[names[k]] = ms[k] / steps
    oracle.lib().orc_set_num_threads(threads_before)
    value = n_pts * steps / dt
    line = {
        "impl": "reference", "metric": "points/sec per ICP iteration", "value": value, "unit": "points/s", "n_gpus": 1,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 geometry / f64 normal equations", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][3], "mode": "icp", "dof": dof, "n_first": n1, "n_second": n2,
                   "buckets": nb_last[0], "correspondences": nc_last[0], "search_radius_m": res, "bucket_m": res, "max_inner": 100, "max_outer": 100,
                   "l2_policy": "n/a (every device buffer is allocated, filled over PCIe and freed inside each call)", "parallelism": "pairs1",
                   "reference_path": ("verbatim reference CUDA kernels (lesson_16.cu, CCUDAAXBSolverWrapper.cpp compiled for sm_100a) through the "
                                      "reference's per-call malloc/H2D/D2H/free pattern + its per-iteration CPU transform and observation assembly on ONE "
                                      "host thread, as upstream (the reference has NO CPU implementation of this path)") if use_ref else
                                     "CPU oracle port (reference kernels unavailable on this box), one thread"},
        "stage_ms": stage_ms,
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": 1, "kind": kind,
                         "sample": f"{steps} full iterations of the same pair, host glue on 1 thread (the reference is single-threaded)"},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=out, flush=True)


def slam_record(pkg, which, args, rank, world, local_rank):
    """Multi-scan 6D SLAM record (BASELINE configs C4 / C5 shape): registerAll sweeps through the C entry point
    m3dreg_slam_sweep — pairs gated at 10 m and sharded over the ranks, ONE NCCL all-reduce of the n_scans x 28
    normal-equation blocks per sweep inside the library, redundant solve.  STRONG scaling: the scan set is fixed, the ranks
    share its pairs.  Scan generation is spread over the ranks and exchanged (it is numpy ray casting, not the product)."""
    import torch
    import torch.distributed as dist
    slam = importlib.import_module("mandala-mapping_b200.slam")
    if which == "c4":
        n_scans, kind, kw, res, modes, dof = args.slam_scans, "hdl32", {}, 1.0, ["icp"], 4
        desc = f"C4: registerAll over {n_scans} synthetic HDL-32E scans x 65 536 points along a loop, 1.0 m buckets, 10 m pair gate, 4-DOF (the reference's live solver, gpu6DSLAM.cpp:575)"
    else:
        n_scans, kind, kw, res, modes, dof = args.slam_c5_scans, "sick", {}, 1.0, ["icp", "ndt"], 6
        desc = (f"C5 shape: registerAll over {n_scans} synthetic rotating-SICK scans x 1 048 576 points, 1.0 m buckets, 10 m pair gate, "
                "ICP and NDT sweeps alternating (BASELINE's 1000 scans do not fit a bench run: the scans are numpy ray casts, 2.4 s each, "
                "spread over worker processes)")
    t0 = time.perf_counter()
    mine = [k for k in range(n_scans) if k % world == rank]
    gen_workers = max(1, min(args.slam_gen_workers, (os.cpu_count() or 1) // max(1, world)))
    scans, truth, init = pkg.synth.slam_scans(n_scans, kind=kind, seed=42, spacing=1.0, only=mine, workers=gen_workers, **kw)
    ctx = pkg.Context(local_rank)
    npts = len(scans[mine[0]])
    if world > 1:
        per = (n_scans + world - 1) // world
        loc = torch.zeros((per, npts * 40), dtype=torch.uint8, device="cuda")
        for q, k in enumerate(mine):
            assert len(scans[k]) == npts
            loc[q] = torch.from_numpy(np.frombuffer(scans[k].tobytes(), dtype=np.uint8).copy()).cuda()
        allb = [torch.empty_like(loc) for _ in range(world)]
        dist.all_gather(allb, loc)
        torch.cuda.synchronize()
        for k in range(n_scans):
            ctx.scan_upload(k, allb[k % world][k // world].data_ptr(), n=npts, on_device=True)
        del allb, loc
    else:
        for k in range(n_scans):
            ctx.scan_upload(k, scans[k])
    gen_s = time.perf_counter() - t0
    drivers = {m: slam.DeviceSweep(ctx, pkg.default_params(res, dof=dof, mode=pkg.MODE_NDT if m == "ndt" else pkg.MODE_ICP), 10.0) for m in modes[:1]}
    for m in modes[1:]:      # the communicator lives in the context: later drivers reuse it
        d = slam.DeviceSweep.__new__(slam.DeviceSweep)
        d.ctx, d.params, d.threshold, d.first_optimised = ctx, pkg.default_params(res, dof=dof, mode=pkg.MODE_NDT if m == "ndt" else pkg.MODE_ICP), 10.0, 0
        d.rank, d.world, d.last_stats = rank, world, None
        drivers[m] = d

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    poses = init.copy()
    for m in modes:                       # warm-up: one sweep per mode from the initial poses (result discarded)
        drivers[m].sweep(init)
    err = [pkg.synth.relative_pose_error(poses, truth)]
    acc_ms, red_ms, wall = [], [], []
    status = None
    barrier()
    for s in range(args.slam_sweeps):
        m = modes[s % len(modes)]
        barrier()
        t1 = time.perf_counter()
        poses, status = drivers[m].sweep(poses)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t1
        st = drivers[m].last_stats
        v = torch.tensor([dt * 1e3, st.accumulate_ms, st.allreduce_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        wall.append(float(v[0])); acc_ms.append(float(v[1])); red_ms.append(float(v[2]))
        err.append(pkg.synth.relative_pose_error(poses, truth))
    st = drivers[modes[0]].last_stats
    ms_sweep = float(np.mean(wall))
    # sharded == unsharded: the all-reduced blocks of a sweep from the initial poses vs the same rows accumulated by this rank alone
    drivers[modes[0]].sweep(init)
    neq = ctx.slam_neq(n_scans)
    rows = sorted(set([0, n_scans // 3, n_scans - 1]))
    pi, pj, _ = pkg.slam_plan(init, [npts] * n_scans, 10.0, 0, 1)
    sel = np.isin(pi, rows)
    solo = torch.zeros(n_scans * 28, dtype=torch.float64, device="cuda")
    ctx.sweep_zero(solo, n_scans)
    ctx.sweep_accumulate(pi[sel], pj[sel], init, drivers[modes[0]].params, solo)
    ctx.synchronize()
    b = solo.cpu().numpy().reshape(-1, 28)[rows]
    a = neq[rows]
    scale = np.abs(b[:, :27]).max(axis=1, keepdims=True) + 1e-300
    dev = float((np.abs(a[:, :27] - b[:, :27]) / scale).max())
    chk = torch.tensor([dev, 0.0 if np.array_equal(a[:, 27], b[:, 27]) else 1.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(chk, op=dist.ReduceOp.MAX)
    ctx.close()
    return {
        "workload": desc, "scaling": "strong", "n_scans": n_scans, "points_per_scan": npts, "pairs": int(st.n_pairs), "pairs_rank0": int(st.n_pairs_mine),
        "modes": modes, "dof": dof, "sweeps": args.slam_sweeps, "ms_per_sweep": ms_sweep, "scans_per_s": n_scans / (ms_sweep * 1e-3),
        "points_per_s": int(st.points_all) / (ms_sweep * 1e-3),
        "accumulate_ms_max_rank": float(np.mean(acc_ms)), "allreduce_wait_ms_max_rank": float(np.mean(red_ms)),
        "timing": "host wall clock around m3dreg_slam_sweep (gate + partition + accumulate + NCCL all-reduce + solve + pose read-back), max over ranks, mean over sweeps",
        "collective": f"one ncclAllReduce of {n_scans} x 28 float64 ({n_scans * 224} bytes) per sweep, issued by the library on its stream",
        "sharded_vs_single_rank": {"rows": rows, "max_rel_dev_normal_equations": float(chk[0]), "counts_identical": bool(chk[1] == 0.0)},
        "relative_pose_error_m": {"initial": err[0], "per_sweep": err[1:]},
        "solved_scans": int((status == 0).sum()), "scan_generation_and_upload_s": gen_s,
    }


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
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="icp", choices=["icp", "ndt"])
    ap.add_argument("--dof", type=int, default=6, choices=[4, 6])
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--slam", default="c4", help="comma list of multi-scan records to add to the line: c4 (100 HDL-32E scans x 65 536, "
                    "registerAll strong scaling), c5 (rotating-SICK scans x 1 048 576, ICP and NDT sweeps alternating), none")
    ap.add_argument("--slam-scans", type=int, default=100)
    ap.add_argument("--slam-c5-scans", type=int, default=16)
    ap.add_argument("--slam-sweeps", type=int, default=6)
    ap.add_argument("--slam-gen-workers", type=int, default=16, help="processes that ray-cast the synthetic scans (per rank; capped by the host's cores)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world, out)
        return

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("mandala-mapping_b200")
    assert torch.cuda.is_available(), "bench.py needs a B200 (the product has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    peak_gbs, peak_src = load_peaks()
    first, second, pose_init, pose2, pose_true, res = make_pair(pkg, args.workload, args.seed + rank)
    n1, n2 = len(first), len(second)
    prm = pkg.default_params(res, dof=args.dof, mode=pkg.MODE_NDT if args.mode == "ndt" else pkg.MODE_ICP)

    ctx = pkg.Context(local_rank)
    # everything (kernels, NCCL, timing events) runs on ONE explicit non-default stream
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.scan_upload(0, first)
    ctx.scan_upload(1, second)
    neq_all = torch.zeros(world * 28, dtype=torch.float64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1: the all-reduce of the normal-equation blocks runs on its own stream and overlaps the next iteration (nothing
    # in the registration of a pair waits for it: in a sweep the reduced blocks are consumed once, by the final solve).
    # Two buffers; the last block of the normal-equation kernel writes the rank's block straight into the buffer
    # (m3dreg_icp_set_neq_out), and every all-reduce is complete before the closing event of the timed region (drain()).
    comm = torch.cuda.Stream() if world > 1 else None
    bufs = [torch.zeros(world * 28, dtype=torch.float64, device="cuda") for _ in range(2)] if world > 1 else []
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_zeroed = [torch.cuda.Event() for _ in range(2)]
    ev_reduced = [torch.cuda.Event() for _ in range(2)]
    step_no = [0]
    if world > 1:
        for k in range(2):
            ev_zeroed[k].record(stream)
            ev_reduced[k].record(stream)

    def one_step():
        if world > 1:
            k = step_no[0] % 2
            step_no[0] += 1
            stream.wait_event(ev_zeroed[k])                     # buffer k was reduced and cleared two steps ago
            ctx.icp_set_neq_out(bufs[k][rank * 28:])            # the kernel that forms the block writes it there
        ctx.icp_step(1)
        if world > 1:
            ev_copied[k].record(stream)
            with torch.cuda.stream(comm):
                comm.wait_event(ev_copied[k])
                dist.all_reduce(bufs[k])
                ev_reduced[k].record(comm)
                neq_all.copy_(bufs[k])                          # the latest reduced blocks, for whoever consumes them
                bufs[k].zero_()
                ev_zeroed[k].record(comm)

    def drain():
        if world > 1:
            for k in range(2):
                stream.wait_event(ev_zeroed[k])
            ctx.icp_set_neq_out(None)

    # ---------------- device-resident timing: value ----------------
    ctx.icp_begin(0, 1, pose_init, pose2, prm)
    launches0 = ctx.launch_count
    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches1 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        one_step()
    drain()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    allreduce_ok = None
    if world > 1:     # the last reduced blocks must hold this rank's last block bit for bit, and every other rank's too
        mine = torch.zeros(28, dtype=torch.float64, device="cuda")
        ctx.icp_copy_neq(mine)
        torch.cuda.synchronize()
        ok = torch.equal(neq_all[rank * 28:(rank + 1) * 28], mine) and bool((neq_all.view(world, 28)[:, 27] > 0).all())
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        allreduce_ok = bool(flag.item())
    launches = ctx.launch_count - launches1
    clocks = sampler.stop() if rank == 0 else None
    pose_out, st = ctx.icp_end()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    value = world * (n1 + n2) / (ms_per_step * 1e-3)
    nb = int(st.n_buckets_last)
    nc = int(st.n_obs_last)

    # ---------------- per-stage event timing: roofline of the dominant kernel ----------------
    ctx.icp_begin(0, 1, pose_out, pose2, prm)
    ctx.set_profiling(True)
    ctx.icp_step(2)                      # the profiling variants of the kernels are loaded lazily on first use: keep that out of the stage times
    ctx.get_stage_ms()
    ctx.nn_evaluations(reset=True)
    ctx.nn_fallbacks(reset=True)
    ctx.icp_step(min(args.steps, 10))
    stage_ms, stage_iters = ctx.get_stage_ms()
    try:
        gph = ctx.grid_phase_ns().astype(np.int64)
        names = ["box", "barrier1", "keys+hist", "barrier2", "sort pass 1 + count scan", "barrier3", "bucket table + sort pass 2", "barrier4",
                 "(sort pass 3)", "candidate sets"]
        grid_phases_us = {names[k]: float(gph[k + 1] - gph[k]) / 1e3 for k in range(10)} if gph[10] > gph[0] > 0 else None
        if grid_phases_us and gph[21] > gph[15] > 0:
            sub = ["digit bases", "ranks (match)", "warp scan", "scatter", "next histogram"]
            grid_phases_us["sort pass 1 detail"] = dict(**{"count scan": float(gph[15] - gph[4]) / 1e3},
                                                        **{sub[k]: float(gph[17 + k] - gph[16 + k]) / 1e3 for k in range(5)})
        if grid_phases_us and gph[27] > gph[22] > 0:
            subc = ["list + bucket record", "setup", "sweep 1 (bin)", "table scan", "sweep 2 (place)"]
            grid_phases_us["one bucket detail"] = dict(**{subc[k]: float(gph[23 + k] - gph[22 + k]) / 1e3 for k in range(5)},
                                                       **{"candidates": int(gph[29] & 0xffffffff), "points": int(gph[29] >> 32), "start after phase begin": float(gph[22] - gph[9]) / 1e3})
    except Exception:
        grid_phases_us = None
    evals_per_query = ctx.nn_evaluations(reset=True) / max(stage_iters, 1) / float(n2)
    fallback_share = ctx.nn_fallbacks(reset=True) / max(stage_iters, 1) / float(n2)
    ctx.set_profiling(False)
    ctx.icp_end()
    stage_ms = stage_ms / max(stage_iters, 1)
    stage_names = ["box pass (transform in registers + bounding box)", "grid (keys, radix sort, bucket table, candidate sets)",
                   "semantic NN (k_nn_search_hull)", "normal equations + solve"]
    dom = int(np.argmax(stage_ms))
    if args.mode == "icp":
        # the dominant kernel of the ICP iteration is the semantic search: ONE launch per iteration = the whole stage
        roof_kernel, roof_stage = "k_nn_search_hull", 2
        nn_bytes = nn_alg_bytes(n1, n2, nb)
    else:
        # NDT (no reference implementation): no search kernel runs; report the stage that dominates, with SURVEY 8d's stage bytes
        # (grid: 20*N1 + 12*B plus the 12*N1 coordinates the per-bucket statistics read; queries: 16*N2; reduction: 4*N2 + 40*Nc)
        roof_stage = dom
        roof_kernel = ["k_transform_soa<true>", "k_grid_head + k_radix_scan/scatter + k_finalize_grid + k_ndt_accumulate_points + k_ndt_finalize_buckets",
                       "k_ndt_accumulate_queries", "k_ndt_accumulate_queries + k_ndt_normal_equations"][dom]
        nn_bytes = [48 * n1, 32 * n1 + 12 * nb, 16 * n2, 20 * n2 + 40 * nc][dom]
    nn_ms = float(stage_ms[roof_stage])
    nn_gbs = nn_bytes / (nn_ms * 1e-3) / 1e9
    iter_bytes = alg_bytes(n1, n2, nb, nc)
    iter_gbs = iter_bytes / (ms_per_step * 1e-3) / 1e9

    # ---------------- end-to-end through the host-buffer C ABI call (H2D/D2H inside the timed region) ----------------
    h_first = torch.from_numpy(np.frombuffer(first.tobytes(), dtype=np.uint8).copy()).pin_memory()
    # second cloud already in the global frame, as the reference hands it over (gpu6DSLAM.cpp:306-307)
    d_tmp = torch.from_numpy(np.frombuffer(second.tobytes(), dtype=np.uint8).copy()).cuda()
    d_out = torch.empty_like(d_tmp)
    o2, t2 = pkg.matrix4_to_euler(pose2)
    ctx.transform(d_tmp, d_out, n2, pkg.euler_to_matrix(o2, t2))
    ctx.synchronize()
    h_second = d_out.cpu().pin_memory()
    del d_tmp, d_out
    h_nn = torch.empty(n2, dtype=torch.int32).pin_memory()
    pose_e2e = np.ascontiguousarray(pose_init, dtype=np.float32).reshape(16).copy()
    e2e_steps = max(1, args.e2e_steps)
    for _ in range(2):
        _e2e_call(ctx, pkg, h_first, n1, h_second, n2, pose_e2e, prm, h_nn)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _e2e_call(ctx, pkg, h_first, n1, h_second, n2, pose_e2e, prm, h_nn)
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    te = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * (n1 + n2) * e2e_steps / float(te.item())

    # ---------------- the reference's schedule from the initial perturbation (gpu6DSLAM.cpp:159-172) ----------------
    # The timed steps above run after the warm-up iterations, i.e. on a nearly aligned pair (the search's best case).  This
    # record times the three steps of registerLastArrivedScan as upstream runs them: 30 iterations each at radius = bucket =
    # 2.5 / 2.0 / 1.0 m, 4-DOF, starting from the SURVEY 8d perturbation: unconverged correspondences and larger buckets.
    schedule = None
    if args.mode == "icp" and world == 1:
        try:
            pose_s = np.ascontiguousarray(pose_init, dtype=np.float32).copy()
            schedule = {"dof": 4, "steps": []}
            for rb in (2.5, 2.0, 1.0):
                sp = pkg.default_params(rb, dof=4)
                pose_w, _ = ctx.icp_pair(0, 1, pose_s, pose2, sp, 1)          # buffers for this bucket size: outside the timed call
                pose_s, sst = ctx.icp_pair(0, 1, pose_s, pose2, sp, 30)
                schedule["steps"].append({"radius_m": rb, "bucket_m": rb, "iterations": int(sst.iterations_run), "us_per_iteration": float(sst.device_ms) * 1e3 / max(int(sst.iterations_run), 1),
                                          "buckets": int(sst.n_buckets_last), "correspondences": int(sst.n_obs_last),
                                          "translation_error_m": float(np.abs(pose_s[:3, 3] - pose_true[:3, 3]).max())})
            schedule["total_ms"] = float(sum(s["us_per_iteration"] * s["iterations"] for s in schedule["steps"]) / 1e3)
        except Exception as ex:  # pragma: no cover
            schedule = {"failed": repr(ex)}

    ctx.close()
    slam_out = {}
    for which in [w for w in args.slam.split(",") if w in ("c4", "c5")]:
        try:
            slam_out[which] = slam_record(pkg, which, args, rank, world, local_rank)
        except Exception as ex:  # pragma: no cover
            slam_out[which] = {"failed": repr(ex)}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            try:
                cpu = cpu_baseline(first, second, pose_init, pose2, res, args.dof)
            except Exception as ex:  # pragma: no cover
                cpu = {"value": None, "unit": "points/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
        working_set = n1 * (32 * 3 + 16 + 16) + n2 * (32 + 32 + 4) + nb * 12
        line = {
            "metric": "points/sec per ICP iteration", "value": value, "unit": "points/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 geometry / f64 normal equations", "data": "synthetic",
            "config": {
                "workload": WORKLOADS[args.workload][3] + (f"; one pair per rank x {world} ranks, NCCL all-reduce of the 28-double normal-equation blocks per step on a second stream (overlaps the next step, drained inside the timed region)" if world > 1 else ""),
                "mode": args.mode, "dof": args.dof, "n_first": n1, "n_second": n2, "buckets": nb, "correspondences": nc,
                "search_radius_m": res, "bucket_m": res, "max_inner": 100, "max_outer": 100,
                "l2_policy": f"no flush: per-iteration working set {working_set / 1e6:.0f} MB " + ("exceeds" if working_set > 126e6 else "is below") + " the 126 MB L2",
                "parallelism": f"pairs{world}",
            },
            "allreduce_check": allreduce_ok,
            "gpu_launches": int(launches),
            "launches_per_step": launches / args.steps,
            "clocks": clocks,
            "roofline": {
                "bound": "hbm", "kernel": roof_kernel, "achieved": nn_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": nn_gbs / peak_gbs,
                "traffic": ncu_traffic("k_nn_search_hull", args.workload) if args.mode == "icp" else None, "peak_source": peak_src, "algorithmic_bytes_per_launch": nn_bytes, "launch_ms": nn_ms,
                "dominant_stage": stage_names[dom],
                "nn_candidate_evaluations_per_query": evals_per_query,
                "nn_queries_on_per_thread_fallback": fallback_share,
                "stage_ms": {stage_names[k]: float(stage_ms[k]) for k in range(4)},
                "grid_phases_us": grid_phases_us,
                "iteration": {"algorithmic_bytes": iter_bytes, "achieved": iter_gbs, "frac": iter_gbs / peak_gbs,
                              "note": "whole iteration (all kernels) vs the HBM roofline, B_alg = 104*N1 + 36*N2 + 24*B + 40*Nc"},
            },
            "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": 40 * (n1 + n2), "d2h_bytes_per_step": 4 * n2 + 64,
                    "steps": e2e_steps, "call": "m3dreg_icp_iteration_host (both 40-B clouds H2D from pinned memory, nn + pose D2H, every step)"},
            "cpu_baseline": cpu,
            "schedule": schedule,
            "slam": slam_out or None,
            "result": {"status": int(st.last_status), "translation_error_m": float(np.abs(pose_out[:3, 3] - pose_true[:3, 3]).max())},
        }
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


def _e2e_call(ctx, pkg, h_first, n1, h_second, n2, pose, prm, h_nn):
    import ctypes as C
    st = pkg.IcpStats()
    rc = pkg.lib().m3dreg_icp_iteration_host(ctx._h, C.c_void_p(h_first.data_ptr()), C.c_int(n1), C.c_void_p(h_second.data_ptr()),
                                             C.c_int(n2), pkg._p(pose), C.byref(prm), C.c_void_p(h_nn.data_ptr()), C.byref(st))
    if rc != 0:
        raise pkg.M3dRegError(rc, "m3dreg_icp_iteration_host")
    return st


if __name__ == "__main__":
    main()
