"""world_size-2 gloo test (CPU) of the multi-GPU sweep logic: pair gating, partitioning, all-reduce of the per-scan
normal equations, redundant solve.  The compute back-end is a CPU stand-in built from the oracle (test
infrastructure); the product back-end (DeviceBackend) is exercised by tests/test_gpu_icp.py and bench.py."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    def __init__(self, scans, params):
        import oracle
        self.o, self.scans, self.params = oracle, scans, params

    def new_neq(self, n):
        import torch
        return torch.zeros(n * 28, dtype=torch.float64)

    def accumulate(self, pair_i, pair_j, poses, neq):
        o = self.o
        neq.zero_()
        out = neq.numpy().reshape(-1, 28)
        for i, j in zip(pair_i, pair_j):
            of1, t1 = o.matrix4_to_euler(poses[i])
            fg = o.transform_cloud(self.scans[i], o.euler_to_matrix(of1, t1))
            sg = o.transform_cloud(self.scans[j], o.euler_to_matrix(*o.matrix4_to_euler(poses[j])))
            nn, *_ = o.semantic_nn(fg, sg, self.params.search_radius, self.params.bucket_size)
            obs = o.build_observations(fg, self.scans[i], sg, nn)
            N, b = o.normal_equations(obs, [t1[0], t1[1], t1[2], of1[0], of1[1], of1[2]], 6)
            out[i, :21] += N[np.triu_indices(6)]
            out[i, 21:27] += b
            out[i, 27] += len(obs)

    def solve(self, neq, poses):
        o = self.o
        q = neq.numpy().reshape(-1, 28)
        new = np.zeros_like(poses)
        status = np.zeros(len(poses), dtype=np.int32)
        for s in range(len(poses)):
            of, t = o.matrix4_to_euler(poses[s])
            p6 = np.array([t[0], t[1], t[2], of[0], of[1], of[2]], dtype=np.float64)
            st = -4
            if q[s, 27] > self.params.obs_threshold:
                st, x = o.solve_packed(q[s], self.params.dof)
                if st == 0:
                    p6[:3] += x[:3]
                    if self.params.dof == 6:
                        p6[3:] += x[3:]
                    else:
                        p6[5] += x[3]
            status[s] = st
            new[s] = o.euler_to_matrix(p6[3:].astype(np.float32), p6[:3].astype(np.float32))
        return new, status


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import oracle
    pkg = importlib.import_module("mandala-mapping_b200")
    slam = importlib.import_module("mandala-mapping_b200.slam")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scans, truth, init = pkg.synth.slam_scans(5, kind="hdl32", seed=9, spacing=1.0, n_azimuth=128)
    prm = oracle.default_params(1.0, dof=4)
    drv = slam.SweepDriver(OracleBackend(scans, prm), [len(s) for s in scans], 10.0)
    poses, status = drv.sweep(init)
    if rank == 0:
        q.put((poses, status, drv.last_pairs))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_covers_all_pairs(pkg):
    slam = importlib.import_module("mandala-mapping_b200.slam")
    poses = pkg.synth.loop_trajectory(40, spacing=1.0).astype(np.float32)
    pi, pj = slam.gate_pairs(poses, 10.0)
    assert len(pi) > 40 and (pi != pj).all()
    d = np.linalg.norm(poses[pi, :3, 3] - poses[pj, :3, 3], axis=1)
    assert (d < 10.0).all()
    sizes = np.full(40, 65536)
    for world in (1, 2, 3, 8):
        parts = slam.partition_pairs(pi, pj, sizes, world)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(len(pi)))          # every pair exactly once
        load = np.array([len(p) for p in parts], dtype=float)
        assert load.max() <= 1.35 * load.mean() + 1


def test_two_rank_sweep_equals_single_process(pkg, oracle):
    import torch.multiprocessing as mp
    slam = importlib.import_module("mandala-mapping_b200.slam")
    scans, truth, init = pkg.synth.slam_scans(5, kind="hdl32", seed=9, spacing=1.0, n_azimuth=128)
    prm = oracle.default_params(1.0, dof=4)
    single, status1 = slam.SweepDriver(OracleBackend(scans, prm), [len(s) for s in scans], 10.0).sweep(init)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    poses2, status2, n_pairs = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert n_pairs == 20
    assert np.array_equal(status1, status2)
    assert np.abs(poses2 - single).max() < 1e-6
    # and both equal the oracle's own registerAll sweep
    poses_o, neq_o, status_o = oracle.register_all_sweep(scans, init, prm, pair_thr=10.0)
    assert np.array_equal(status_o, status1)
    assert np.abs(poses_o - single).max() < 1e-6


def test_cpp_plan_equals_python_plan(pkg):
    """m3dreg_slam_plan (the C++ pair gate + partition m3dreg_slam_sweep runs; pure host code) == slam.gate_pairs /
    slam.partition_pairs for several world sizes, scan sizes and gates; number_of_last_EOZ restricts i."""
    slam = importlib.import_module("mandala-mapping_b200.slam")
    rng = np.random.default_rng(3)
    for n, spread, thr in ((40, 15.0, 10.0), (100, 30.0, 10.0), (12, 2.0, 10.0), (30, 50.0, 5.0)):
        poses = np.tile(np.eye(4, dtype=np.float32), (n, 1, 1))
        poses[:, :3, 3] = rng.uniform(-spread, spread, (n, 3)).astype(np.float32)
        sizes = rng.integers(1000, 70000, n).astype(np.int32)
        pi, pj = slam.gate_pairs(poses, thr)
        for world in (1, 2, 3, 4, 8):
            a, b, ow = pkg.slam_plan(poses, sizes, thr, 0, world)
            assert np.array_equal(a, pi) and np.array_equal(b, pj)
            ow_py = np.zeros(len(pi), dtype=np.int32)
            for r, idx in enumerate(slam.partition_pairs(pi, pj, sizes, world)):
                ow_py[idx] = r
            assert np.array_equal(ow, ow_py), (n, world)
        a, b, ow = pkg.slam_plan(poses, sizes, thr, n - 3, 2)
        keep = pi >= n - 3
        assert np.array_equal(a, pi[keep]) and np.array_equal(b, pj[keep])


def test_measured_cost_plan_balances_what_point_counts_cannot(pkg):
    """m3dreg_slam_plan_measured: equal-sized scans whose groups cost very different device time (overlap differs) are
    balanced by the measured time; every pair keeps exactly one owner and groups stay whole unless heavier than a share."""
    synth = pkg.synth
    truth = synth.loop_trajectory(40, 1.0)
    sizes = np.full(40, 65536, dtype=np.int32)
    rng = np.random.default_rng(3)
    cpp = rng.uniform(0.01, 0.05, 40)
    cpp[5] = 0.0                                              # one group not measured: takes the mean of the others
    for world in (2, 4, 8):
        pi, pj, ow = pkg.slam_plan_measured(truth, sizes, cpp, 10.0, 0, world)
        pi0, pj0, ow0 = pkg.slam_plan(truth, sizes, 10.0, 0, world)
        assert np.array_equal(pi, pi0) and np.array_equal(pj, pj0) and ow.min() >= 0 and ow.max() < world
        eff = np.where(cpp > 0, cpp, cpp[cpp > 0].mean())
        load = np.array([eff[pi[ow == r]].sum() for r in range(world)])
        load0 = np.array([eff[pi[ow0 == r]].sum() for r in range(world)])
        assert load.max() / load.mean() < 1.06, (world, load)
        assert load.max() <= load0.max() + 1e-12            # never worse than the point-count plan under the measured costs
        for i in np.unique(pi):                              # groups stay whole (none is heavier than 1.25 shares here)
            assert len(np.unique(ow[pi == i])) == 1
