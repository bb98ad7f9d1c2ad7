"""GPU tests of the fused, device-resident loops against the CPU oracle loops."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _second_global(oracle, second, pose2):
    return oracle.transform_cloud(second, oracle.euler_to_matrix(*oracle.matrix4_to_euler(pose2)))


@pytest.mark.parametrize("dof", [6, 4])
def test_icp_pair_tracks_oracle(pkg, oracle, ctx, hdl_pair_small, dof):
    first, second, pose_init, pose2, pose_true = hdl_pair_small
    ctx.scan_clear()
    ctx.scan_upload(0, first)
    ctx.scan_upload(1, second)
    prm = pkg.default_params(0.5, dof=dof)
    oprm = oracle.default_params(0.5, dof=dof)
    sg = _second_global(oracle, second, pose2)
    pose_d, pose_o = pose_init.copy(), pose_init.copy()
    agree = []
    for it in range(12):
        pose_d, st = ctx.icp_pair(0, 1, pose_d, pose2, prm, 1)
        nn_d = ctx.export_last_nn(len(second))
        status, pose_o, n_obs, x_o, nn_o = oracle.icp_iteration(first, sg, pose_o, oprm, want_nn=True)
        assert st.last_status == 0 and status == 0
        agree.append(float((nn_d == nn_o).mean()))
        assert st.n_obs_last == n_obs or agree[-1] < 1.0
    # north-star: >= 85 % identical correspondences in the end-to-end loop
    assert min(agree) >= 0.85, agree
    # final poses: 1e-5 m / 1e-6 rad
    o_d, t_d = pkg.matrix4_to_euler(pose_d)
    o_o, t_o = oracle.matrix4_to_euler(pose_o)
    assert np.abs(t_d - t_o).max() < 1e-5, (t_d, t_o, agree)
    assert np.abs(o_d - o_o).max() < 1e-6, (o_d, o_o, agree)


def test_icp_pair_multi_iteration_equals_single_steps(pkg, ctx, hdl_pair_small):
    """iterations=N in one call == N calls with iterations=1 (state is carried as the float 4x4 either way)."""
    first, second, pose_init, pose2, _ = hdl_pair_small
    ctx.scan_clear()
    ctx.scan_upload(0, first)
    ctx.scan_upload(1, second)
    prm = pkg.default_params(0.5)
    pose_a, st = ctx.icp_pair(0, 1, pose_init, pose2, prm, 6)
    assert st.iterations_run == 6
    pose_b = pose_init.copy()
    for _ in range(6):
        pose_b, _ = ctx.icp_pair(0, 1, pose_b, pose2, prm, 1)
    assert np.array_equal(pose_a, pose_b)


def test_neq_out_equals_copy(pkg, ctx, hdl_pair_small):
    """m3dreg_icp_set_neq_out: the block the normal-equation kernel writes into the caller's buffer (what the N > 1 bench
    hands to NCCL) is the block icp_copy_neq returns, for every iteration; switching it off stops the writes."""
    import torch
    first, second, pose_init, pose2, _ = hdl_pair_small
    ctx.scan_clear()
    ctx.scan_upload(0, first)
    ctx.scan_upload(1, second)
    prm = pkg.default_params(0.5)
    ctx.icp_begin(0, 1, pose_init, pose2, prm)
    buf = torch.zeros(3 * 28, dtype=torch.float64, device="cuda")
    ref = torch.zeros(28, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()          # the context runs on its own stream
    try:
        for it in range(3):
            ctx.icp_set_neq_out(buf[28:])
            ctx.icp_step(1)
            ctx.icp_copy_neq(ref)
            ctx.synchronize()
            torch.cuda.synchronize()
            assert torch.equal(buf[28:56], ref) and float(ref[27]) > 100
            assert float(buf[:28].abs().sum()) == 0.0 and float(buf[56:].abs().sum()) == 0.0
        ctx.icp_set_neq_out(None)
        buf.zero_()
        ctx.icp_step(1)
        ctx.synchronize()
        torch.cuda.synchronize()
        assert float(buf.abs().sum()) == 0.0
    finally:
        ctx.icp_set_neq_out(None)
        ctx.icp_end()


def test_export_hooks_match_oracle(pkg, oracle, ctx, hdl_pair_small):
    first, second, pose_init, pose2, _ = hdl_pair_small
    ctx.scan_clear()
    ctx.scan_upload(0, first)
    ctx.scan_upload(1, second)
    prm = pkg.default_params(0.5)
    ctx.icp_pair(0, 1, pose_init, pose2, prm, 1)
    gp, table, buckets = ctx.export_last_grid(len(first))
    pose1 = oracle.euler_to_matrix(*oracle.matrix4_to_euler(pose_init))
    fg = oracle.transform_cloud(first, pose1)
    gp_o = oracle.grid_params(fg, 0.5)
    assert gp.tobytes() == gp_o.tobytes()
    buckets_o, table_o = oracle.build_grid(fg, gp_o)
    assert table.tobytes() == table_o.tobytes()
    assert np.array_equal(buckets["index_begin"], buckets_o["index_begin"])
    assert np.array_equal(buckets["number_of_points"], buckets_o["number_of_points"])


def test_icp_iteration_host_entry(pkg, oracle, ctx, hdl_pair_small):
    first, second, pose_init, pose2, _ = hdl_pair_small
    sg = _second_global(oracle, second, pose2)
    prm = pkg.default_params(0.5)
    nn = np.zeros(len(second), dtype=np.int32)
    pose_d, st = ctx.icp_iteration_host(first, sg, pose_init, prm, nn_out=nn)
    status, pose_o, n_obs, x_o, nn_o = oracle.icp_iteration(first, sg, pose_init, oracle.default_params(0.5), want_nn=True)
    assert st.last_status == 0 and status == 0
    assert np.array_equal(nn, nn_o)
    assert st.n_obs_last == n_obs
    assert np.allclose(np.array(st.x_last), x_o, rtol=1e-7, atol=1e-10)
    assert np.abs(pose_d - pose_o).max() < 2e-7


def test_too_few_observations_leaves_pose(pkg, ctx, synth):
    first = synth.random_cloud(500, seed=1, extent=(1, 1, 1))
    second = synth.random_cloud(500, seed=2, extent=(1, 1, 1))
    second["x"] += 50.0          # no overlap at all
    ctx.scan_clear()
    ctx.scan_upload(0, first)
    ctx.scan_upload(1, second)
    prm = pkg.default_params(0.5)
    pose0 = synth.pose_matrix(0.1, 0.2, 0.3, 0.01, 0.02, 0.03).astype(np.float32)
    pose, st = ctx.icp_pair(0, 1, pose0, np.eye(4, dtype=np.float32), prm, 2)
    assert st.last_status == pkg.E_TOO_FEW_OBS and st.n_obs_last == 0
    assert np.array_equal(pose, pose0)


def test_sweep_matches_oracle(pkg, oracle, ctx, synth):
    """registerAll Jacobi sweep: per-scan normal equations and new poses vs the oracle's sweep."""
    import torch
    scans, truth, init = synth.slam_scans(5, kind="hdl32", seed=9, spacing=1.0, n_azimuth=256)
    ctx.scan_clear()
    for k, s in enumerate(scans):
        ctx.scan_upload(k, s)
    for dof in (4, 6):
        prm = pkg.default_params(1.0, dof=dof)
        poses_o, neq_o, status_o = oracle.register_all_sweep(scans, init, oracle.default_params(1.0, dof=dof), pair_thr=10.0)
        pairs = [(i, j) for i in range(5) for j in range(5) if i != j]
        d_neq = torch.zeros(5 * 28, dtype=torch.float64, device="cuda")
        ctx.sweep_zero(d_neq, 5)
        ctx.sweep_accumulate([p[0] for p in pairs], [p[1] for p in pairs], init, prm, d_neq)
        poses_d, status_d = ctx.sweep_solve(d_neq, init, prm)
        neq_d = d_neq.cpu().numpy().reshape(5, 28)
        assert np.array_equal(neq_d[:, 27], neq_o[:, 27])                     # identical correspondence counts
        scale = np.abs(neq_o[:, :27]).max(axis=1, keepdims=True)
        assert (np.abs(neq_d[:, :27] - neq_o[:, :27]) <= 1e-10 * scale).all()
        assert np.array_equal(status_d, status_o)
        assert np.abs(poses_d - poses_o).max() < 1e-6
        # sharding invariance: two halves accumulated separately and added == one pass (what the all-reduce does)
        d_a = torch.zeros_like(d_neq)
        d_b = torch.zeros_like(d_neq)
        half = len(pairs) // 2
        ctx.sweep_accumulate([p[0] for p in pairs[:half]], [p[1] for p in pairs[:half]], init, prm, d_a)
        ctx.sweep_accumulate([p[0] for p in pairs[half:]], [p[1] for p in pairs[half:]], init, prm, d_b)
        s = (d_a + d_b).cpu().numpy().reshape(5, 28)
        assert (np.abs(s[:, :27] - neq_d[:, :27]) <= 1e-12 * scale).all() and np.array_equal(s[:, 27], neq_d[:, 27])


def test_slam_sweep_entry_matches_oracle(pkg, oracle, ctx, synth):
    """m3dreg_slam_sweep (gate + partition + accumulate + solve in ONE C call, registerAll of gpu6DSLAM.cpp:424-597) on one
    rank vs the oracle's sweep: identical counts, normal equations to round-off, poses and statuses; then
    number_of_last_EOZ semantics: only the last scans move, the others keep their pose bit for bit."""
    scans, truth, init = synth.slam_scans(5, kind="hdl32", seed=9, spacing=1.0, n_azimuth=256)
    ctx.scan_clear()
    for k, s in enumerate(scans):
        ctx.scan_upload(k, s)
    for dof in (4, 6):
        prm = pkg.default_params(1.0, dof=dof)
        poses_o, neq_o, status_o = oracle.register_all_sweep(scans, init, oracle.default_params(1.0, dof=dof), pair_thr=10.0)
        poses_d, status_d, st = ctx.slam_sweep(init, prm, 10.0, 0)
        neq_d = ctx.slam_neq(5)
        assert st.n_pairs == 20 and st.n_pairs_mine == 20 and st.world == 1
        assert np.array_equal(neq_d[:, 27], neq_o[:, 27])
        scale = np.abs(neq_o[:, :27]).max(axis=1, keepdims=True)
        assert (np.abs(neq_d[:, :27] - neq_o[:, :27]) <= 1e-10 * scale).all()
        assert np.array_equal(status_d, status_o)
        assert np.abs(poses_d - poses_o).max() < 1e-6
    # registerAll(..., number_of_last_EOZ = 2): scans 0..2 are neighbours only
    prm = pkg.default_params(1.0, dof=4)
    poses_l, status_l, st = ctx.slam_sweep(init, prm, 10.0, 3)
    assert st.n_pairs == 8
    assert np.array_equal(poses_l[:3], np.asarray(init, dtype=np.float32).reshape(-1, 4, 4)[:3])
    poses_full, _, _ = ctx.slam_sweep(init, prm, 10.0, 0)
    assert np.array_equal(poses_l[3:], poses_full[3:])       # Jacobi: a scan's update does not depend on who else is optimised
    # a tight gate leaves scans without pairs: they are still replaced by their Euler round trip, status TOO_FEW_OBS
    poses_t, status_t, st = ctx.slam_sweep(init, prm, 0.5, 0)
    assert st.n_pairs == 0 and (status_t == pkg.E_TOO_FEW_OBS).all()


def _nccl_rank(rank, world, port, q):
    import importlib, os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("mandala-mapping_b200")
    slam = importlib.import_module("mandala-mapping_b200.slam")
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)      # only carries the 128-byte NCCL id; the all-reduce is the library's
    scans, truth, init = pkg.synth.slam_scans(6, kind="hdl32", seed=9, spacing=1.0, n_azimuth=256)
    ctx = pkg.Context(rank)
    for k, s in enumerate(scans):
        ctx.scan_upload(k, s)
    prm = pkg.default_params(1.0, dof=4)
    drv = slam.DeviceSweep(ctx, prm, 10.0)
    poses, status = drv.sweep(init)
    neq = ctx.slam_neq(6)
    q.put((rank, poses, status, neq, int(drv.last_stats.n_pairs_mine)))
    ctx.close()
    dist.destroy_process_group()


def test_slam_sweep_two_ranks_nccl(pkg, ctx, synth):
    """Two ranks, two GPUs, the library's own NCCL all-reduce: both ranks end with the single-rank result."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    scans, truth, init = synth.slam_scans(6, kind="hdl32", seed=9, spacing=1.0, n_azimuth=256)
    ctx.scan_clear()
    for k, s in enumerate(scans):
        ctx.scan_upload(k, s)
    prm = pkg.default_params(1.0, dof=4)
    poses_1, status_1, _ = ctx.slam_sweep(init, prm, 10.0, 0)
    neq_1 = ctx.slam_neq(6)
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_nccl_rank, args=(r, 2, 29591, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
    assert res[0][4] > 0 and res[1][4] > 0 and res[0][4] + res[1][4] == 30
    for r in range(2):
        assert np.array_equal(res[r][2], status_1)
        assert np.array_equal(res[r][3][:, 27], neq_1[:, 27])
        scale = np.abs(neq_1[:, :27]).max(axis=1, keepdims=True)
        assert (np.abs(res[r][3][:, :27] - neq_1[:, :27]) <= 1e-12 * scale).all()
        assert np.abs(res[r][1] - poses_1).max() < 1e-6
    assert np.array_equal(res[0][1], res[1][1])          # both ranks hold the same poses bit for bit
