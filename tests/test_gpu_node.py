"""ROS-free replay of the reference's per-scan driver (gpu6DSLAM::registerSingleScan,
src/gpu6DSLAM.cpp:4-222) through include/m3dreg_node.h, against
  * the composition of the individual C-ABI calls (cut-off, noise filter, downsampling, classification: bit-exact),
  * an oracle replay of the schedule (registerLastArrivedScan x 3 steps, registerAll over the last 3 scans x 3 steps) on the
    node's processed scans: poses within 1e-3 m / 1e-4 (the float pose chaining upstream is Eigen — unpinned — and a chained
    multi-scan run compounds last-ulp differences through re-matched correspondences; single loops are pinned to 1e-5 m in
    test_gpu_icp.py / test_gpu_parity_full.py),
  * the files it writes (PCD + three XML models) and loadmapfromfile / getMetascan / callbackInitialPose."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _params(pkg):
    p = pkg.node_default_params()
    p.cutoff_z_min = -3.0                      # the synthetic scanner sits 2 m above the floor of its own frame
    p.viewpoint[:] = [0.0, 0.0, 0.0]
    p.slam_registerLastArrivedScan_number_of_iterations_step[:] = [8, 8, 8]
    p.slam_registerAll_number_of_iterations_step[:] = [2, 2, 2]
    return p


def _raw(scan):
    r = scan.copy()
    r["normal_x"] = 0; r["normal_y"] = 0; r["normal_z"] = 0; r["label"] = 7
    return r


def _inv(m):
    return np.linalg.inv(m.astype(np.float64)).astype(np.float32)


def test_register_single_scan_replay(pkg, ctx, synth, oracle, tmp_path):
    scans, truth, _ = synth.slam_scans(4, kind="hdl32", seed=51, spacing=0.6, n_azimuth=256)
    # odometry: the truth with a translation and yaw error per scan — what the live 4-DOF solver (x, y, z, yaw) can correct
    rng = np.random.default_rng(52)
    init = np.stack([synth.pose_matrix(*(truth[k][:3, 3] + rng.normal(0, 0.04, 3)), 0.0, 0.0,
                                       np.arctan2(truth[k][1, 0], truth[k][0, 0]) + rng.normal(0, 0.01)).astype(np.float32) for k in range(4)])
    prm = _params(pkg)
    root = tmp_path / "map"
    node = pkg.Node(ctx, prm, str(root))
    stats = [node.register_single_scan(_raw(s), init[k], f"2016{k:04d}T000000") for k, s in enumerate(scans)]
    assert len(node) == 4

    # -- pre-registration: the node's processed scan == the composition of the individual calls ------------------------
    for k, s in enumerate(scans):
        r = _raw(s)
        keep = (r["z"] < 15) & (r["z"] > -3.0) & (r["x"] * r["x"] + r["y"] * r["y"] > np.float32(1.5))
        c = np.ascontiguousarray(r[keep])
        assert stats[k].n_raw == len(r) and stats[k].n_after_cutoff == len(c)
        c, _ = ctx.remove_noise(c, 0.5, 1.0, 3)
        assert stats[k].n_after_noise_removal == len(c)
        c, _ = ctx.downsample(c, 0.3, 0.3)
        assert stats[k].n_after_downsampling == len(c)
        c = ctx.classify(c, 1.0, 10.0, 1.0, 15, 1.0, 100, 100, (0.0, 0.0, 0.0))
        got = node.scan(k)
        assert all(np.array_equal(got[f].view(np.uint32 if got[f].dtype.itemsize == 4 else got[f].dtype),
                                  c[f].view(np.uint32 if c[f].dtype.itemsize == 4 else c[f].dtype)) for f in c.dtype.names), k
        assert (got["label"] != 1).sum() > 0.3 * len(got)              # surfaces were recognised: the semantic search has something to key on
    assert stats[0].pair_iterations == 0 and stats[0].sweeps == 0
    assert stats[1].pair_iterations == 24 and stats[1].sweeps == 0           # two scans: registerAll(.., 3) returns at once (gpu6DSLAM.cpp:428)
    assert stats[2].sweeps == 6 and stats[3].sweeps == 6 and stats[3].sweep_solved_last == 3

    # -- oracle replay of pose chaining + schedule on the processed scans ---------------------------------------------
    proc = [node.scan(k) for k in range(4)]
    vreg, last_mtf = [], None
    for k in range(4):
        mtf = init[k].astype(np.float32)
        if k == 0:
            vreg.append(mtf.copy())
        else:
            inc = (_inv(last_mtf) @ mtf).astype(np.float32)
            vreg.append((vreg[-1] @ inc).astype(np.float32))
            li = _inv(vreg[-1])
            vreg = [(mtf @ (li @ m)).astype(np.float32) for m in vreg]
            i, j = k, k - 1
            for step, rb in enumerate((2.5, 2.0, 1.0)):
                op = oracle.default_params(rb, dof=4)
                for _ in range(8):
                    sg = oracle.transform_cloud(proc[j], oracle.euler_to_matrix(*oracle.matrix4_to_euler(vreg[j])))
                    _, vreg[i], _, _, _ = oracle.icp_iteration(proc[i], sg, vreg[i], op)
            for step, rb in enumerate((2.5, 2.0, 1.0)):
                op = oracle.default_params(rb, dof=4)
                for _ in range(2 if k + 1 >= 3 else 0):
                    newp, _, _ = oracle.register_all_sweep(proc[:k + 1], np.stack(vreg), op, 10.0, first_optimised=k + 1 - 3)
                    vreg = [newp[q].copy() for q in range(k + 1)]
        last_mtf = mtf
    for k in range(4):
        reg, tf = node.pose(k)
        assert np.abs(reg[:3, 3] - vreg[k][:3, 3]).max() < 1e-3, (k, reg[:3, 3], vreg[k][:3, 3])
        assert np.abs(reg[:3, :3] - vreg[k][:3, :3]).max() < 1e-4, k
    # the registered trajectory stays near the truth (consecutive relative poses).  Point-to-point matching between scans from
    # DIFFERENT viewpoints is biased by the ring pattern on the floor (it pulls towards zero relative motion), so a short
    # schedule on 0.3 m-voxel scans is not expected to beat 4 cm odometry; convergence of the registration itself is tested
    # on same-viewpoint pairs (test_gpu_icp.py, test_gpu_parity_full.py) — here the point is parity with the oracle replay
    reg = np.stack([node.pose(k)[0] for k in range(4)])
    e_reg, e_odo = synth.relative_pose_error(reg, truth), synth.relative_pose_error(init, truth)
    print(f"relative pose error: odometry {e_odo:.4f} m -> registered {e_reg:.4f} m")
    assert e_reg < 0.25, (e_reg, e_odo)

    # -- files: raw + processed PCD per scan, three models per scan ----------------------------------------------------
    ids = [node.scan_id(k) for k in range(4)]
    assert ids[2] == "scan_20160002T000000"
    for k in range(4):
        t = f"2016{k:04d}T000000"
        raw_back = pkg.pcd_read(root / "rawData" / f"scan_{t}.pcd")
        assert len(raw_back) == len(scans[k]) and np.array_equal(raw_back["x"], scans[k]["x"])
        assert np.array_equal(pkg.pcd_read(root / "processedData" / f"scan_{t}.pcd")["label"], proc[k]["label"])
        for stem in ("tfModel_", "tfModelProcessedData_", "registeredData_"):
            assert (root / f"{stem}{t}.xml").exists()
    m = pkg.Model()
    assert m.load(root / "registeredData_20160003T000000.xml") and m.scan_ids() == ids
    for k in range(4):
        assert np.allclose(m.affine(ids[k]), node.pose(k)[0], rtol=2e-5, atol=2e-5)        # six significant digits on disk
    tfm = pkg.Model()
    assert tfm.load(root / "tfModel_20160003T000000.xml") and np.allclose(tfm.affine(ids[1]), init[1], rtol=2e-5, atol=2e-5)

    # -- getMetascan: every scan through its registered pose, concatenated ---------------------------------------------
    meta = node.metascan()
    assert len(meta) == sum(len(p) for p in proc)
    off = 0
    for k in range(4):
        want = oracle.transform_cloud(proc[k], node.pose(k)[0])
        seg = meta[off:off + len(want)]
        assert np.array_equal(seg["x"], want["x"]) and np.array_equal(seg["normal_z"], want["normal_z"]) and np.array_equal(seg["label"], want["label"])
        off += len(want)

    # -- loadmapfromfile into a fresh node, registerAll() service, callbackInitialPose ---------------------------------
    node2 = pkg.Node(ctx, prm, None)
    node2.load_map(str(root / "registeredData_20160003T000000.xml"))
    assert len(node2) == 4 and [node2.scan_id(k) for k in range(4)] == ids
    for k in range(4):
        assert np.array_equal(node2.scan(k)["x"], proc[k]["x"]) and np.allclose(node2.pose(k)[0], node.pose(k)[0], rtol=2e-5, atol=2e-5)
    before = np.stack([node2.pose(k)[0] for k in range(4)])
    assert node2.register_all() == 4
    after = np.stack([node2.pose(k)[0] for k in range(4)])
    op = oracle.default_params(0.5, dof=4)
    want, _, st = oracle.register_all_sweep(proc, before, op, 10.0)
    assert (st == 0).all() and np.abs(after[:, :3, 3] - want[:, :3, 3]).max() < 1e-5 and np.abs(after[:, :3, :3] - want[:, :3, :3]).max() < 1e-6
    anchor = synth.pose_matrix(5.0, -3.0, 0.5, 0.0, 0.0, 0.7).astype(np.float32)
    node2.set_initial_pose(anchor)
    moved = np.stack([node2.pose(k)[0] for k in range(4)])
    closest = int(np.argmin(np.linalg.norm(after[:, :3, 3] - anchor[:3, 3], axis=1)))
    assert np.allclose(moved[closest], anchor, atol=1e-5)                                   # pose * closest^-1 * initial
    for k in range(4):                                                                      # every pose: pose * closest^-1 * initial
        want_k = after[k].astype(np.float64) @ np.linalg.inv(after[closest].astype(np.float64)) @ anchor
        assert np.allclose(moved[k], want_k, atol=1e-4), k
    node2.close()
    node.close()
