"""CPU tests of the oracle (oracle/m3d_oracle.c) against independent numpy restatements and against the golden
vectors under tests/golden/ (produced on a B200 by the reference's own kernels, tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _f32(x):
    return np.asarray(x, dtype=np.float32)


def test_grid_params_and_keys_numpy(oracle, synth):
    c = synth.random_cloud(5000, seed=1)
    gp = oracle.grid_params(c, 0.5, ext=1.0)
    mn = _f32([c[a].min() for a in "xyz"]) - np.float32(1.0)
    mx = _f32([c[a].max() for a in "xyz"]) + np.float32(1.0)
    assert [gp["min_X"][0], gp["min_Y"][0], gp["min_Z"][0]] == list(mn)
    assert [gp["max_X"][0], gp["max_Y"][0], gp["max_Z"][0]] == list(mx)
    nb = (((mx - mn) / np.float32(0.5)) + np.float32(1.0)).astype(np.int32)
    assert [gp["nb_X"][0], gp["nb_Y"][0], gp["nb_Z"][0]] == list(nb)
    assert gp["number_of_buckets"][0] == int(nb[0]) * int(nb[1]) * int(nb[2])
    keys = oracle.bucket_keys(c, gp)
    ix = ((c["x"] - mn[0]) / np.float32(0.5)).astype(np.int32)
    iy = ((c["y"] - mn[1]) / np.float32(0.5)).astype(np.int32)
    iz = ((c["z"] - mn[2]) / np.float32(0.5)).astype(np.int32)
    assert np.array_equal(keys, ix * nb[1] * nb[2] + iy * nb[2] + iz)
    assert keys.min() >= 0 and keys.max() < gp["number_of_buckets"][0]


def test_build_grid_stable_and_table(oracle, synth):
    c = synth.random_cloud(20000, seed=2, extent=(3, 3, 1))
    gp = oracle.grid_params(c, 0.5)
    buckets, table = oracle.build_grid(c, gp)
    keys = oracle.bucket_keys(c, gp)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(table["index_of_point"], order)
    assert np.array_equal(table["index_of_bucket"], keys[order])
    occ = np.unique(keys)
    cnt = np.bincount(keys, minlength=len(buckets))
    first_alone = cnt[keys[order[0]]] == 1
    for b in occ:
        lo = np.searchsorted(keys[order], b, "left")
        hi = np.searchsorted(keys[order], b, "right")
        if first_alone and lo == 1:
            assert tuple(buckets[b]) == (-1, hi, 0)      # reference quirk (lesson_16.cu:154-158)
        else:
            assert tuple(buckets[b]) == (lo, hi, hi - lo)
    empty = np.setdiff1d(np.arange(len(buckets)), occ)
    assert (buckets["index_begin"][empty] == -1).all() and (buckets["number_of_points"][empty] == 0).all()


def test_first_element_quirk(oracle, synth):
    """Element 0 of the sorted table alone in its bucket => the second occupied bucket is invisible."""
    c = synth.random_cloud(64, seed=3, extent=(2, 2, 1))
    c["x"][0], c["y"][0], c["z"][0] = -5.0, -5.0, -1.0        # a lone point in the lowest cell
    gp = oracle.grid_params(c, 1.0)
    buckets, table = oracle.build_grid(c, gp)
    assert table["index_of_point"][0] == 0
    b1 = table["index_of_bucket"][1]
    assert buckets["index_begin"][b1] == -1 and buckets["number_of_points"][b1] == 0
    # queries sitting exactly on the points of that bucket find nothing there
    nn = oracle.nn_search(c, c, table, buckets, gp, 0.01)
    members = table["index_of_point"][table["index_of_bucket"] == b1]
    assert (nn[members] == -1).all()
    others = np.setdiff1d(np.arange(len(c)), members)
    assert (nn[others] == others).all()


def _brute_nn(first, second, gp, buckets, table, radius, max_in, max_out, gate):
    """Independent (slow) numpy restatement of the reference semantics, for small clouds."""
    mn = _f32([gp["min_X"][0], gp["min_Y"][0], gp["min_Z"][0]])
    mx = _f32([gp["max_X"][0], gp["max_Y"][0], gp["max_Z"][0]])
    res = _f32([gp["res_X"][0], gp["res_Y"][0], gp["res_Z"][0]])
    nb = [int(gp["nb_X"][0]), int(gp["nb_Y"][0]), int(gp["nb_Z"][0])]
    r2 = np.float32(radius) * np.float32(radius)
    out = np.full(len(second), -1, dtype=np.int32)
    for qi, q in enumerate(second):
        p = _f32([q["x"], q["y"], q["z"]])
        if (p < mn).any() or (p > mx).any():
            continue
        ic = ((p - mn) / res).astype(np.int32)
        best, best_i = np.float32(1e8), -1
        for i in range(-1 if ic[0] else 0, 1 if ic[0] == nb[0] - 1 else 2):
            for j in range(-1 if ic[1] else 0, 1 if ic[1] == nb[1] - 1 else 2):
                for k in range(-1 if ic[2] else 0, 1 if ic[2] == nb[2] - 1 else 2):
                    cell = (ic[0] + i) * nb[1] * nb[2] + (ic[1] + j) * nb[2] + ic[2] + k
                    n = buckets["number_of_points"][cell]
                    if n <= 0:
                        continue
                    cap = max_in if (i, j, k) == (0, 0, 0) else max_out
                    if cap <= 0:
                        continue
                    step = 1 if cap >= n else max(n // cap, 1)
                    for l in range(buckets["index_begin"][cell], buckets["index_end"][cell], step):
                        c = first[table["index_of_point"][l]]
                        dx, dy, dz = p[0] - c["x"], p[1] - c["y"], p[2] - c["z"]
                        t = np.float32(dy * dy)
                        t = np.float32(np.float64(dx) * np.float64(dx) + np.float64(t))
                        dist = np.float32(np.float64(dz) * np.float64(dz) + np.float64(t))
                        t = np.float32(q["normal_y"] * c["normal_y"])
                        t = np.float32(np.float64(q["normal_x"]) * np.float64(c["normal_x"]) + np.float64(t))
                        dot = np.float32(np.float64(q["normal_z"]) * np.float64(c["normal_z"]) + np.float64(t))
                        if c["label"] == q["label"] and gate(dot) and dist <= r2 and dist < best:
                            best, best_i = dist, table["index_of_point"][l]
        out[qi] = best_i
    return out


@pytest.mark.parametrize("max_in,max_out", [(100, 100), (3, 2), (1, 0)])
def test_nn_against_numpy(oracle, synth, max_in, max_out):
    first = synth.random_cloud(600, seed=4, extent=(1.5, 1.5, 0.8), n_labels=2)
    second = synth.random_cloud(200, seed=5, extent=(1.8, 1.8, 0.9), n_labels=2)
    gp = oracle.grid_params(first, 0.5, ext=0.2)
    buckets, table = oracle.build_grid(first, gp)
    nn = oracle.nn_search(first, second, table, buckets, gp, 0.5, max_in, max_out)
    ref = _brute_nn(first, second, gp, buckets, table, 0.5, max_in, max_out, lambda d: bool(oracle.lib().orc_angle_gate(float(d))))
    assert np.array_equal(nn, ref)
    assert (nn >= 0).sum() > 20


def test_angle_gate_properties(oracle):
    g = lambda d: bool(oracle.lib().orc_angle_gate(float(np.float32(d))))
    assert not g(0.0) and not g(-0.0)            # zero normals: exactly 90 deg -> rejected (SURVEY B-3)
    assert g(1.0) and g(0.5) and g(0.57) and g(1e-3)
    assert not g(-1e-3) and not g(-0.7) and not g(-1.0)
    assert not g(np.nextafter(np.float32(1.0), np.float32(2.0)))   # dot = 1.0000001 -> acosf NaN -> rejected
    assert not g(float("nan")) and not g(float("inf"))
    # monotone threshold near zero: exactly one switch from reject to accept
    xs = np.float32(2.0) ** np.arange(-40, -10, 0.25, dtype=np.float32)
    acc = np.array([g(x) for x in xs])
    assert acc[-1] and not acc[0] and (np.diff(acc.astype(int)) >= 0).all()


def test_normal_equations_against_numpy(oracle, synth):
    rng = np.random.default_rng(0)
    n = 500
    obs = np.zeros(n, dtype=synth.OBS_DTYPE)
    for f in obs.dtype.names:
        obs[f] = rng.normal(size=n).astype(np.float32)
    obs["P"] = rng.uniform(0.01, 1, n).astype(np.float32)
    pose6 = [0.3, -0.2, 0.1, 0.05, -0.1, 0.7]
    om, fi, ka = pose6[3:]
    so, co, sf, cf, sk, ck = np.sin(om), np.cos(om), np.sin(fi), np.cos(fi), np.sin(ka), np.cos(ka)
    R = np.array([[cf * ck, -cf * sk, sf],
                  [co * sk + so * sf * ck, co * ck - so * sf * sk, -so * cf],
                  [so * sk - co * sf * ck, so * ck + co * sf * sk, co * cf]])
    # analytic Jacobian check by finite differences of R(om,fi,ka) p0
    def Rm(a, b, c):
        return synth.pose_matrix(0, 0, 0, a, b, c)[:3, :3]
    assert np.allclose(R, Rm(om, fi, ka))
    A = np.zeros((3 * n, 6))
    l = np.zeros(3 * n)
    P = np.zeros(3 * n)
    eps = 1e-6
    for k in range(n):
        p0 = np.array([obs["x0"][k], obs["y0"][k], obs["z0"][k]], dtype=np.float64)
        J = np.stack([(Rm(om + eps, fi, ka) - Rm(om - eps, fi, ka)) @ p0, (Rm(om, fi + eps, ka) - Rm(om, fi - eps, ka)) @ p0,
                      (Rm(om, fi, ka + eps) - Rm(om, fi, ka - eps)) @ p0], axis=1) / (2 * eps)
        A[3 * k:3 * k + 3, :3] = -np.eye(3)
        A[3 * k:3 * k + 3, 3:] = -J
        l[3 * k:3 * k + 3] = [obs["x_diff"][k], obs["y_diff"][k], obs["z_diff"][k]]
        P[3 * k:3 * k + 3] = obs["P"][k]
    N_np = A.T @ (P[:, None] * A)
    b_np = A.T @ (P * l)
    N, b = oracle.normal_equations(obs, pose6, 6)
    assert np.allclose(N, N_np, rtol=1e-6, atol=1e-6) and np.allclose(b, b_np, rtol=1e-6, atol=1e-6)
    assert np.allclose(N, N.T)
    N4, b4 = oracle.normal_equations(obs, pose6, 4)
    sel = [0, 1, 2, 5]
    assert np.allclose(N4, N[np.ix_(sel, sel)], rtol=1e-13) and np.allclose(b4, b[sel], rtol=1e-13)
    info, x = oracle.chol_solve(N, b)
    assert info == 0 and np.allclose(x, np.linalg.solve(N, b), rtol=1e-9)
    st, p_new, x2 = oracle.register_ls(obs, pose6, 6)
    assert st == 0 and np.allclose(p_new, np.array(pose6) + x, rtol=1e-12)
    st, p4, x4 = oracle.register_ls(obs, pose6, 4)
    assert st == 0 and np.allclose(p4[[3, 4]], pose6[3:5]) and np.allclose(p4[5], pose6[5] + x4[3])


def test_chol_not_spd(oracle):
    N = np.eye(6)
    N[3, 3] = -1.0
    info, _ = oracle.chol_solve(N, np.ones(6))
    assert info == 4


def test_euler_helpers(oracle, synth):
    rng = np.random.default_rng(1)
    for _ in range(50):
        o = rng.uniform(-1.2, 1.2, 3).astype(np.float32)
        t = rng.uniform(-10, 10, 3).astype(np.float32)
        m = oracle.euler_to_matrix(o, t)
        assert np.allclose(m, synth.pose_matrix(t[0], t[1], t[2], o[0], o[1], o[2]), atol=3e-7)
        o2, t2 = oracle.matrix4_to_euler(m)
        assert np.allclose(o2, o, atol=2e-6) and np.array_equal(t2, t)


def test_transform_matches_matrix_product(oracle, synth):
    c = synth.random_cloud(1000, seed=6)
    m = synth.pose_matrix(1.0, -2.0, 0.5, 0.1, -0.2, 0.3).astype(np.float32)
    out = oracle.transform_cloud(c, m)
    xyz = np.stack([c["x"], c["y"], c["z"]], 1).astype(np.float64) @ m[:3, :3].astype(np.float64).T + m[:3, 3]
    assert np.allclose(np.stack([out["x"], out["y"], out["z"]], 1), xyz, atol=1e-5)
    assert np.array_equal(out["label"], c["label"]) and np.array_equal(out["ring"], c["ring"])


def test_icp_converges_to_truth(oracle, synth):
    f, s, p_init, p2, p_true = synth.scan_pair("hdl32", seed=11, n_azimuth=512)
    prm = oracle.default_params(0.5)
    sg = oracle.transform_cloud(s, oracle.euler_to_matrix(*oracle.matrix4_to_euler(p2)))
    pose = p_init.copy()
    for _ in range(25):
        st, pose, n_obs, _, _ = oracle.icp_iteration(f, sg, pose, prm)
        assert st == 0
    assert np.abs(pose[:3, 3] - p_true[:3, 3]).max() < 0.01
    o, _ = oracle.matrix4_to_euler(pose)
    assert np.abs(o[:2]).max() < 1e-3 and abs(o[2]) < 0.03     # yaw locks within ~2 azimuth steps (2*pi/512) of the ring pattern


def test_register_all_sweep_sane(oracle, synth):
    scans, truth, init = synth.slam_scans(4, kind="hdl32", seed=5, spacing=1.0, n_azimuth=256)
    prm = oracle.default_params(1.0, dof=4)
    poses = init.copy()
    e0 = np.abs(poses[:, :3, 3] - truth[:, :3, 3]).max()
    for _ in range(3):
        poses, neq, status = oracle.register_all_sweep(scans, poses, prm, pair_thr=10.0)
    assert (status == 0).all() and neq.shape == (4, 28) and (neq[:, 27] > 100).all()
    assert np.isfinite(poses).all()
    # point-to-point ICP between sparse ring scans from different viewpoints is biased towards pulling the
    # viewpoints together (floor rings), so only sanity is asserted: poses stay near the truth
    assert np.abs(poses[:, :3, 3] - truth[:, :3, 3]).max() < e0 + 0.25
    # packed systems are symmetric positive definite in their 4-DOF sub-block
    for k in range(4):
        N = np.zeros((6, 6))
        N[np.triu_indices(6)] = neq[k, :21]
        N = N + np.triu(N, 1).T
        sel = [0, 1, 2, 5]
        assert np.linalg.eigvalsh(N[np.ix_(sel, sel)]).min() > 0


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "*.npz"))))
def test_golden_vectors(oracle, path):
    """Golden vectors = outputs of the reference's own kernels on a B200 (see tests/golden/make_golden.py)."""
    g = np.load(path)
    first, second = g["first"].view(oracle.POINT_DTYPE).reshape(-1), g["second"].view(oracle.POINT_DTYPE).reshape(-1)
    radius, bucket, ext = float(g["radius"]), float(g["bucket"]), float(g["ext"])
    max_in, max_out = int(g["max_inner"]), int(g["max_outer"])
    nn, gp, table, buckets = oracle.semantic_nn(first, second, radius, bucket, ext, max_in, max_out)
    assert gp.tobytes() == g["grid_params"].tobytes()
    assert table.tobytes() == g["table"].tobytes()
    rb = g["buckets"].view(oracle.BUCKET_DTYPE).reshape(-1)
    # the quirk bucket's index_end is a write race upstream (1 or the run end): compare everything else
    quirk = (rb["index_begin"] == -1) & (rb["index_end"] != -1)
    assert quirk.sum() <= 1
    assert np.array_equal(rb["index_begin"], buckets["index_begin"])
    assert np.array_equal(rb["number_of_points"], buckets["number_of_points"])
    assert np.array_equal(rb["index_end"][~quirk], buckets["index_end"][~quirk])
    assert np.array_equal(nn, g["nn"])
    if "obs" in g:
        obs = g["obs"].view(oracle.OBS_DTYPE).reshape(-1)
        for dof in (6, 4):
            N, b = oracle.normal_equations(obs, g["pose6"], dof)
            assert np.allclose(N, g[f"AtPA{dof}"], rtol=1e-10, atol=1e-9)
            assert np.allclose(b, g[f"AtPl{dof}"], rtol=1e-10, atol=1e-9)
            st, p_new, x = oracle.register_ls(obs, g["pose6"], dof)
            assert st == 0 and np.allclose(x, g[f"x{dof}"], rtol=1e-7, atol=1e-10)


def test_ndt_definition_against_numpy(oracle, synth):
    """orc_ndt_normal_equations against a direct numpy evaluation of its definition."""
    first = synth.hdl32_scan(seed=71, n_azimuth=128)
    second = synth.hdl32_scan(seed=72, n_azimuth=128)
    pose = synth.pose_matrix(0.05, -0.02, 2.0, 0.004, -0.003, 0.01).astype(np.float32)
    fg = oracle.transform_cloud(first, pose)
    gp = oracle.grid_params(fg, 1.0)
    buckets, table = oracle.build_grid(fg, gp)
    pose6 = [0.05, -0.02, 2.0, 0.004, -0.003, 0.01]
    n_obs, neq = oracle.ndt_normal_equations(fg, first, second, table, buckets, gp, pose6)
    keys = oracle.bucket_keys(second, gp)
    mn = np.array([gp["min_X"][0], gp["min_Y"][0], gp["min_Z"][0]], dtype=np.float64)
    nbY, nbZ = int(gp["nb_Y"][0]), int(gp["nb_Z"][0])
    eps = 1e-6
    def Rm(a, b, c):
        return synth.pose_matrix(0, 0, 0, a, b, c)[:3, :3]
    om, fi, ka = pose6[3:]
    dR = [(Rm(om + eps, fi, ka) - Rm(om - eps, fi, ka)) / (2 * eps), (Rm(om, fi + eps, ka) - Rm(om, fi - eps, ka)) / (2 * eps),
          (Rm(om, fi, ka + eps) - Rm(om, fi, ka - eps)) / (2 * eps)]
    N = np.zeros((6, 6)); rhs = np.zeros(6); cnt = 0
    inside = ((second["x"] >= gp["min_X"][0]) & (second["x"] <= gp["max_X"][0]) & (second["y"] >= gp["min_Y"][0]) &
              (second["y"] <= gp["max_Y"][0]) & (second["z"] >= gp["min_Z"][0]) & (second["z"] <= gp["max_Z"][0]))
    cache = {}
    for qi in np.nonzero(inside)[0]:
        b = int(keys[qi])
        if buckets["number_of_points"][b] < 5:
            continue
        if b not in cache:
            idx = table["index_of_point"][buckets["index_begin"][b]:buckets["index_end"][b]]
            pg = np.stack([fg["x"][idx], fg["y"][idx], fg["z"][idx]], 1).astype(np.float64)
            pl = np.stack([first["x"][idx], first["y"][idx], first["z"][idx]], 1).astype(np.float64)
            S = np.cov(pg.T) + (0.05 * 1.0) ** 2 * np.eye(3)
            cache[b] = (pg.mean(0), pl.mean(0), np.linalg.inv(S))
        mu_g, mu_l, W = cache[b]
        q = np.array([second["x"][qi], second["y"][qi], second["z"][qi]], dtype=np.float64)
        A = -np.hstack([np.eye(3), np.stack([d @ mu_l for d in dR], 1)])
        N += A.T @ W @ A
        rhs += A.T @ W @ (mu_g - q)
        cnt += 1
    assert cnt == n_obs and neq[27] == cnt
    assert np.allclose(neq[:21], N[np.triu_indices(6)], rtol=1e-6, atol=1e-6 * np.abs(N).max())
    assert np.allclose(neq[21:27], rhs, rtol=1e-6, atol=1e-6 * np.abs(rhs).max())
    st, x = oracle.solve_packed(neq, 6)
    assert st == 0 and np.allclose(x, np.linalg.solve(N, rhs), rtol=1e-4, atol=1e-8)


def test_ndt_converges(oracle, synth):
    f, s, p_init, p2, p_true = synth.scan_pair("hdl32", seed=11, n_azimuth=512)
    prm = oracle.default_params(1.0, mode=1)
    sg = oracle.transform_cloud(s, oracle.euler_to_matrix(*oracle.matrix4_to_euler(p2)))
    pose = p_init.copy()
    for _ in range(8):
        st, pose, n_obs, _, _ = oracle.icp_iteration(f, sg, pose, prm)
        assert st == 0 and n_obs > 1000
    assert np.abs(pose[:3, 3] - p_true[:3, 3]).max() < 5e-3
    o, _ = oracle.matrix4_to_euler(pose)
    assert np.abs(o).max() < 2e-3


def _points_from_xyz(synth, xyz):
    pts = np.zeros(len(xyz), dtype=synth.POINT_DTYPE)
    pts["x"], pts["y"], pts["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    pts["normal_z"] = 1.0
    return pts


def test_ndt_oracle_matches_independent_numpy_fixture(oracle, synth):
    """NDT has no reference implementation; tests/golden/ndt/make_ndt_golden.py states its definition a third time with numpy
    / scipy library routines only (np.cov, np.linalg.inv, scipy Rotation, finite-difference Jacobian).  The oracle must
    reproduce that fixture, so that oracle and CUDA kernels are not only checked against each other."""
    g = np.load(os.path.join(HERE, "golden", "ndt", "ndt_planes.npz"))
    fg, fl, q = _points_from_xyz(synth, g["glob"]), _points_from_xyz(synth, g["local"]), _points_from_xyz(synth, g["queries"])
    gp = oracle.grid_params(fg, float(g["res"]), ext=float(g["ext"]))
    buckets, table = oracle.build_grid(fg, gp)
    _, neq = oracle.ndt_normal_equations(fg, fl, q, table, buckets, gp, g["pose6"])
    assert int(neq[27]) == int(g["n_obs"])
    scale = np.abs(g["neq28"][:27]).max()
    assert (np.abs(neq[:27] - g["neq28"][:27]) <= 1e-7 * scale).all()      # finite-difference Jacobian in the fixture: ~1e-9
    st, x = oracle.solve_packed(neq, 6)
    assert st == 0 and np.allclose(x, g["x"], rtol=1e-6, atol=1e-10)
