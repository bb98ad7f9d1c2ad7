"""CPU tests of the oracle's pre-registration steps against independent numpy restatements
of the definitions in src/lesson_16.cu:740-1239, and against the fixtures the reference's own kernels produced on a
B200 (tests/golden/preproc/*.npz, written by tests/golden/make_golden_preproc.py)."""
import glob
import os

import numpy as np
import pytest

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "preproc", "*.npz")))


def _keys(oracle, cloud, res, ext):
    gp = oracle.grid_params(cloud, res, ext=ext)
    return gp, oracle.bucket_keys(cloud, gp)


def test_noise_filter_definition(oracle, synth):
    cloud = synth.random_cloud(20000, seed=3, extent=(6, 5, 2))
    cloud["x"][0], cloud["y"][0], cloud["z"][0] = -9.0, -9.0, -3.0
    for thr in (0, 3, 25):
        out, m = oracle.remove_noise(cloud, 0.5, 1.0, thr)
        gp, keys = _keys(oracle, cloud, 0.5, 1.0)
        buckets, table = oracle.build_grid(cloud, gp)
        want = buckets["number_of_points"][keys] > thr           # the dense table's count (0 for the quirk bucket)
        assert np.array_equal(m.astype(bool), want)
        assert len(out) == int(want.sum()) and np.array_equal(out["x"], cloud["x"][want])
    # the quirk: element 0 is alone in its bucket, the next occupied bucket keeps count 0 and loses its points even at threshold 0
    out, m = oracle.remove_noise(cloud, 0.5, 1.0, 0)
    counts = np.bincount(keys, minlength=int(gp["number_of_buckets"][0]))
    lost = (counts[keys] > 0) & (m == 0)
    assert lost.any() and len(np.unique(keys[lost])) == 1


def test_downsampling_definition(oracle, synth):
    cloud = synth.hdl32_scan(seed=4, n_azimuth=256)
    for res, ext in ((0.3, 0.3), (1.0, 1.0)):
        out, m = oracle.downsample(cloud, res, ext)
        gp, keys = _keys(oracle, cloud, res, ext)
        buckets, table = oracle.build_grid(cloud, gp)
        # one point per bucket that has an index_begin: the smallest original index of the bucket (stable sort)
        order = np.lexsort((np.arange(len(cloud)), keys))
        first = order[np.r_[True, keys[order][1:] != keys[order][:-1]]]
        has_begin = buckets["index_begin"][keys[first]] != -1
        want = np.zeros(len(cloud), dtype=bool)
        want[first[has_begin]] = True
        assert np.array_equal(m.astype(bool), want)
        assert len(out) == int(want.sum())
        assert int(want.sum()) == int((buckets["index_begin"] != -1).sum())


def _numpy_classify_point(cloud, table, buckets, gp, pos, radius, cap_in, cap_out):
    """mean (float32, visit order) and covariance of one sorted position, straight from the definition"""
    nbx, nby, nbz = (int(gp[f][0]) for f in ("nb_X", "nb_Y", "nb_Z"))
    key, idx = int(table["index_of_bucket"][pos]), int(table["index_of_point"][pos])
    ix, iy, iz = key // (nby * nbz), (key % (nby * nbz)) // nbz, key % nbz
    p = np.array([cloud["x"][idx], cloud["y"][idx], cloud["z"][idx]], dtype=np.float32)
    nbrs = []
    for i in range(0 if ix == 0 else -1, 1 if ix == nbx - 1 else 2):
        for j in range(0 if iy == 0 else -1, 1 if iy == nby - 1 else 2):
            for k in range(0 if iz == 0 else -1, 1 if iz == nbz - 1 else 2):
                nb = key + i * nby * nbz + j * nbz + k
                if nb < 0 or nb >= int(gp["number_of_buckets"][0]):
                    continue
                n = int(buckets["number_of_points"][nb])
                cap = cap_in if nb == key else cap_out
                if n <= 0 or cap <= 0:
                    continue
                it = 1 if cap >= n else max(n // cap, 1)
                for l in range(int(buckets["index_begin"][nb]), int(buckets["index_end"][nb]), it):
                    q = int(table["index_of_point"][l])
                    c = np.array([cloud["x"][q], cloud["y"][q], cloud["z"][q]], dtype=np.float32)
                    d = (p - c).astype(np.float32)
                    dist = np.sqrt(np.float32(np.float32(d[1] * d[1]) + np.float32(d[0] * d[0]) + np.float32(d[2] * d[2])))   # not fused: tolerance below
                    if dist <= np.float32(radius):
                        nbrs.append(c)
    return p, nbrs


def test_classification_definition(oracle, synth):
    cloud = synth.hdl32_scan(seed=5, n_azimuth=128)
    raw = cloud.copy()
    raw["normal_x"] = 0; raw["normal_y"] = 0; raw["normal_z"] = 0; raw["label"] = 7
    kw = dict(radius=1.0, curvature_threshold=10.0, ground_z=1.0, plane_points=15, ext=1.0, max_inner=100, max_outer=100, viewpoint=(0.0, 0.0, 2.0))
    out, mean, table = oracle.classify(raw, **kw)
    gp = oracle.grid_params(raw, 1.0, ext=1.0)
    buckets, table2 = oracle.build_grid(raw, gp)
    assert table.tobytes() == table2.tobytes()
    rng = np.random.default_rng(0)
    checked = planes = 0
    for pos in rng.choice(len(raw), 150, replace=False):
        p, nbrs = _numpy_classify_point(raw, table, buckets, gp, int(pos), 1.0, 100, 100)
        idx = int(table["index_of_point"][pos])
        if len(nbrs) < 3:
            assert not mean[pos].any()
            continue
        m = np.zeros(3, dtype=np.float32)
        for c in nbrs:
            m = (m + c).astype(np.float32)
        m = (m / np.float32(len(nbrs))).astype(np.float32)
        assert np.allclose(mean[pos], m, rtol=0, atol=2e-6), (pos, mean[pos], m)
        if len(nbrs) < 15:
            assert out["label"][idx] == 1 and out["normal_x"][idx] == 0
            continue
        d = np.stack(nbrs).astype(np.float64) - m.astype(np.float64)
        cov = d.T @ d / len(nbrs)
        w, v = np.linalg.eigh(cov)                               # ascending
        n = v[:, 0]
        got = np.array([out["normal_x"][idx], out["normal_y"][idx], out["normal_z"][idx]], dtype=np.float64)
        assert abs(abs(got @ n) - 1.0) < 1e-5, (pos, got, n)
        assert got @ (np.array([0.0, 0.0, 2.0]) - p.astype(np.float64)) >= 0          # flipped towards the viewpoint
        ratio = w[1] / w[0]
        if abs(ratio - 10.0) > 1e-3:
            is_plane = ratio > 10.0
            if is_plane:
                planes += 1
                want = (3 if p[2] < 1.0 else 2) if abs(got[2]) > 0.7 else 0
            else:
                want = 1
            assert out["label"][idx] == want, (pos, ratio, out["label"][idx], want)
        checked += 1
    assert checked > 30 and planes > 10


def test_yaw_sweep_recovers_rotation(oracle, synth):
    first, second, _, _, _ = synth.scan_pair("hdl32", seed=6, n_azimuth=128)
    rot = synth.pose_matrix(0.0, 0.0, 0.0, 0.0, 0.0, -np.deg2rad(6.0)).astype(np.float32)
    second_rot = oracle.transform_cloud(second, rot)
    best, n, counts = oracle.find_best_yaw(first, second_rot, None, None, bucket=1.0, ext=1.0, radius=0.3, max_inner=50, max_outer=50,
                                           angle_start=-12.0, angle_finish=12.0, angle_step=2.0)
    assert best == 6.0 and n == counts.max() and len(counts) == 13
    # the first maximum wins: a flat profile returns the start angle's index only if it is strictly exceeded nowhere
    assert int(np.argmax(counts)) == list(np.arange(-12.0, 12.1, 2.0)).index(6.0)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_golden_preproc_vectors(oracle, path):
    """Outputs of the REFERENCE's kernels on a B200 (make_golden_preproc.py) replayed through the CPU oracle."""
    g = np.load(path)
    cloud = np.frombuffer(g["cloud"].tobytes(), dtype=oracle.POINT_DTYPE).copy()
    _, m = oracle.remove_noise(cloud, float(g["noise_res"]), float(g["noise_ext"]), int(g["noise_threshold"]))
    assert np.array_equal(m, g["noise_markers"])
    _, m = oracle.downsample(cloud, float(g["down_res"]), float(g["down_ext"]))
    assert np.array_equal(m, g["down_markers"])
    kw = dict(radius=float(g["cls_radius"]), curvature_threshold=float(g["cls_curvature"]), ground_z=float(g["cls_ground_z"]),
              plane_points=int(g["cls_plane_points"]), ext=float(g["cls_ext"]), max_inner=int(g["cls_max_inner"]), max_outer=int(g["cls_max_outer"]),
              viewpoint=tuple(float(v) for v in g["cls_viewpoint"]))
    out, mean, table = oracle.classify(cloud, **kw)
    assert table.tobytes() == g["cls_table"].tobytes()
    assert mean.tobytes() == g["cls_mean"].tobytes()             # d_mean: bit-exact
    want = np.frombuffer(g["cls_cloud"].tobytes(), dtype=oracle.POINT_DTYPE)
    same = out["label"] == want["label"]
    assert same.mean() >= 0.999, same.mean()
    for f in ("normal_x", "normal_y", "normal_z"):
        assert np.abs(out[f][same] - want[f][same]).max() < 1e-4
    if "yaw_counts" in g.files:
        first = np.frombuffer(g["yaw_first"].tobytes(), dtype=oracle.POINT_DTYPE).copy()
        second = np.frombuffer(g["yaw_second"].tobytes(), dtype=oracle.POINT_DTYPE).copy()
        a = g["yaw_args"]
        best, n, counts = oracle.find_best_yaw(first, second, None, None, bucket=float(a[0]), ext=float(a[1]), radius=float(a[2]),
                                               max_inner=int(a[3]), max_outer=int(a[4]), angle_start=float(a[5]), angle_finish=float(a[6]), angle_step=float(a[7]))
        assert np.array_equal(counts, g["yaw_counts"]) and best == float(g["yaw_best"])
