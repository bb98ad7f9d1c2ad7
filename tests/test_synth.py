import numpy as np


def test_point_layout(synth):
    d = synth.POINT_DTYPE
    assert d.itemsize == 40
    assert [d.fields[n][1] for n in ("x", "y", "z", "intensity", "ring", "normal_x", "label", "rgb")] == [0, 4, 8, 12, 16, 20, 32, 36]
    assert synth.HASH_DTYPE.itemsize == 8 and synth.BUCKET_DTYPE.itemsize == 12
    assert synth.OBS_DTYPE.itemsize == 28 and synth.GRID_PARAMS_DTYPE.itemsize == 64
    assert synth.GRID_PARAMS_DTYPE.fields["number_of_buckets"][1] == 40


def test_hdl32_deterministic(synth):
    a = synth.hdl32_scan(seed=3, n_azimuth=128)
    b = synth.hdl32_scan(seed=3, n_azimuth=128)
    assert len(a) == 32 * 128
    assert a.tobytes() == b.tobytes()
    assert set(np.unique(a["label"])) <= {0, 1, 2, 3}
    assert np.isfinite(a["x"]).all() and np.isfinite(a["normal_x"]).all()


def test_normals_dot_le_one(synth):
    a = synth.hdl32_scan(seed=5, n_azimuth=256)
    n = np.stack([a["normal_x"], a["normal_y"], a["normal_z"]], axis=1)
    assert (synth._f32_dot_self(n) <= np.float32(1.0)).all()
    assert (synth._f32_dot_self(n) > np.float32(0.999)).all()


def test_sick_scan_shape(synth):
    a = synth.rotating_sick_scan(seed=1, n_beams=64, n_profiles=32)
    assert len(a) == 64 * 32
    # expressed in the unit frame: the room spans 40 x 30 x 6 m around a sensor at z = 2 m
    assert a["z"].min() > -2.2 and a["z"].max() < 4.2


def test_pair_and_trajectory(synth):
    f, s, p_init, p2, p_true = synth.scan_pair("hdl32", seed=2, n_azimuth=64)
    assert f.shape == s.shape and p_init.shape == (4, 4)
    assert np.allclose(p_true, p2)
    assert abs(p_init[0, 3] - p_true[0, 3] - 0.10) < 1e-6
    tr = synth.loop_trajectory(80, spacing=1.0)
    d = np.linalg.norm(tr[1:, :3, 3] - tr[:-1, :3, 3], axis=1)
    assert np.all(d < 1.0 + 1e-6) and np.all(d > 0.9)          # unit spacing along the loop (chords on arcs)
    assert np.abs(tr[:, 0, 3]).max() <= 12.5 + 1e-9 and np.abs(tr[:, 1, 3]).max() <= 7.5 + 1e-9


def test_slam_scans_by_worker_processes_match_the_serial_path(synth):
    """bench.py spreads the ray casts of a multi-scan set over child processes: same scans, same poses."""
    a, ta, ia = synth.slam_scans(5, kind="hdl32", seed=7, n_azimuth=64)
    b, tb, ib = synth.slam_scans(5, kind="hdl32", seed=7, n_azimuth=64, workers=3, only=[0, 2, 3])
    assert np.array_equal(ta, tb) and np.array_equal(ia, ib)
    assert b[1] is None and b[4] is None
    for k in (0, 2, 3):
        assert b[k].dtype == a[k].dtype and b[k].dtype.itemsize == 40
        assert b[k].tobytes() == a[k].tobytes()
