"""GPU parity of the pre-registration steps through the C ABI: product vs the reference's own
kernels (oracle/_ref, call sequences of cudaWrapper.cpp:118-342 and 662-836) vs the CPU oracle.
  * noise filter, downsampling: markers and surviving points bit-exact;
  * classification: sorted table and the float mean per sorted position (the reference's d_mean) bit-exact; normals and
    labels within the stated tolerance — the 3x3 decomposition is a Jacobi iteration here and a closed-form cubic
    upstream (third party, src/cuda_SVD.cu), so a label may differ only where lambda_mid / lambda_min sits on the
    threshold: label agreement >= 99.9 %, normals of agreeing plane points within 1e-4, of edge points within 1e-4 for
    99 % of them (the direction of a near-degenerate eigenvector is ill-conditioned for either solver);
  * yaw sweep: matches per angle (the semantic NN count) identical, same winning angle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# defaults of the reference node (include/gpu6DSLAM.h:163-176, 212-223)
NOISE = dict(res=0.5, ext=1.0, threshold=3)
DOWN = dict(res=0.3, ext=0.3)                      # gpu6DSLAM.cpp:71 passes the resolution as the box extension
CLS = dict(radius=1.0, curvature_threshold=10.0, ground_z=1.0, plane_points=15, ext=1.0, max_inner=100, max_outer=100,
           viewpoint=(0.0, 0.0, 2.0))


def _clouds(synth):
    yield "hdl16k", synth.hdl32_scan(seed=31, n_azimuth=512)
    yield "sick16k", synth.rotating_sick_scan(seed=32, n_beams=128, n_profiles=128)
    yield "rand", synth.random_cloud(30000, seed=33, extent=(6, 5, 2))
    q = synth.random_cloud(5000, seed=34, extent=(3, 3, 1))
    q["x"][0], q["y"][0], q["z"][0] = -9.0, -9.0, -3.0          # element 0 alone in the first bucket: the quirk bucket loses its points
    yield "quirk", q
    yield "tiny", synth.random_cloud(7, seed=35)


def _same_points(a, b):
    return all(np.array_equal(a[f].view(np.uint32 if a[f].dtype.itemsize == 4 else a[f].dtype),
                              b[f].view(np.uint32 if b[f].dtype.itemsize == 4 else b[f].dtype)) for f in a.dtype.names)


def test_noise_filter_and_downsampling_bit_exact(ctx, synth, oracle, ref):
    from tests import refwrap
    for name, cloud in _clouds(synth):
        for thr in (NOISE["threshold"], 0, 40):
            out, m = ctx.remove_noise(cloud, NOISE["res"], NOISE["ext"], thr)
            m_ref = refwrap.remove_noise_host(cloud, NOISE["res"], NOISE["ext"], thr)
            out_o, m_o = oracle.remove_noise(cloud, NOISE["res"], NOISE["ext"], thr)
            assert np.array_equal(m, m_ref), (name, thr)
            assert np.array_equal(m, m_o), (name, thr)
            assert len(out) == int(m.sum()) and _same_points(out, cloud[m != 0]), (name, thr)
        for res, ext in ((DOWN["res"], DOWN["ext"]), (1.0, 1.0)):
            out, m = ctx.downsample(cloud, res, ext)
            m_ref = refwrap.downsample_host(cloud, res, ext)
            out_o, m_o = oracle.downsample(cloud, res, ext)
            assert np.array_equal(m, m_ref), (name, res)
            assert np.array_equal(m, m_o), (name, res)
            assert len(out) == int(m.sum()) and _same_points(out, cloud[m != 0]), (name, res)


def _classify_agreement(a, b, what):
    """labels of two classified clouds + normals where both say plane-like (labels 0, 2, 3)"""
    same = a["label"] == b["label"]
    frac = float(same.mean())
    na = np.stack([a["normal_x"], a["normal_y"], a["normal_z"]], 1).astype(np.float64)
    nb = np.stack([b["normal_x"], b["normal_y"], b["normal_z"]], 1).astype(np.float64)
    both = same & (a["label"] != 1)
    dn = float(np.abs(na[both] - nb[both]).max()) if both.any() else 0.0
    # edge points carry a normal too (the plane test failed, the decomposition still ran).  Their two small eigenvalues may
    # be close, and the eigenvector error of ANY solver grows with 1 / gap (upstream's cubic works on the SQUARED spectrum of
    # A^T A and is the less accurate of the two): the bulk must agree tightly, the worst case only loosely
    edge = same & (a["label"] == 1)
    d = np.abs(na[edge] - nb[edge]).max(axis=1) if edge.any() else np.zeros(1)
    de = (float(np.quantile(d, 0.99)), float(d.max()))
    return frac, dn, de


def test_classification_vs_reference_and_oracle(ctx, synth, oracle, ref):
    from tests import refwrap
    for name, cloud in _clouds(synth):
        # a scan as it arrives: no normals, no labels
        raw = cloud.copy()
        raw["normal_x"] = 0; raw["normal_y"] = 0; raw["normal_z"] = 0; raw["label"] = 7
        for radius, caps in ((CLS["radius"], (100, 100)), (0.5, (20, 10))):
            kw = dict(CLS, radius=radius, max_inner=caps[0], max_outer=caps[1])
            got, mean, table = ctx.classify(raw, want_debug=True, **kw)
            want, mean_r, table_r = refwrap.classify_host(raw, **kw)
            orc, mean_o, table_o = oracle.classify(raw, **kw)
            assert table.tobytes() == table_r.tobytes() == table_o.tobytes(), (name, radius)
            assert mean.tobytes() == mean_r.tobytes(), (name, radius, "d_mean differs from the reference kernels")
            assert mean_o.tobytes() == mean_r.tobytes(), (name, radius, "oracle d_mean differs from the reference kernels")
            for other, tag in ((want, "reference"), (orc, "oracle")):
                frac, dn, de = _classify_agreement(got, other, tag)
                assert frac >= 0.999, (name, radius, tag, frac)
                assert dn < 1e-4 and de[0] < 1e-4 and de[1] < 2e-2, (name, radius, tag, dn, de)
            # untouched fields stay untouched
            for f in ("x", "y", "z", "intensity", "ring", "rgb"):
                assert np.array_equal(got[f], raw[f])
        labels = np.bincount(got["label"].clip(0, 7), minlength=8)
        if name in ("hdl16k", "sick16k"):
            assert labels[0] > 0 and labels[3] > 0, (name, labels)          # walls and floor are found in the room


def test_classification_recovers_analytic_labels(ctx, synth):
    """The synthetic room's analytic labels (what the registration tests use) against the classifier's."""
    cloud = synth.hdl32_scan(seed=36, n_azimuth=1024)
    raw = cloud.copy()
    raw["normal_x"] = 0; raw["normal_y"] = 0; raw["normal_z"] = 0; raw["label"] = 7
    got = ctx.classify(raw, **CLS)
    planar = cloud["label"] != 1
    classified_planar = got["label"] != 1
    # dense surfaces are recognised as planes; the dot product with the analytic normal is ~1 where both agree
    both = planar & classified_planar
    assert both.sum() > 0.5 * planar.sum(), (int(both.sum()), int(planar.sum()))
    dot = (got["normal_x"] * cloud["normal_x"] + got["normal_y"] * cloud["normal_y"] + got["normal_z"] * cloud["normal_z"])[both]
    assert np.median(np.abs(dot)) > 0.99


def test_yaw_sweep_counts_identical(ctx, synth, oracle, ref):
    from tests import refwrap
    first, second, pose_first, pose_second, _ = synth.scan_pair("hdl32", seed=37, n_azimuth=256)
    yaw = np.deg2rad(7.5)
    rot = synth.pose_matrix(0.0, 0.0, 0.0, 0.0, 0.0, -yaw).astype(np.float32)
    second_rot = oracle.transform_cloud(second, rot)                     # the sweep has to turn it back by +7.5 degrees
    args = dict(bucket=1.0, ext=1.0, radius=0.3, max_inner=50, max_outer=50, angle_start=-12.0, angle_finish=12.0, angle_step=1.5)
    ident = np.eye(4, dtype=np.float32)
    for m2, m1inv in ((None, None), (ident, ident), (synth.pose_matrix(0.02, -0.01, 0.0, 0.0, 0.0, 0.01).astype(np.float32), None)):
        best, best_n, counts = ctx.find_best_yaw(first, second_rot, m2, m1inv, **args)
        best_r, counts_r = refwrap.find_best_yaw_host(first, second_rot, m2, m1inv, args["bucket"], args["ext"], args["radius"],
                                                      args["max_inner"], args["max_outer"], args["angle_start"], args["angle_finish"], args["angle_step"])
        best_o, best_n_o, counts_o = oracle.find_best_yaw(first, second_rot, m2, m1inv, **args)
        assert np.array_equal(counts, counts_r), (counts, counts_r)
        assert np.array_equal(counts, counts_o)
        assert best == best_r == best_o and best_n == int(counts.max())
        assert abs(best - 7.5) <= 1.5


def test_preproc_edge_cases_and_argument_errors(pkg, ctx, synth, oracle):
    """One point, coincident points, a cloud in a single bucket, zero survivors; invalid arguments report, never crash."""
    import ctypes as C
    one = synth.random_cloud(1, seed=70)
    out, m = ctx.remove_noise(one, 0.5, 1.0, 0)
    assert m.tolist() == oracle.remove_noise(one, 0.5, 1.0, 0)[1].tolist() and len(out) == int(m.sum())
    out, m = ctx.downsample(one, 0.3, 0.3)
    assert m.tolist() == [1] and len(out) == 1
    got = ctx.classify(one, **CLS)
    assert got["label"][0] == 1 and got["normal_x"][0] == 0 and got["normal_z"][0] == 0            # fewer than three neighbours: edge, no normal
    same = synth.random_cloud(300, seed=71)
    same["x"], same["y"], same["z"] = 1.25, -0.5, 0.75                                             # 300 coincident points
    out, m = ctx.remove_noise(same, 0.5, 1.0, 3)
    assert m.all() and len(out) == 300
    out, m = ctx.downsample(same, 0.3, 0.3)
    assert m.sum() == 1 and m[0] == 1                                                               # the smallest original index of the bucket stays
    got = ctx.classify(same, **CLS)
    want, _, _ = oracle.classify(same, **CLS)
    assert np.array_equal(got["label"], want["label"])                                             # zero covariance: edge everywhere
    sparse = synth.random_cloud(500, seed=72, extent=(40, 40, 10))
    out, m = ctx.remove_noise(sparse, 0.5, 1.0, 5)
    assert np.array_equal(m, oracle.remove_noise(sparse, 0.5, 1.0, 5)[1]) and len(out) == int(m.sum())
    out, m = ctx.remove_noise(sparse, 0.5, 1.0, 1000)
    assert len(out) == 0 and not m.any()                                                            # nothing survives
    # argument errors through the raw C ABI
    L = pkg.lib()
    kept = C.c_int(0)
    buf = np.zeros(4, dtype=pkg.POINT_DTYPE)
    assert L.m3dreg_remove_noise_host(ctx._h, None, C.c_int(4), C.c_float(0.5), C.c_float(1.0), C.c_int(3), pkg._p(buf), C.byref(kept), None) == pkg.E_INVALID_ARG
    assert L.m3dreg_remove_noise_host(ctx._h, pkg._p(buf), C.c_int(0), C.c_float(0.5), C.c_float(1.0), C.c_int(3), pkg._p(buf), C.byref(kept), None) == pkg.E_INVALID_ARG
    assert L.m3dreg_downsample_host(ctx._h, pkg._p(buf), C.c_int(4), C.c_float(0.0), C.c_float(1.0), pkg._p(buf), C.byref(kept), None) == pkg.E_INVALID_ARG
    assert L.m3dreg_classify_host(ctx._h, pkg._p(buf), C.c_int(4), C.c_float(-1.0), C.c_float(10.0), C.c_float(1.0), C.c_int(15), C.c_float(1.0),
                                  C.c_int(100), C.c_int(100), C.c_float(0), C.c_float(0), C.c_float(0), None, None) == pkg.E_INVALID_ARG
    best = C.c_float(0)
    assert L.m3dreg_find_best_yaw_host(ctx._h, pkg._p(buf), C.c_int(4), pkg._p(buf), C.c_int(4), None, None, C.c_float(1.0), C.c_float(1.0), C.c_float(0.3),
                                       C.c_int(50), C.c_int(50), C.c_float(-3.0), C.c_float(3.0), C.c_float(0.0), C.byref(best), None, None, C.c_int(0)) == pkg.E_INVALID_ARG
    # a grid that does not fit int32 is refused (upstream overflows silently, lesson_16.cu:80)
    huge = synth.random_cloud(64, seed=73, extent=(3000, 3000, 3000))
    with pytest.raises(pkg.M3dRegError) as ei:
        ctx.remove_noise(huge, 0.001, 1.0, 0)
    assert ei.value.status == pkg.E_TOO_MANY_BUCKETS
    # the context is still usable afterwards
    out, m = ctx.downsample(sparse, 1.0, 1.0)
    assert np.array_equal(m, oracle.downsample(sparse, 1.0, 1.0)[1])
