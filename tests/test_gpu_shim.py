"""C++ host (tests/csrc/shim_host.cpp) over include/cuda_wrapper_shim.hpp: the reference's CCudaWrapper call pattern
(src/gpu6DSLAM.cpp:264-422) and the fused device loop, both against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from tests import native

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dof", [6, 4])
def test_cpp_host_matches_oracle(pkg, oracle, hdl_pair_small, tmp_path, dof):
    first, second, pose_init, pose2, _ = hdl_pair_small
    exe = native.build_shim_host()
    first.tofile(tmp_path / "first.bin")
    second.tofile(tmp_path / "second.bin")
    np.concatenate([pose_init.reshape(-1), pose2.reshape(-1)]).astype(np.float32).tofile(tmp_path / "poses.bin")
    iters = 4
    out = tmp_path / "out.bin"
    subprocess.check_call([exe, str(tmp_path / "first.bin"), str(tmp_path / "second.bin"), str(tmp_path / "poses.bin"), "0.5",
                           str(iters), str(dof), str(out)], timeout=300)
    raw = np.fromfile(out, dtype=np.uint8)
    pose_legacy = raw[:64].view(np.float32).reshape(4, 4)
    pose_fused = raw[64:128].view(np.float32).reshape(4, 4)
    n2 = int(raw[128:132].view(np.int32)[0])
    nn_last = raw[132:132 + 4 * n2].view(np.int32)
    assert n2 == len(second)

    prm = oracle.default_params(0.5, dof=dof)
    sg = oracle.transform_cloud(second, oracle.euler_to_matrix(*oracle.matrix4_to_euler(pose2)))
    pose = pose_init.copy()
    nn_o = None
    for _ in range(iters):
        _, pose, _, _, nn_o = oracle.icp_iteration(first, sg, pose, prm, want_nn=True)
    # the legacy loop uses the very same arithmetic as the oracle: NN bit-exact, poses to fp64-summation round-off
    assert np.array_equal(nn_last, nn_o)
    assert np.abs(pose_legacy[:3, 3] - pose[:3, 3]).max() < 1e-5
    assert np.abs(pose_legacy[:3, :3] - pose[:3, :3]).max() < 1e-6
    assert np.abs(pose_fused[:3, 3] - pose[:3, 3]).max() < 1e-5
    assert np.abs(pose_fused[:3, :3] - pose[:3, :3]).max() < 1e-6


def test_cpp_preprocessing_through_the_shim(pkg, ctx, oracle, synth, tmp_path):
    """tests/csrc/shim_preproc.cpp: removeNoiseNaive -> downsampling -> classify and findBestYaw through the shim used as
    `class CCudaWrapper` (full swap) == the same steps through the C ABI, bit for bit."""
    exe = native.build_shim_program("shim_preproc")
    scan = synth.hdl32_scan(seed=61, n_azimuth=256)
    raw = scan.copy()
    raw["normal_x"] = 0; raw["normal_y"] = 0; raw["normal_z"] = 0; raw["label"] = 7
    other = oracle.transform_cloud(synth.hdl32_scan(seed=62, n_azimuth=256), synth.pose_matrix(0, 0, 0, 0, 0, -np.deg2rad(6.0)).astype(np.float32))
    raw.tofile(tmp_path / "scan.bin"); scan.tofile(tmp_path / "first.bin"); other.tofile(tmp_path / "other.bin")
    subprocess.check_call([exe, str(tmp_path / "scan.bin"), str(tmp_path / "first.bin"), str(tmp_path / "other.bin"), str(tmp_path / "proc.bin"),
                           str(tmp_path / "yaw.bin")], timeout=300)
    got = np.fromfile(tmp_path / "proc.bin", dtype=pkg.POINT_DTYPE)
    keep = (raw["z"] < 15) & (raw["z"] > -3) & (raw["x"] * raw["x"] + raw["y"] * raw["y"] > np.float32(1.5))
    c, _ = ctx.remove_noise(np.ascontiguousarray(raw[keep]), 0.5, 1.0, 3)
    c, _ = ctx.downsample(c, 0.3, 0.3)
    c = ctx.classify(c, 1.0, 10.0, 1.0, 15, 1.0, 100, 100, (0.0, 0.0, 0.0))
    assert len(got) == len(c)
    for f in c.dtype.names:      # field by field: bytes 18-19 of the 40-byte record are padding
        assert np.array_equal(got[f].view(np.uint32 if got[f].dtype.itemsize == 4 else got[f].dtype),
                              c[f].view(np.uint32 if c[f].dtype.itemsize == 4 else c[f].dtype)), f
    yaw = np.fromfile(tmp_path / "yaw.bin", dtype=np.float32)
    best, best_n, counts = ctx.find_best_yaw(scan, other, np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32), bucket=1.0, ext=1.0, radius=0.3,
                                             max_inner=50, max_outer=50, angle_start=-12.0, angle_finish=12.0, angle_step=1.5)
    assert yaw[0] == np.float32(best) == np.float32(6.0)
    want = oracle.euler_to_matrix(np.array([0.0, 0.0, np.float32(np.float64(best) * np.pi / 180.0)], dtype=np.float32), np.zeros(3, dtype=np.float32))
    assert np.array_equal(yaw[1:13].reshape(3, 4), want[:3, :])
