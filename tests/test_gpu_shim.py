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
