"""Generates tests/golden/preproc/*.npz on a B200 box by running the REFERENCE's own kernels (oracle/_ref/libm3dref.so)
behind the call sequences of CCudaWrapper::removeNoiseNaive / downsampling / classify / findBestYaw
(src/cudaWrapper.cpp:118-342, 662-836) on small seeded clouds:

    gpurun -- python tests/golden/make_golden_preproc.py     # writes gpurun_out/golden/preproc/*.npz; copy to tests/golden/preproc/

tests/test_oracle_preproc.py::test_golden_preproc_vectors replays them through the CPU oracle without a GPU."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import oracle
    from tests import refwrap
    synth = importlib.import_module("mandala-mapping_b200.synth")
    out_dir = os.path.join(ROOT, "gpurun_out", "golden", "preproc")
    os.makedirs(out_dir, exist_ok=True)
    raw = lambda c: np.frombuffer(c.tobytes(), dtype=np.uint8)
    cases = {"hdl_4k": synth.hdl32_scan(seed=201, n_azimuth=128), "sick_4k": synth.rotating_sick_scan(seed=202, n_beams=64, n_profiles=64)}
    q = synth.random_cloud(3000, seed=203, extent=(3, 3, 1))
    q["x"][0], q["y"][0], q["z"][0] = -9.0, -9.0, -3.0
    cases["quirk"] = q
    for name, cloud in cases.items():
        c = cloud.copy()
        c["normal_x"] = 0; c["normal_y"] = 0; c["normal_z"] = 0; c["label"] = 7
        d = dict(cloud=raw(c), noise_res=0.5, noise_ext=1.0, noise_threshold=3, down_res=0.3, down_ext=0.3,
                 cls_radius=1.0, cls_curvature=10.0, cls_ground_z=1.0, cls_plane_points=15, cls_ext=1.0, cls_max_inner=100, cls_max_outer=100,
                 cls_viewpoint=np.array([0.0, 0.0, 2.0], dtype=np.float32))
        d["noise_markers"] = refwrap.remove_noise_host(c, 0.5, 1.0, 3)
        d["down_markers"] = refwrap.downsample_host(c, 0.3, 0.3)
        out, mean, table = refwrap.classify_host(c, 1.0, 10.0, 1.0, 15, 1.0, 100, 100, (0.0, 0.0, 2.0))
        d["cls_cloud"], d["cls_mean"], d["cls_table"] = raw(out), mean, raw(table)
        if name == "hdl_4k":
            second = oracle.transform_cloud(synth.hdl32_scan(seed=204, n_azimuth=128), synth.pose_matrix(0, 0, 0, 0, 0, -np.deg2rad(4.5)).astype(np.float32))
            args = np.array([1.0, 1.0, 0.3, 50, 50, -9.0, 9.0, 1.5], dtype=np.float32)
            best, counts = refwrap.find_best_yaw_host(cloud, second, None, None, 1.0, 1.0, 0.3, 50, 50, -9.0, 9.0, 1.5)
            d["yaw_first"], d["yaw_second"], d["yaw_args"], d["yaw_counts"], d["yaw_best"] = raw(cloud), raw(second), args, counts, best
            print(name, "yaw best", best, counts.tolist())
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **d)
        print(name, "kept", int(d["noise_markers"].sum()), int(d["down_markers"].sum()), "labels", np.bincount(out["label"].clip(0, 7), minlength=8).tolist())


if __name__ == "__main__":
    main()
