"""Generates tests/golden/*.npz on a B200 box by running the REFERENCE's own kernels (oracle/_ref/libm3dref.so,
built from /root/reference by oracle/Makefile) on small seeded inputs:

    gpurun -- python tests/golden/make_golden.py      # writes gpurun_out/golden/*.npz; copy them to tests/golden/

Each file holds inputs (raw bytes of the 40-B points) and the reference's outputs: gridParameters, sorted
hashElement table, dense bucket table, NN indices, and for the registration part the observation vector built
from that NN result with AtPA/AtPl (cuBLAS) and the solution x (cuSOLVER potrf/potrs) for 6 and 4 DOF.
tests/test_oracle.py::test_golden_vectors replays them through the CPU oracle on a box without a GPU."""
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
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    cases = {
        "hdl_4k": (synth.hdl32_scan(seed=101, n_azimuth=128), synth.hdl32_scan(seed=102, n_azimuth=128), 0.5, 0.5, 1.0, 100, 100),
        "rand_stride": (synth.random_cloud(6000, seed=103, extent=(1.5, 1.5, 0.8)), synth.random_cloud(3000, seed=104, extent=(1.6, 1.6, 0.9)), 0.5, 0.5, 1.0, 5, 3),
        "sick_4k": (synth.rotating_sick_scan(seed=105, n_beams=64, n_profiles=64), synth.rotating_sick_scan(seed=106, n_beams=64, n_profiles=64), 1.0, 1.0, 1.0, 100, 100),
    }
    q = synth.random_cloud(2000, seed=107, extent=(2, 2, 1))
    q["x"][0], q["y"][0], q["z"][0] = -7.0, -7.0, -2.0
    cases["quirk"] = (q, synth.random_cloud(2000, seed=108, extent=(2, 2, 1)), 0.5, 0.5, 1.0, 100, 100)
    for name, (first, second, radius, bucket, ext, mi, mo) in cases.items():
        m = synth.pose_matrix(0.05, -0.03, 0.01, 0.004, -0.006, 0.01).astype(np.float32)
        fg = refwrap.transform_host(first, m)            # reference device transform kernel
        nn, gp, table, buckets = refwrap.nn_search_host(fg, second, radius, bucket, ext, mi, mo)
        d = dict(first=np.frombuffer(fg.tobytes(), dtype=np.uint8), second=np.frombuffer(second.tobytes(), dtype=np.uint8),
                 radius=radius, bucket=bucket, ext=ext, max_inner=mi, max_outer=mo,
                 grid_params=np.frombuffer(gp.tobytes(), dtype=np.uint8), table=np.frombuffer(table.tobytes(), dtype=np.uint8),
                 buckets=np.frombuffer(buckets.tobytes(), dtype=np.uint8), nn=nn)
        obs = oracle.build_observations(fg, first, second, nn)
        if len(obs) > 100:
            pose6 = np.array([0.05, -0.03, 0.01, 0.004, -0.006, 0.01])
            d["obs"] = np.frombuffer(obs.tobytes(), dtype=np.uint8)
            d["pose6"] = pose6
            for dof in (6, 4):
                N, b = refwrap.normal_equations_host(obs, pose6, dof)
                st, p_new, x = refwrap.register_ls_host(obs, pose6, dof)
                assert st == 0
                d[f"AtPA{dof}"], d[f"AtPl{dof}"], d[f"x{dof}"] = N, b, x
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **d)
        print(name, "matched", int((nn >= 0).sum()), "of", len(nn), "buckets", len(buckets))


if __name__ == "__main__":
    main()
