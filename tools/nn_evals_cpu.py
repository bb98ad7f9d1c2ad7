#!/usr/bin/env python
"""CPU-side tuning aid: candidate evaluations per query of the device NN search logic (host instantiation of nn_core.cuh)
on the C2 pair, at the initial (perturbed) pose and along an oracle ICP run.  TEST TOOLING (uses oracle/)."""
import importlib, sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("mandala-mapping_b200")
import oracle
from tests import native

kind = sys.argv[1] if len(sys.argv) > 1 else "sick"
iters = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "3", "10"])]
bucket = float(sys.argv[3]) if len(sys.argv) > 3 else (1.0 if kind == "sick" else 0.5)
first, second, pose_init, pose2, _ = pkg.synth.scan_pair(kind, seed=42)
prm = oracle.default_params(bucket)
sg = oracle.transform_cloud(second, oracle.euler_to_matrix(*oracle.matrix4_to_euler(pose2)))
pose = pose_init.copy()
for it in range(max(iters) + 1):
    if it in iters:
        p1 = oracle.euler_to_matrix(*oracle.matrix4_to_euler(pose))
        fg = oracle.transform_cloud(first, p1)
        t0 = time.time()
        nn_o, gp, table, buckets = oracle.semantic_nn(fg, sg, bucket, bucket)
        t1 = time.time()
        nn_e, ev = native.nn_emul_search(fg, sg, table, buckets, gp, bucket)
        t2 = time.time()
        print(f"iter {it}: evals/query {ev / len(sg):8.1f}  match {np.array_equal(nn_e, nn_o)}  matched {(nn_o >= 0).mean():.3f}  oracle {t1 - t0:.1f}s emul {t2 - t1:.1f}s", flush=True)
    _, pose, _, _, _ = oracle.icp_iteration(first, sg, pose, prm, want_nn=True)
