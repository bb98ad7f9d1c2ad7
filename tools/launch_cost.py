#!/usr/bin/env python
"""CPU launch cost vs GPU time of one fused iteration: python tools/launch_cost.py [workload]  (needs a GPU)"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("mandala-mapping_b200")
import bench
w = sys.argv[1] if len(sys.argv) > 1 else "c1"
first, second, pose_init, pose2, pose_true, res = bench.make_pair(pkg, w, 42)
prm = pkg.default_params(res)
ctx = pkg.Context(0)
ctx.scan_upload(0, first); ctx.scan_upload(1, second)
ctx.icp_begin(0, 1, pose_init, pose2, prm)
ctx.icp_step(10); ctx.synchronize()
n = 200
t0 = time.perf_counter(); ctx.icp_step(n); t1 = time.perf_counter(); ctx.synchronize(); t2 = time.perf_counter()
print(f"{w}: enqueue {1e6 * (t1 - t0) / n:.1f} us/iteration (CPU), complete {1e6 * (t2 - t0) / n:.1f} us/iteration, launches/iter {ctx.launch_count / (n + 10):.1f}")
ctx.icp_end(); ctx.close()
