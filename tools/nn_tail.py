#!/usr/bin/env python
"""Work distribution of k_nn_search_hull inside the stream (profiling build of the kernel, %globaltimer stamps):
when do the persistent warps run dry, how long are the chunks.  usage: python tools/nn_tail.py [c1|c2|c3] [iterations]"""
import importlib, os, sys
os.environ["M3DREG_NN_DIAG"] = "1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module("mandala-mapping_b200")
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
first, second, pose_init, pose2, pose_true, res = bench.make_pair(pkg, wl, 42)
prm = pkg.default_params(res)
ctx = pkg.Context(0)
ctx.scan_upload(0, first); ctx.scan_upload(1, second)
ctx.icp_begin(0, 1, pose_init, pose2, prm)
ctx.icp_step(iters)
ctx.set_profiling(True)
for k in range(3):
    ctx.icp_step(1)
    ctx.synchronize()
    d = ctx.grid_phase_ns().astype(np.uint64)
    start = int(~d[0] & np.uint64(0xFFFFFFFFFFFFFFFF))
    span = (int(d[1]) - start) / 1e3
    warps, chunks = int(d[3]), int(d[6])
    print(f"{wl}: span {span:.1f} us, mean warp exit {int(d[2]) / max(warps, 1) / 1e3:.1f} us ({warps} warps), chunks {chunks}: mean {int(d[5]) / max(chunks, 1) / 1e3:.2f} us, "
          f"longest {int(d[4]) / 1e3:.1f} us, > 20 us: {int(d[7])}, > 50 us: {int(d[8])}; chunks that searched per lane: {int(d[10])}, mean {int(d[9]) / max(int(d[10]), 1) / 1e3:.1f} us")
ctx.set_profiling(False)
ctx.icp_end()
ctx.close()
