#!/usr/bin/env python
"""Work distribution of the LAST k_nn_search_hull launch of a C4-like sweep (see tools/nn_tail.py)."""
import importlib, os, sys
os.environ["M3DREG_NN_DIAG"] = "1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("mandala-mapping_b200")
slam = importlib.import_module("mandala-mapping_b200.slam")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
scans, truth, init = pkg.synth.slam_scans(n, kind="hdl32", seed=42, spacing=1.0)
prm = pkg.default_params(1.0, dof=4)
ctx = pkg.Context(0)
for k, s in enumerate(scans):
    ctx.scan_upload(k, s)
drv = slam.DeviceSweep(ctx, prm, 10.0)
drv.sweep(init)
ctx.set_profiling(True)
for rep in range(2):
    drv.sweep(init)
    ctx.synchronize()
    d = ctx.grid_phase_ns().astype(np.uint64)
    start = int(~d[0] & np.uint64(0xFFFFFFFFFFFFFFFF))
    span = (int(d[1]) - start) / 1e3
    warps, chunks = int(d[3]), int(d[6])
    ev = ctx.nn_evaluations(reset=True); fb = ctx.nn_fallbacks(reset=True)
    print(f"sweep launch: span {span:.1f} us, mean warp exit {int(d[2]) / max(warps, 1) / 1e3:.1f} us ({warps} warps), chunks {chunks}: mean {int(d[5]) / max(chunks, 1) / 1e3:.2f} us, "
          f"longest {int(d[4]) / 1e3:.1f} us, > 20 us: {int(d[7])}, > 50 us: {int(d[8])}; per-lane chunks: {int(d[10])}, mean {int(d[9]) / max(int(d[10]), 1) / 1e3:.1f} us, "
          f"share of chunk time {int(d[9]) / max(int(d[5]), 1):.2f}; sweep totals: evaluations {ev}, per-lane queries {fb}")
ctx.close()
