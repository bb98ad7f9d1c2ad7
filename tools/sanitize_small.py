#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): fused ICP on a 16k-point pair, a batched
sweep over 4 small scans and a stage-level NN call with spatially sorted queries.  Usage (GPU box):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("mandala-mapping_b200")
import torch

first, second, pose_init, pose2, _ = pkg.synth.scan_pair("hdl32", seed=7, n_azimuth=256)
prm = pkg.default_params(0.5)
ctx = pkg.Context(0)
ctx.scan_upload(0, first)
ctx.scan_upload(1, second)
pose, st = ctx.icp_pair(0, 1, pose_init, pose2, prm, 3)
nn = ctx.export_last_nn(len(second))
print("icp", st.iterations_run, st.n_obs_last, int((nn >= 0).sum()))
scans, truth, init = pkg.synth.slam_scans(4, kind="hdl32", seed=9, spacing=1.0, n_azimuth=128)
ctx.scan_clear()
for k, s in enumerate(scans):
    ctx.scan_upload(k, s)
pairs = [(i, j) for i in range(4) for j in range(4) if i != j]
d_neq = torch.zeros(4 * 28, dtype=torch.float64, device="cuda")
prm1 = pkg.default_params(1.0, dof=6)
ctx.sweep_zero(d_neq, 4)
ctx.sweep_accumulate([p[0] for p in pairs], [p[1] for p in pairs], init, prm1, d_neq)
poses, status = ctx.sweep_solve(d_neq, init, prm1)
print("sweep", status.tolist(), float(d_neq.cpu().numpy().reshape(4, 28)[:, 27].sum()))
f = pkg.synth.random_cloud(20000, seed=61, extent=(2, 2, 0.05))
q = pkg.synth.random_cloud(4000, seed=62, extent=(2, 2, 0.05))
order = np.lexsort((q["y"], q["x"], q["label"]))
for r, b in ((0.5, 0.5), (1.0, 0.4)):
    nn = ctx.semantic_nn_host(f, q[order].copy(), r, b)
    print("nn", r, b, int((nn >= 0).sum()))
ctx.close()
