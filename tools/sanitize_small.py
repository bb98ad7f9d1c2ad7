#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): fused ICP on a 16k-point pair, a batched
sweep over 4 small scans, a stage-level NN call with spatially sorted queries, m3dreg_slam_sweep, NDT, the pre-registration
steps (noise filter, downsampling, classification, yaw sweep) and the node replay of two scans.  Usage (GPU box):
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
# round 2: the sweep as one call, NDT mode, pre-registration kernels, node replay
slam = importlib.import_module("mandala-mapping_b200.slam")
drv = slam.DeviceSweep(ctx, prm1, 10.0)
poses2, st2 = drv.sweep(init)
print("slam_sweep", st2.tolist())
prm_ndt = pkg.default_params(1.0, dof=6, mode=pkg.MODE_NDT)
pose_n, st_n = ctx.icp_pair(0, 1, init[0], init[1], prm_ndt, 2)
print("ndt", st_n.iterations_run, st_n.n_obs_last)
raw = scans[0].copy()
raw["normal_x"] = 0; raw["normal_y"] = 0; raw["normal_z"] = 0; raw["label"] = 7
c, m = ctx.remove_noise(raw, 0.5, 1.0, 3)
c, m = ctx.downsample(c, 0.3, 0.3)
c = ctx.classify(c, 1.0, 10.0, 1.0, 15, 1.0, 100, 100, (0.0, 0.0, 0.0))
print("preproc", len(raw), len(c), np.bincount(c["label"].clip(0, 7), minlength=4)[:4].tolist())
best, n_best, counts = ctx.find_best_yaw(scans[0], scans[1], None, None, bucket=1.0, ext=1.0, radius=0.3, max_inner=50, max_outer=50,
                                         angle_start=-3.0, angle_finish=3.0, angle_step=1.5)
print("yaw", best, n_best, counts.tolist())
np_ = pkg.node_default_params()
np_.cutoff_z_min = -3.0
np_.viewpoint[:] = [0.0, 0.0, 0.0]
np_.slam_registerLastArrivedScan_number_of_iterations_step[:] = [2, 2, 2]
np_.slam_registerAll_number_of_iterations_step[:] = [1, 1, 1]
node = pkg.Node(ctx, np_, None)
for k in range(3):
    r = scans[k].copy()
    r["normal_x"] = 0; r["normal_y"] = 0; r["normal_z"] = 0; r["label"] = 7
    stn = node.register_single_scan(r, init[k], f"t{k}")
print("node", len(node), stn.pair_iterations, stn.sweeps, len(node.metascan()))
node.close()
ctx.close()
