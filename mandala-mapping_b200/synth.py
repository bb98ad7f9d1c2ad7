"""Deterministic synthetic scans for the registration hot path.

The reference ships no sample data (``*.pcd`` is git-ignored upstream), so every test and benchmark
input is generated here: a 40 x 30 x 6 m box room with 8 axis-aligned boxes and 4 pillars, ray-cast
from either

* an HDL-32E-like spinning lidar (32 rings, 2048 azimuth steps -> 65 536 points), or
* the m3d rotating-SICK unit: a 270 deg / N-beam planar profile
  (``m3d/m3d_aggregator/src/m3d_aggregator.cpp:269-285``) whose laser frame is ``RPY(0, -pi/2, ang)`` at
  (-0.0835, 0, 0.1835) m from the unit frame (``m3d/m3dunit_base/src/encoder_node_li.cpp:90-98``),
  accumulated while ``ang`` sweeps 1.1*pi (``m3d_aggregator.cpp:30``).

Points are emitted directly as the reference's 40-byte ``PointXYZIRNLRGB`` records
(``gpu_6dslam/gpu_6dslam/include/custom_point_types.h:8-20``) in the SENSOR (local) frame with analytic unit
normals facing the sensor and analytic semantic labels (lesson_16.h:9-12), so the registration path can be
exercised without the classification stage.  Everything is numpy + PCG64 with fixed seeds.
"""
from __future__ import annotations

import math

import numpy as np

LABEL_PLANE, LABEL_EDGE, LABEL_CEILING, LABEL_GROUND = 0, 1, 2, 3

#: numpy view of lidar_pointcloud::PointXYZIRNLRGB (40 bytes, align 4)
POINT_DTYPE = np.dtype(
    {
        "names": ["x", "y", "z", "intensity", "ring", "normal_x", "normal_y", "normal_z", "label", "rgb"],
        "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4", "<f4", "<f4", "<i4", "<f4"],
        "offsets": [0, 4, 8, 12, 16, 20, 24, 28, 32, 36],
        "itemsize": 40,
    }
)
HASH_DTYPE = np.dtype([("index_of_point", "<i4"), ("index_of_bucket", "<i4")])
BUCKET_DTYPE = np.dtype([("index_begin", "<i4"), ("index_end", "<i4"), ("number_of_points", "<i4")])
OBS_DTYPE = np.dtype([(n, "<f4") for n in ("x_diff", "y_diff", "z_diff", "x0", "y0", "z0", "P")])
GRID_PARAMS_DTYPE = np.dtype(
    {
        "names": ["min_X", "min_Y", "min_Z", "max_X", "max_Y", "max_Z", "nb_X", "nb_Y", "nb_Z",
                  "number_of_buckets", "res_X", "res_Y", "res_Z"],
        "formats": ["<f4"] * 6 + ["<i4"] * 3 + ["<i8"] + ["<f4"] * 3,
        "offsets": [0, 4, 8, 12, 16, 20, 24, 28, 32, 40, 48, 52, 56],
        "itemsize": 64,
    }
)

ROOM_LO = np.array([-20.0, -15.0, 0.0])
ROOM_HI = np.array([20.0, 15.0, 6.0])

# 8 axis-aligned boxes (lo, hi) and 4 full-height pillars (cx, cy, r): all at least 1.5 m away from the
# 25 x 15 m rounded-rectangle trajectory so the aggregator's +-1 m self-exclusion box never triggers.
BOXES = np.array(
    [
        [[-8.0, -3.0, 0.0], [-5.0, -1.0, 2.5]],
        [[-2.0, 2.0, 0.0], [1.0, 5.0, 1.5]],
        [[4.0, -4.0, 0.0], [7.0, -2.0, 3.0]],
        [[5.0, 2.0, 0.0], [8.5, 4.5, 2.0]],
        [[-18.5, -13.5, 0.0], [-15.5, -10.5, 4.0]],
        [[15.0, 10.0, 0.0], [18.5, 13.5, 2.2]],
        [[-18.0, 10.5, 0.0], [-15.0, 13.0, 1.2]],
        [[15.5, -13.0, 0.0], [18.0, -10.5, 5.0]],
    ]
)
PILLARS = np.array([[-4.0, 3.0, 0.6], [2.0, -3.5, 0.5], [-16.5, 0.0, 0.8], [16.5, 0.5, 0.7]])

_EDGE_BAND = 0.05


def _raycast(origin: np.ndarray, dirs: np.ndarray):
    """First hit of rays ``origin + t*dirs`` (world frame, float64) with the scene.

    Returns (t, normal[N,3], label[N])."""
    n = dirs.shape[0]
    o = np.broadcast_to(origin, (n, 3))
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / dirs
    # room (always hit from inside)
    bound = np.where(dirs > 0, ROOM_HI, ROOM_LO)
    tr = (bound - o) * inv
    tr = np.where(np.isfinite(tr) & (dirs != 0), tr, np.inf)
    axis = np.argmin(tr, axis=1)
    t = tr[np.arange(n), axis]
    normal = np.zeros((n, 3))
    normal[np.arange(n), axis] = -np.sign(dirs[np.arange(n), axis])
    label = np.full(n, LABEL_PLANE, dtype=np.int32)
    label[(axis == 2) & (dirs[:, 2] < 0)] = LABEL_GROUND
    label[(axis == 2) & (dirs[:, 2] > 0)] = LABEL_CEILING

    for lo, hi in BOXES:
        t1 = (lo - o) * inv
        t2 = (hi - o) * inv
        tlo = np.minimum(t1, t2)
        thi = np.maximum(t1, t2)
        par = dirs == 0
        inside = (o >= lo) & (o <= hi)
        tlo = np.where(par, np.where(inside, -np.inf, np.inf), tlo)
        thi = np.where(par, np.where(inside, np.inf, -np.inf), thi)
        tn = tlo.max(axis=1)
        tf = thi.min(axis=1)
        hit = (tf >= tn) & (tn > 1e-9) & (tn < t)
        if not hit.any():
            continue
        ax = np.argmax(tlo, axis=1)
        idx = np.nonzero(hit)[0]
        t[idx] = tn[idx]
        normal[idx] = 0.0
        normal[idx, ax[idx]] = -np.sign(dirs[idx, ax[idx]])
        p = o[idx] + tn[idx, None] * dirs[idx]
        d_edge = np.minimum(np.abs(p - lo), np.abs(p - hi))
        d_edge[np.arange(idx.size), ax[idx]] = np.inf
        # the floor contact line is not a box edge
        lab = np.where(d_edge[:, :2].min(axis=1) < _EDGE_BAND, LABEL_EDGE, LABEL_PLANE)
        top_edge = (np.abs(p[:, 2] - hi[2]) < _EDGE_BAND) & (ax[idx] != 2)
        lab = np.where(top_edge, LABEL_EDGE, lab)
        label[idx] = lab

    for cx, cy, r in PILLARS:
        ox = o[:, 0] - cx
        oy = o[:, 1] - cy
        a = dirs[:, 0] ** 2 + dirs[:, 1] ** 2
        b = ox * dirs[:, 0] + oy * dirs[:, 1]
        c = ox * ox + oy * oy - r * r
        disc = b * b - a * c
        with np.errstate(divide="ignore", invalid="ignore"):
            tc = (-b - np.sqrt(np.where(disc > 0, disc, np.nan))) / a
        z = o[:, 2] + tc * dirs[:, 2]
        hit = np.isfinite(tc) & (tc > 1e-9) & (tc < t) & (z >= ROOM_LO[2]) & (z <= ROOM_HI[2])
        if not hit.any():
            continue
        idx = np.nonzero(hit)[0]
        t[idx] = tc[idx]
        nx = (ox[idx] + tc[idx] * dirs[idx, 0]) / r
        ny = (oy[idx] + tc[idx] * dirs[idx, 1]) / r
        normal[idx, 0] = nx
        normal[idx, 1] = ny
        normal[idx, 2] = 0.0
        label[idx] = LABEL_PLANE
    return t, normal, label


def _f32_dot_self(n: np.ndarray) -> np.ndarray:
    """n.n evaluated like the NN kernel evaluates dot products: fma(nz,nz, fma(nx,nx, ny*ny)) in f32."""
    nx, ny, nz = (n[:, i].astype(np.float32) for i in range(3))
    t = (ny * ny).astype(np.float32)
    t = (nx.astype(np.float64) * nx.astype(np.float64) + t.astype(np.float64)).astype(np.float32)
    return (nz.astype(np.float64) * nz.astype(np.float64) + t.astype(np.float64)).astype(np.float32)


def _shrink_normals(nrm: np.ndarray) -> np.ndarray:
    """Round normals to f32 such that their f32 self-dot is <= 1 (a dot that
    rounds to 1.0000001 makes acosf return NaN and the reference rejects the match)."""
    out = nrm.astype(np.float32)
    for _ in range(8):
        bad = _f32_dot_self(out) > np.float32(1.0)
        if not bad.any():
            break
        out[bad] = (out[bad].astype(np.float64) * (1.0 - 2.0 ** -23)).astype(np.float32)
    return out


def pose_matrix(tx, ty, tz, om, fi, ka) -> np.ndarray:
    """4x4 float64 matrix T(t) * Rx(om) * Ry(fi) * Rz(ka) (lesson_16.cu:278-293 convention)."""
    co, so, cf, sf, ck, sk = math.cos(om), math.sin(om), math.cos(fi), math.sin(fi), math.cos(ka), math.sin(ka)
    m = np.eye(4)
    m[:3, :3] = [
        [cf * ck, -cf * sk, sf],
        [co * sk + so * sf * ck, co * ck - so * sf * sk, -so * cf],
        [so * sk - co * sf * ck, so * ck + co * sf * sk, co * cf],
    ]
    m[:3, 3] = [tx, ty, tz]
    return m


def _emit(origin_w, dirs_w, sensor_pose, ring, seed, noise_sigma, chunk=1 << 18) -> np.ndarray:
    """Ray-cast and express the hits in the sensor frame as PointXYZIRNLRGB records."""
    n = dirs_w.shape[0]
    rng = np.random.Generator(np.random.PCG64(seed))
    pts = np.zeros(n, dtype=POINT_DTYPE)
    Rinv = sensor_pose[:3, :3].T
    tinv = -Rinv @ sensor_pose[:3, 3]
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        d = dirs_w[s:e]
        o = origin_w if origin_w.ndim == 1 else origin_w[s:e]
        t, nrm, lab = _raycast(o, d)
        t = t + rng.normal(0.0, noise_sigma, size=t.shape)
        pw = o + t[:, None] * d
        pl = pw @ Rinv.T + tinv
        nl = nrm @ Rinv.T
        nl32 = _shrink_normals(nl)
        pts["x"][s:e] = pl[:, 0]
        pts["y"][s:e] = pl[:, 1]
        pts["z"][s:e] = pl[:, 2]
        pts["normal_x"][s:e] = nl32[:, 0]
        pts["normal_y"][s:e] = nl32[:, 1]
        pts["normal_z"][s:e] = nl32[:, 2]
        pts["label"][s:e] = lab
        pts["intensity"][s:e] = np.clip(1.0 / np.maximum(t, 0.5), 0.0, 1.0)
    pts["ring"] = ring.astype(np.uint16)
    return pts


def hdl32_scan(sensor_pose: np.ndarray | None = None, seed: int = 42, n_azimuth: int = 2048,
               noise_sigma: float = 0.01) -> np.ndarray:
    """HDL-32E-like scan: 32 rings x n_azimuth steps, firing order (azimuth-major), sensor frame."""
    if sensor_pose is None:
        sensor_pose = pose_matrix(0, 0, 2.0, 0, 0, 0)
    elev = np.deg2rad(np.linspace(-30.67, 10.67, 32))
    az = np.arange(n_azimuth) * (2.0 * math.pi / n_azimuth)
    A, E = np.meshgrid(az, elev, indexing="ij")
    d_local = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], axis=-1).reshape(-1, 3)
    ring = np.tile(np.arange(32), n_azimuth)
    dirs_w = d_local @ sensor_pose[:3, :3].T
    return _emit(sensor_pose[:3, 3].copy(), dirs_w, sensor_pose, ring, seed, noise_sigma)


def rotating_sick_scan(sensor_pose: np.ndarray | None = None, seed: int = 42, n_beams: int = 1024,
                       n_profiles: int = 1024, noise_sigma: float = 0.01) -> np.ndarray:
    """m3d rotating-SICK scan: n_profiles profiles of n_beams beams over a 1.1*pi sweep, unit frame."""
    if sensor_pose is None:
        sensor_pose = pose_matrix(0, 0, 2.0, 0, 0, 0)
    a = np.deg2rad(-135.0) + np.arange(n_beams) * (np.deg2rad(270.0) / n_beams)
    ang = np.arange(n_profiles) * (1.1 * math.pi / n_profiles)
    ca, sa = np.cos(a), np.sin(a)
    # laser frame in the unit frame: R = Rz(ang) * Ry(-pi/2); beam (ca, sa, 0) -> Ry(-pi/2): (0*ca.., ) see below
    # Ry(-pi/2) maps (x, y, z) -> (-z, y, x); for a planar beam (ca, sa, 0): (0, sa, ca)
    bx, by, bz = np.zeros_like(ca), sa, ca
    cang, sang = np.cos(ang)[:, None], np.sin(ang)[:, None]
    dx = cang * bx[None, :] - sang * by[None, :]
    dy = sang * bx[None, :] + cang * by[None, :]
    dz = np.broadcast_to(bz[None, :], dx.shape)
    d_local = np.stack([dx, dy, dz], axis=-1).reshape(-1, 3)
    o_local = np.array([-0.0835, 0.0, 0.1835])
    ring = np.tile(np.arange(n_beams) & 0xFFFF, n_profiles)
    R, t = sensor_pose[:3, :3], sensor_pose[:3, 3]
    pts = _emit(R @ o_local + t, d_local @ R.T, sensor_pose, ring, seed, noise_sigma)
    # aggregator self-exclusion box (m3d_aggregator.cpp:65-73, +-1 m): the scene keeps clear of it
    inside = (np.abs(pts["x"]) < 1.0) & (np.abs(pts["y"]) < 1.0) & (np.abs(pts["z"]) < 1.0)
    assert not inside.any(), "scene geometry inside the aggregator's self-exclusion box"
    return pts


#: ground-truth sensor pose of the FIRST scan of a pair and the perturbation applied to its initial guess
PAIR_TRUE_POSE = (0.0, 0.0, 2.0, 0.0, 0.0, 0.0)
PAIR_PERTURBATION = (0.10, -0.05, 0.02, 0.010, -0.015, 0.030)


def scan_pair(kind: str = "hdl32", seed: int = 42, **kw):
    """A registration pair (SURVEY.md §8d "pair perturbation").

    Returns ``(first_local, second_local, pose_first_init, pose_second, pose_first_true)`` with 4x4 float32
    row-major poses.  Both scans see the same scene from the same pose with independent range noise
    (point-to-point ICP between scans from different viewpoints is biased by the ring pattern on the
    floor, so only this set-up has the true pose as its fixed point); the first scan's initial guess is
    the truth offset by PAIR_PERTURBATION."""
    gen = {"hdl32": hdl32_scan, "sick": rotating_sick_scan}[kind]
    pose2 = pose_matrix(0, 0, 2.0, 0, 0, 0)
    pose1_true = pose_matrix(*PAIR_TRUE_POSE)
    first = gen(pose1_true, seed=seed, **kw)
    second = gen(pose2, seed=seed + 1, **kw)
    t = PAIR_TRUE_POSE
    p = PAIR_PERTURBATION
    pose1_init = pose_matrix(t[0] + p[0], t[1] + p[1], t[2] + p[2], t[3] + p[3], t[4] + p[4], t[5] + p[5])
    return (first, second, pose1_init.astype(np.float32), pose2.astype(np.float32), pose1_true.astype(np.float32))


def loop_trajectory(n_scans: int, spacing: float = 1.0, size=(25.0, 15.0), corner_radius: float = 3.0,
                    z: float = 2.0) -> np.ndarray:
    """n_scans poses (float64 4x4) every `spacing` m along a rounded-rectangle loop, yaw tangent."""
    w, h = size
    r = corner_radius
    sx, sy = w - 2 * r, h - 2 * r
    # piecewise: bottom edge (+x), corner, right edge (+y), corner, top edge (-x), corner, left edge (-y), corner
    pieces = [sx, 0.5 * math.pi * r, sy, 0.5 * math.pi * r, sx, 0.5 * math.pi * r, sy, 0.5 * math.pi * r]
    total = sum(pieces)
    poses = np.zeros((n_scans, 4, 4))
    for k in range(n_scans):
        s = (k * spacing) % total
        i = 0
        while s > pieces[i]:
            s -= pieces[i]
            i += 1
        q, local = divmod(i, 2)
        # start points / headings of the 4 straight edges
        starts = [(-sx / 2, -h / 2, 0.0), (w / 2, -sy / 2, 0.5 * math.pi), (sx / 2, h / 2, math.pi), (-w / 2, sy / 2, 1.5 * math.pi)]
        x0, y0, hd = starts[q]
        if local == 0:
            x = x0 + s * math.cos(hd)
            y = y0 + s * math.sin(hd)
            yaw = hd
        else:
            ex = x0 + pieces[2 * q] * math.cos(hd)
            ey = y0 + pieces[2 * q] * math.sin(hd)
            cx = ex - r * math.sin(hd)
            cy = ey + r * math.cos(hd)
            th = s / r
            yaw = hd + th
            x = cx + r * math.sin(yaw)
            y = cy - r * math.cos(yaw)
        poses[k] = pose_matrix(x, y, z, 0.0, 0.0, yaw)
    return poses


def _scan_job(job):
    """One scan of slam_scans()."""
    kind, pose, seed, kw = job
    return {"hdl32": hdl32_scan, "sick": rotating_sick_scan}[kind](pose, seed=seed, **kw)


_WORKER_CODE = """
import importlib.util, json, sys
import numpy as np
spec = importlib.util.spec_from_file_location("m3d_synth_worker", sys.argv[1])
synth = importlib.util.module_from_spec(spec); spec.loader.exec_module(synth)
job = json.load(open(sys.argv[2]))
truth = synth.loop_trajectory(job["n_scans"], job["spacing"])
for k in job["scans"]:
    sc = synth._scan_job((job["kind"], truth[k], job["seed"] + k, job["kw"]))
    np.save(job["dir"] + "/scan_%d.npy" % k, sc)
"""


def _scans_by_workers(todo, n_scans, kind, seed, spacing, kw, workers):
    """Generate the scans `todo` with `workers` plain child processes (python -c, this file loaded by path, results as
    .npy files in a temporary directory): independent of how the caller's main module was started."""
    import json
    import os
    import subprocess
    import sys
    import tempfile
    made = {}
    with tempfile.TemporaryDirectory(prefix="m3d_scans_") as d:
        procs = []
        for w in range(workers):
            share = todo[w::workers]
            if not share:
                continue
            jf = os.path.join(d, "job_%d.json" % w)
            with open(jf, "w") as f:
                json.dump({"n_scans": n_scans, "spacing": spacing, "kind": kind, "seed": seed, "kw": kw, "scans": share, "dir": d}, f)
            procs.append(subprocess.Popen([sys.executable, "-c", _WORKER_CODE, os.path.abspath(__file__), jf]))
        for p in procs:
            if p.wait() != 0:
                raise RuntimeError("scan generation worker failed")
        for k in todo:
            made[k] = np.load(os.path.join(d, "scan_%d.npy" % k))
    return made


def slam_scans(n_scans: int, kind: str = "hdl32", seed: int = 42, spacing: float = 1.0,
               drift_sigma_t: float = 0.05, drift_sigma_r: float = 0.01, only=None, workers: int = 1, **kw):
    """Multi-scan data set (C4/C5): scans in their local frames, true poses, drifted initial poses.
    only: iterable of scan indices to generate (the others are None) — poses are always complete, so that the ranks of a
    multi-GPU job can each generate a share of the scans and exchange them.
    workers > 1: the scans (independent numpy ray casts, seed + k each) are generated by that many child processes —
    same arrays as the serial path."""
    truth = loop_trajectory(n_scans, spacing)
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    want = None if only is None else set(int(k) for k in only)
    todo = [k for k in range(n_scans) if want is None or k in want]
    if workers > 1 and len(todo) > 1:
        made = _scans_by_workers(todo, n_scans, kind, seed, spacing, kw, min(workers, len(todo)))
    else:
        made = {k: _scan_job((kind, truth[k], seed + k, kw)) for k in todo}
    scans, init = [], np.zeros((n_scans, 4, 4), dtype=np.float32)
    for k in range(n_scans):
        scans.append(made.get(k))
        dt = rng.normal(0, drift_sigma_t, 3)
        dr = rng.normal(0, drift_sigma_r, 3)
        yaw = math.atan2(truth[k][1, 0], truth[k][0, 0])
        init[k] = pose_matrix(truth[k][0, 3] + dt[0], truth[k][1, 3] + dt[1], truth[k][2, 3] + dt[2],
                              dr[0], dr[1], yaw + dr[2]).astype(np.float32)
    return scans, truth.astype(np.float32), init


def relative_pose_error(poses, truth) -> float:
    """Gauge-free trajectory error: mean translation error of the relative pose between consecutive scans (a sweep moves
    every scan, so the absolute error against the truth also contains a rigid drift of the whole trajectory)."""
    p = np.asarray(poses, dtype=np.float64).reshape(-1, 4, 4)
    t = np.asarray(truth, dtype=np.float64).reshape(-1, 4, 4)
    e = []
    for k in range(len(p) - 1):
        rel_e = np.linalg.inv(p[k]) @ p[k + 1]
        rel_t = np.linalg.inv(t[k]) @ t[k + 1]
        e.append(np.linalg.norm((np.linalg.inv(rel_t) @ rel_e)[:3, 3]))
    return float(np.mean(e)) if e else 0.0


def concat_points(scans) -> np.ndarray:
    """Concatenate point arrays keeping the 40-byte layout (np.concatenate re-packs padded dtypes to 38 B)."""
    out = np.zeros(sum(len(s) for s in scans), dtype=POINT_DTYPE)
    o = 0
    for s in scans:
        assert s.dtype.itemsize == 40
        out[o:o + len(s)] = s
        o += len(s)
    return out


def random_cloud(n: int, seed: int = 0, extent=(10.0, 8.0, 3.0), n_labels: int = 4, unit_normals: bool = True) -> np.ndarray:
    """Unstructured random cloud for edge-case parity tests (uniform points, random normals/labels)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pts = np.zeros(n, dtype=POINT_DTYPE)
    pts["x"] = rng.uniform(-extent[0], extent[0], n)
    pts["y"] = rng.uniform(-extent[1], extent[1], n)
    pts["z"] = rng.uniform(0, extent[2], n)
    v = rng.normal(size=(n, 3))
    if unit_normals:
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        v = _shrink_normals(v)
    pts["normal_x"], pts["normal_y"], pts["normal_z"] = v[:, 0], v[:, 1], v[:, 2]
    pts["label"] = rng.integers(0, n_labels, n)
    pts["ring"] = rng.integers(0, 32, n)
    pts["intensity"] = rng.uniform(0, 1, n)
    return pts
