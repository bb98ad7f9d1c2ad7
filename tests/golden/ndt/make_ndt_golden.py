"""Independent numpy / scipy statement of the NDT normal equations, and the fixture it produces.

NDT does not exist in the reference: its definition of record is BASELINE.json's north-star sentence —
"per-bucket mean and covariance" of the gridded cloud, point-to-distribution observation equations assembled into the
same 6x6 / 6x1 system with the 3x3 block weight Sigma^-1 — which oracle/m3d_oracle.c (orc_ndt_normal_equations) and the CUDA
kernels (k_ndt_*) both implement.  So that the two are not only checked against each other, THIS script states the same
mathematics a third time with library routines only (np.mean, np.cov, np.linalg.inv, scipy's Rotation for R and a
central difference for dR/d(angles)) and writes tests/golden/ndt/ndt_planes.npz:

    python tests/golden/ndt/make_ndt_golden.py

Inputs are generated here from a seeded numpy Generator (three noisy planes, 3 000 + 2 000 points) and stored in the
fixture, so neither the product's nor the oracle's generators are involved.  It imports NOTHING from this repository.

Definition (res = bucket size, ext = box extension):
  grid       min/max of the gridded (global) cloud widened by ext (float32, as cudaCalculateGridParams, lesson_16.cu:64-91),
             cell = trunc((v - min) / res) in float32, bucket = ix * nbY * nbZ + iy * nbZ + iz
  bucket b   with n >= 5 points: mu_g = mean of its global points, mu_l = mean of the same points in the scan's local frame,
             Sigma = unbiased covariance of the global points + (0.05 res)^2 I, W = Sigma^-1 (skipped if not positive)
  query q    inside the widened box whose bucket is usable: l = mu_g - q, A = -[I | J], J = d(R mu_l)/d(om, fi, ka) with
             R = Rx(om) Ry(fi) Rz(ka);  N += A^T W A, rhs += A^T W l, n_obs += 1
  output     21 upper-triangular entries of N (row-major), 6 of rhs, n_obs; x = solution of N x = rhs.
"""
import os

import numpy as np
from scipy.spatial.transform import Rotation

HERE = os.path.dirname(os.path.abspath(__file__))


def planes(rng, n, shift):
    """points on three noisy planes (floor, two walls) in a 6 x 5 x 2.5 m corner"""
    k = n // 3
    a = np.column_stack([rng.uniform(0, 6, k), rng.uniform(0, 5, k), rng.normal(0, 0.01, k)])
    b = np.column_stack([rng.uniform(0, 6, k), rng.normal(0, 0.01, k), rng.uniform(0, 2.5, k)])
    c = np.column_stack([rng.normal(0, 0.01, n - 2 * k), rng.uniform(0, 5, n - 2 * k), rng.uniform(0, 2.5, n - 2 * k)])
    return (np.vstack([a, b, c]) + shift).astype(np.float32)


def rot(om, fi, ka):
    return Rotation.from_euler("XYZ", [om, fi, ka]).as_matrix()      # intrinsic X-Y'-Z'' = Rx(om) Ry(fi) Rz(ka)


def ndt_normal_equations(local, pose6, queries, res, ext):
    t, ang = np.asarray(pose6[:3], dtype=np.float64), np.asarray(pose6[3:], dtype=np.float64)
    R = rot(*ang)
    # the transformed cloud as float32 values (what the registration stores / grids)
    glob = (local.astype(np.float64) @ R.T + t).astype(np.float32)
    mn = glob.min(axis=0) - np.float32(ext)
    mx = glob.max(axis=0) + np.float32(ext)
    nb = ((mx - mn) / np.float32(res) + np.float32(1.0)).astype(np.int32)
    cell = ((glob - mn) / np.float32(res)).astype(np.int32)
    key = cell[:, 0] * nb[1] * nb[2] + cell[:, 1] * nb[2] + cell[:, 2]
    stats = {}
    eps = (0.05 * res) ** 2
    for b in np.unique(key):
        sel = key == b
        if sel.sum() < 5:
            continue
        g = glob[sel].astype(np.float64)
        S = np.cov(g.T, ddof=1) + eps * np.eye(3)
        if np.linalg.det(S) <= 0:
            continue
        stats[int(b)] = (g.mean(axis=0), local[sel].astype(np.float64).mean(axis=0), np.linalg.inv(S))
    h = 1e-6
    dR = [(rot(*(ang + h * e)) - rot(*(ang - h * e))) / (2 * h) for e in np.eye(3)]
    N, rhs, n_obs = np.zeros((6, 6)), np.zeros(6), 0
    for q in queries:
        if (q < mn).any() or (q > mx).any():
            continue
        c = ((q - mn) / np.float32(res)).astype(np.int32)
        b = int(c[0] * nb[1] * nb[2] + c[1] * nb[2] + c[2])
        if b not in stats:
            continue
        mu_g, mu_l, W = stats[b]
        J = np.column_stack([d @ mu_l for d in dR])
        A = -np.hstack([np.eye(3), J])
        l = mu_g - q.astype(np.float64)
        N += A.T @ W @ A
        rhs += A.T @ W @ l
        n_obs += 1
    return N, rhs, n_obs, glob


def main():
    rng = np.random.default_rng(20240611)
    local = planes(rng, 3000, np.array([-2.0, -1.5, 0.0]))
    # Two identical points below everything else in x, y and z: they define the box minimum and share the first bucket.
    # The reference's bucket table has a first-element quirk (lesson_16.cu:148-158: if sorted element 0 is alone in its bucket,
    # the NEXT occupied bucket is lost), and the point that defines the minimum sits exactly `ext` from it, i.e. on a cell
    # boundary when ext == res — one ulp of difference in the transform would decide whether the quirk fires.  A pair of
    # duplicates takes that sensitivity out of the fixture (it is a property of the reference's grid, not of NDT).
    local = np.vstack([np.array([[-3.0, -2.5, -0.5], [-3.0, -2.5, -0.5]], dtype=np.float32), local])
    queries = planes(rng, 2000, np.array([-1.97, -1.52, 0.01]))
    pose6 = np.array([0.04, -0.03, 0.02, 0.006, -0.004, 0.012])
    res, ext = 1.0, 1.0
    N, rhs, n_obs, glob = ndt_normal_equations(local, pose6, queries, res, ext)
    x = np.linalg.solve(N, rhs)
    packed = np.concatenate([N[np.triu_indices(6)], rhs, [n_obs]])
    np.savez_compressed(os.path.join(HERE, "ndt_planes.npz"), local=local, glob=glob, queries=queries, pose6=pose6, res=res, ext=ext,
                        neq28=packed, x=x, n_obs=n_obs)
    print("n_obs", n_obs, "x", x)


if __name__ == "__main__":
    main()
