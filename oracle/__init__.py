"""ctypes binding of the CPU oracle (oracle/m3d_oracle.c) and of the compiled reference
(oracle/_ref/libm3dref.so).  TEST INFRASTRUCTURE: importable only from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs.  The product package never imports this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "_build", "libm3d_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libm3dref.so")

_P = C.c_void_p
_lib = None
_ref = None


class RegParams(C.Structure):
    _fields_ = [("search_radius", C.c_float), ("bucket_size", C.c_float), ("bbox_extension", C.c_float),
                ("max_inner", C.c_int32), ("max_outer", C.c_int32), ("obs_threshold", C.c_int32),
                ("weight", C.c_float * 4), ("dof", C.c_int32), ("mode", C.c_int32)]


def default_params(radius=0.5, bucket=None, dof=6, mode=0) -> RegParams:
    """Reference defaults (include/gpu6DSLAM.h:179-210): ext 1.0, INNER/OUTER 100, threshold 100, weights 10/1/10/10."""
    p = RegParams()
    p.search_radius = radius
    p.bucket_size = radius if bucket is None else bucket
    p.bbox_extension = 1.0
    p.max_inner = 100
    p.max_outer = 100
    p.obs_threshold = 100
    p.weight[:] = [10.0, 1.0, 10.0, 10.0]
    p.dof = dof
    p.mode = mode
    return p


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc) if missing/stale.  Building the checker is not using it."""
    src = os.path.join(_HERE, "m3d_oracle.c")
    hdr = os.path.join(_HERE, "m3d_oracle.h")
    stale = (not os.path.exists(ORACLE_SO)) or any(
        os.path.getmtime(f) > os.path.getmtime(ORACLE_SO) for f in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


def build_ref() -> str | None:
    """Compile the reference's own sources (only where /root/reference exists)."""
    if os.path.isdir("/root/reference/gpu_6dslam/gpu_6dslam/src"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return REF_SO if os.path.exists(REF_SO) else None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(ORACLE_SO)
        L.orc_num_threads.restype = C.c_int
        L.orc_angle_gate.argtypes = [C.c_float]
        L.orc_angle_gate.restype = C.c_int
        L.orc_nn_count_evaluations.restype = C.c_int64
        L.orc_build_observations.restype = C.c_int
        _lib = L
    return _lib


def ref_available() -> bool:
    return os.path.exists(REF_SO)


def ref() -> C.CDLL:
    global _ref
    if _ref is None:
        _ref = C.CDLL(REF_SO)
    return _ref


def _ptr(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    if a.dtype.names and "normal_x" in a.dtype.names:
        assert a.dtype.itemsize == 40, "point array lost its 40-byte layout (np.concatenate re-packs it)"
    return a.ctypes.data_as(_P)


# --------------------------------------------------------------------------------------------------
# numpy-level wrappers of the oracle
# --------------------------------------------------------------------------------------------------
from importlib import import_module as _imp  # noqa: E402

_synth = _imp("mandala-mapping_b200.synth")
POINT_DTYPE, HASH_DTYPE, BUCKET_DTYPE, OBS_DTYPE, GRID_PARAMS_DTYPE = (
    _synth.POINT_DTYPE, _synth.HASH_DTYPE, _synth.BUCKET_DTYPE, _synth.OBS_DTYPE, _synth.GRID_PARAMS_DTYPE)


def grid_params(cloud, rx, ry=None, rz=None, ext=1.0):
    ry = rx if ry is None else ry
    rz = rx if rz is None else rz
    out = np.zeros(1, dtype=GRID_PARAMS_DTYPE)
    lib().orc_grid_params_compute(_ptr(cloud), C.c_int(len(cloud)), C.c_float(rx), C.c_float(ry), C.c_float(rz),
                                  C.c_float(ext), _ptr(out))
    return out


def bucket_keys(cloud, gp):
    keys = np.zeros(len(cloud), dtype=np.int32)
    lib().orc_bucket_keys(_ptr(cloud), C.c_int(len(cloud)), _ptr(gp), _ptr(keys))
    return keys


def build_grid(cloud, gp):
    nb = int(gp["number_of_buckets"][0])
    buckets = np.zeros(nb, dtype=BUCKET_DTYPE)
    table = np.zeros(len(cloud), dtype=HASH_DTYPE)
    lib().orc_build_grid(_ptr(cloud), C.c_int(len(cloud)), _ptr(gp), _ptr(buckets), _ptr(table))
    return buckets, table


def nn_search(first, second, table, buckets, gp, radius, max_inner=100, max_outer=100):
    nn = np.zeros(len(second), dtype=np.int32)
    lib().orc_nn_search(_ptr(first), C.c_int(len(first)), _ptr(second), C.c_int(len(second)), _ptr(table),
                        _ptr(buckets), _ptr(gp), C.c_float(radius), C.c_int(max_inner), C.c_int(max_outer), _ptr(nn))
    return nn


def nn_count_evaluations(second, buckets, gp, max_inner=100, max_outer=100) -> int:
    return int(lib().orc_nn_count_evaluations(_ptr(second), C.c_int(len(second)), _ptr(buckets), _ptr(gp),
                                              C.c_int(max_inner), C.c_int(max_outer)))


def semantic_nn(first, second, radius, bucket=None, ext=1.0, max_inner=100, max_outer=100):
    """CCudaWrapper::semanticNearestNeighbourhoodSearch restated (cudaWrapper.cpp:344-424)."""
    bucket = radius if bucket is None else bucket
    gp = grid_params(first, bucket, ext=ext)
    buckets, table = build_grid(first, gp)
    nn = nn_search(first, second, table, buckets, gp, radius, max_inner, max_outer)
    return nn, gp, table, buckets


def build_observations(first_global, first_local, second_global, nn, weights=(10.0, 1.0, 10.0, 10.0)):
    obs = np.zeros(max(1, len(second_global)), dtype=OBS_DTYPE)
    w = np.asarray(weights, dtype=np.float32)
    n = lib().orc_build_observations(_ptr(first_global), _ptr(first_local), _ptr(second_global),
                                     C.c_int(len(second_global)), _ptr(nn), _ptr(w), _ptr(obs))
    return obs[:n].copy()


def normal_equations(obs, pose6, dof=6):
    N = np.zeros(dof * dof)
    b = np.zeros(dof)
    p = np.asarray(pose6, dtype=np.float64)
    lib().orc_normal_equations(_ptr(obs), C.c_int(len(obs)), _ptr(p), C.c_int(dof), _ptr(N), _ptr(b))
    return N.reshape(dof, dof).T.copy(), b  # column-major -> [row, col]


def chol_solve(N, b):
    n = len(b)
    A = np.asfortranarray(np.asarray(N, dtype=np.float64))
    Af = np.ascontiguousarray(A.T).reshape(-1)  # column-major buffer
    x = np.zeros(n)
    bb = np.asarray(b, dtype=np.float64)
    info = lib().orc_chol_solve(_ptr(Af), _ptr(bb), C.c_int(n), _ptr(x))
    return info, x


def register_ls(obs, pose6, dof=6):
    p = np.array(pose6, dtype=np.float64)
    x = np.zeros(6)
    st = lib().orc_register_ls(_ptr(obs), C.c_int(len(obs)), _ptr(p), C.c_int(dof), _ptr(x))
    return st, p, x[:dof]


def matrix4_to_euler(m):
    m = np.ascontiguousarray(m, dtype=np.float32).reshape(16)
    o = np.zeros(3, dtype=np.float32)
    t = np.zeros(3, dtype=np.float32)
    lib().orc_matrix4_to_euler(_ptr(m), _ptr(o), _ptr(t))
    return o, t


def euler_to_matrix(omfika, xyz):
    o = np.ascontiguousarray(omfika, dtype=np.float32)
    t = np.ascontiguousarray(xyz, dtype=np.float32)
    m = np.zeros(16, dtype=np.float32)
    lib().orc_euler_to_matrix(_ptr(o), _ptr(t), _ptr(m))
    return m.reshape(4, 4)


def transform_cloud(cloud, m):
    m = np.ascontiguousarray(m, dtype=np.float32).reshape(16)
    out = np.zeros_like(cloud)
    lib().orc_transform_cloud(_ptr(cloud), _ptr(out), C.c_int(len(cloud)), _ptr(m))
    return out


def ndt_normal_equations(first_global, first_local, second_global, table, buckets, gp, pose6):
    neq = np.zeros(28)
    p = np.asarray(pose6, dtype=np.float64).copy()
    lib().orc_ndt_normal_equations.restype = C.c_int64
    n = lib().orc_ndt_normal_equations(_ptr(first_global), _ptr(first_local), C.c_int(len(first_global)), _ptr(second_global),
                                       C.c_int(len(second_global)), _ptr(table), _ptr(buckets), _ptr(gp), _ptr(p), _ptr(neq))
    return int(n), neq


def solve_packed(neq, dof=6):
    x = np.zeros(6)
    q = np.ascontiguousarray(neq, dtype=np.float64)
    st = lib().orc_solve_packed(_ptr(q), C.c_int(dof), _ptr(x))
    return st, x[:dof]


def icp_iteration(first_local, second_global, pose, params: RegParams, want_nn=False):
    pose = np.ascontiguousarray(pose, dtype=np.float32).reshape(16).copy()
    scratch = np.zeros_like(first_local)
    nn = np.zeros(len(second_global), dtype=np.int32) if want_nn else None
    n_obs = C.c_int64(0)
    x = np.zeros(6)
    st = lib().orc_icp_iteration(_ptr(first_local), C.c_int(len(first_local)), _ptr(second_global),
                                 C.c_int(len(second_global)), _ptr(pose), C.byref(params), _ptr(scratch),
                                 _ptr(nn) if want_nn else None, C.byref(n_obs), _ptr(x))
    return st, pose.reshape(4, 4), int(n_obs.value), x, nn


def register_all_sweep(scans, poses, params: RegParams, pair_thr=10.0, first_optimised=0):
    off = np.zeros(len(scans) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(s) for s in scans])
    allp = _synth.concat_points(scans)
    poses = np.ascontiguousarray(poses, dtype=np.float32).reshape(len(scans), 16).copy()
    neq = np.zeros((len(scans), 28))
    status = np.zeros(len(scans), dtype=np.int32)
    lib().orc_register_all_sweep_last(_ptr(allp), _ptr(off), C.c_int(len(scans)), _ptr(poses), C.byref(params),
                                      C.c_float(pair_thr), C.c_int(first_optimised), _ptr(neq), _ptr(status))
    return poses.reshape(-1, 4, 4), neq, status


# -- pre-registration steps ------------------------------------------------------
def _m34(m):
    return None if m is None else np.ascontiguousarray(np.asarray(m, dtype=np.float32).reshape(-1)[:12])


def _optr(a):
    return None if a is None else _ptr(a)


def remove_noise(cloud, res, ext, threshold):
    """-> (filtered cloud, markers) as CCudaWrapper::removeNoiseNaive (cudaWrapper.cpp:118-179)."""
    markers = np.zeros(len(cloud), dtype=np.uint8)
    lib().orc_remove_noise_markers(_ptr(cloud), C.c_int(len(cloud)), C.c_float(res), C.c_float(ext), C.c_int(threshold), _ptr(markers))
    return np.ascontiguousarray(cloud[markers != 0]), markers


def downsample(cloud, res, ext):
    markers = np.zeros(len(cloud), dtype=np.uint8)
    lib().orc_downsample_markers(_ptr(cloud), C.c_int(len(cloud)), C.c_float(res), C.c_float(ext), _ptr(markers))
    return np.ascontiguousarray(cloud[markers != 0]), markers


def classify(cloud, radius, curvature_threshold, ground_z, plane_points, ext, max_inner, max_outer, viewpoint=(0.0, 0.0, 0.0)):
    """-> (classified copy, d_mean per sorted position, sorted table) as CCudaWrapper::classify (cudaWrapper.cpp:264-342)."""
    out = np.ascontiguousarray(cloud).copy()
    mean = np.zeros((len(out), 3), dtype=np.float32)
    table = np.zeros(len(out), dtype=HASH_DTYPE)
    lib().orc_classify(_ptr(out), C.c_int(len(out)), C.c_float(radius), C.c_float(curvature_threshold), C.c_float(ground_z),
                       C.c_int(plane_points), C.c_float(ext), C.c_int(max_inner), C.c_int(max_outer),
                       C.c_float(viewpoint[0]), C.c_float(viewpoint[1]), C.c_float(viewpoint[2]), _ptr(mean), _ptr(table))
    return out, mean, table


def yaw_matrices(angle_start, angle_finish, angle_step):
    """Angles (float accumulation as the reference's loop, cudaWrapper.cpp:761) and their row-major 3x4 yaw matrices
    (AngleAxis product = quaternion path = euler_to_matrix(0, 0, rad))."""
    angles, mats = [], []
    a = np.float32(angle_start)
    while a <= np.float32(angle_finish):
        rad = np.float32(float(a) * np.pi / 180.0)
        m = euler_to_matrix(np.array([0.0, 0.0, rad], dtype=np.float32), np.zeros(3, dtype=np.float32))
        angles.append(float(a)); mats.append(np.asarray(m, dtype=np.float32).reshape(-1)[:12].copy())
        a = np.float32(a + np.float32(angle_step))
    return np.array(angles, dtype=np.float32), np.ascontiguousarray(np.stack(mats))


def find_best_yaw(first, second, second_transform=None, first_transform_inverse=None, bucket=1.0, ext=1.0, radius=1.0,
                  max_inner=100, max_outer=100, angle_start=-30.0, angle_finish=30.0, angle_step=0.5):
    """-> (best angle, its count, counts per angle) as CCudaWrapper::findBestYaw (cudaWrapper.cpp:662-836)."""
    angles, mats = yaw_matrices(angle_start, angle_finish, angle_step)
    counts = np.zeros(len(angles), dtype=np.int32)
    L = lib()
    L.orc_find_best_yaw.restype = C.c_int
    a, b = _m34(second_transform), _m34(first_transform_inverse)
    best = L.orc_find_best_yaw(_ptr(first), C.c_int(len(first)), _ptr(second), C.c_int(len(second)), _optr(a), _optr(b),
                               C.c_float(bucket), C.c_float(ext), C.c_float(radius), C.c_int(max_inner), C.c_int(max_outer),
                               _ptr(mats), C.c_int(len(angles)), _ptr(counts))
    return (float(angles[best]) if best >= 0 else float(angle_start)), (int(counts[best]) if best >= 0 else 0), counts
