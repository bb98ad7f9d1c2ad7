"""numpy-level wrappers of oracle/_ref/libm3dref.so (the reference's own kernels). TEST INFRASTRUCTURE."""
import ctypes as C

import numpy as np

import oracle

_P = C.c_void_p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_P)


def nn_search_host(first, second, radius, bucket, ext=1.0, max_inner=100, max_outer=100, export=True):
    r = oracle.ref()
    nn = np.full(len(second), -7, dtype=np.int32)
    gp = np.zeros(1, dtype=oracle.GRID_PARAMS_DTYPE)
    table = np.zeros(len(first), dtype=oracle.HASH_DTYPE) if export else None
    cap = 1
    if export:
        g = oracle.grid_params(first, bucket, ext=ext)     # only to size the export buffer
        cap = int(g["number_of_buckets"][0]) + 8
    buckets = np.zeros(cap, dtype=oracle.BUCKET_DTYPE) if export else None
    st = r.ref_nn_search_host(_ptr(first), C.c_int(len(first)), _ptr(second), C.c_int(len(second)), C.c_float(radius),
                              C.c_float(bucket), C.c_float(ext), C.c_int(max_inner), C.c_int(max_outer), _ptr(nn), _ptr(gp),
                              _ptr(table), _ptr(buckets), C.c_longlong(cap))
    assert st == 0, st
    if export:
        buckets = buckets[: int(gp["number_of_buckets"][0])].copy()
    return nn, gp, table, buckets


def grid_params_host(cloud, rx, ry, rz, ext):
    gp = np.zeros(1, dtype=oracle.GRID_PARAMS_DTYPE)
    st = oracle.ref().ref_grid_params_host(_ptr(cloud), C.c_int(len(cloud)), C.c_float(rx), C.c_float(ry), C.c_float(rz), C.c_float(ext), _ptr(gp))
    assert st == 0, st
    return gp


def transform_host(cloud, m):
    out = cloud.copy()
    m = np.ascontiguousarray(m, dtype=np.float32).reshape(-1)[:12].copy()
    st = oracle.ref().ref_transform_host(_ptr(out), C.c_int(len(out)), _ptr(m))
    assert st == 0, st
    return out


def normal_equations_host(obs, pose6, dof):
    p = np.asarray(pose6, dtype=np.float64).copy()
    N = np.zeros(dof * dof)
    b = np.zeros(dof)
    st = oracle.ref().ref_normal_equations_host(_ptr(obs), C.c_int(len(obs)), _ptr(p), C.c_int(dof), _ptr(N), _ptr(b))
    assert st == 0, st
    return N.reshape(dof, dof).T.copy(), b


def register_ls_host(obs, pose6, dof):
    p = np.asarray(pose6, dtype=np.float64).copy()
    x = np.zeros(6)
    st = oracle.ref().ref_register_ls_host(_ptr(obs), C.c_int(len(obs)), _ptr(p), C.c_int(dof), _ptr(x))
    return st, p, x[:dof]


# -- pre-registration steps: the reference's kernels behind the call sequences of cudaWrapper.cpp:118-342, 662-836 --
def remove_noise_host(cloud, res, ext, threshold):
    markers = np.zeros(len(cloud), dtype=np.uint8)
    st = oracle.ref().ref_remove_noise_host(_ptr(cloud), C.c_int(len(cloud)), C.c_float(res), C.c_float(ext), C.c_int(threshold), _ptr(markers))
    assert st == 0, st
    return markers


def downsample_host(cloud, res, ext):
    markers = np.zeros(len(cloud), dtype=np.uint8)
    st = oracle.ref().ref_downsample_host(_ptr(cloud), C.c_int(len(cloud)), C.c_float(res), C.c_float(ext), _ptr(markers))
    assert st == 0, st
    return markers


def classify_host(cloud, radius, curvature_threshold, ground_z, plane_points, ext, max_inner, max_outer, viewpoint=(0.0, 0.0, 0.0)):
    out = np.ascontiguousarray(cloud).copy()
    mean = np.zeros((len(out), 3), dtype=np.float32)
    table = np.zeros(len(out), dtype=oracle.HASH_DTYPE)
    st = oracle.ref().ref_classify_host(_ptr(out), C.c_int(len(out)), C.c_float(radius), C.c_float(curvature_threshold), C.c_float(ground_z),
                                        C.c_int(plane_points), C.c_float(ext), C.c_int(max_inner), C.c_int(max_outer),
                                        C.c_float(viewpoint[0]), C.c_float(viewpoint[1]), C.c_float(viewpoint[2]), _ptr(mean), _ptr(table))
    assert st == 0, st
    return out, mean, table


def find_best_yaw_host(first, second, second_transform, first_transform_inverse, bucket, ext, radius, max_inner, max_outer,
                       angle_start, angle_finish, angle_step):
    angles, mats = oracle.yaw_matrices(angle_start, angle_finish, angle_step)
    counts = np.zeros(len(angles), dtype=np.int32)
    best = C.c_int(-1)
    a = None if second_transform is None else np.ascontiguousarray(np.asarray(second_transform, dtype=np.float32).reshape(-1)[:12])
    b = None if first_transform_inverse is None else np.ascontiguousarray(np.asarray(first_transform_inverse, dtype=np.float32).reshape(-1)[:12])
    st = oracle.ref().ref_find_best_yaw_host(_ptr(first), C.c_int(len(first)), _ptr(second), C.c_int(len(second)), _ptr(a), _ptr(b),
                                             C.c_float(bucket), C.c_float(ext), C.c_float(radius), C.c_int(max_inner), C.c_int(max_outer),
                                             _ptr(mats), C.c_int(len(angles)), _ptr(counts), C.byref(best))
    assert st == 0, st
    return (float(angles[best.value]) if best.value >= 0 else float(angle_start)), counts
