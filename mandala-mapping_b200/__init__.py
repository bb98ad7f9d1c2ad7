"""m3dreg-b200: B200-native registration hot path of gpu_6dslam behind the reference's CUDA-wrapper surface.

This Python package is only the thin host-side binding used by the tests, ``bench.py`` and the multi-GPU
driver: the product is ``libm3dreg.so`` (hand-written sm_100a kernels + the C ABI of ``include/m3dreg.h``).
The package name contains a hyphen (it is mandated by the repository layout), so import it with::

    import importlib
    m3d = importlib.import_module("mandala-mapping_b200")

There is NO CPU fallback: :func:`lib` raises if the CUDA library is missing, and ``m3dreg_create`` fails
without an sm_100 device.  Nothing in here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import synth
from .synth import BUCKET_DTYPE, GRID_PARAMS_DTYPE, HASH_DTYPE, OBS_DTYPE, POINT_DTYPE  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("M3DREG_LIB_PATH") or os.path.join(_HERE, "libm3dreg.so")   # override: tuning builds only
CSRC = os.path.join(_HERE, "csrc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared"]

# status codes (include/m3dreg.h)
OK, E_INVALID_ARG, E_TOO_MANY_BUCKETS, E_NOT_SPD, E_TOO_FEW_OBS, E_BAD_SLOT, E_NO_DEVICE, E_SIZE_MISMATCH = 0, -1, -2, -3, -4, -5, -6, -7
E_NO_NCCL, E_NCCL = -8, -9
MODE_ICP, MODE_NDT = 0, 1

#: every symbol include/m3dreg.h declares (checked by tests/test_cabi.py against the built library)
EXPORTS = [
    "m3dreg_version", "m3dreg_status_string", "m3dreg_create", "m3dreg_destroy", "m3dreg_warm_up",
    "m3dreg_set_stream", "m3dreg_get_stream", "m3dreg_synchronize", "m3dreg_launch_count",
    "m3dreg_calculate_grid_params", "m3dreg_calculate_grid", "m3dreg_nn_search", "m3dreg_transform",
    "m3dreg_normal_equations", "m3dreg_solve_chol", "m3dreg_solve_observations",
    "m3dreg_semantic_nn_host", "m3dreg_register_ls_host", "m3dreg_matrix4_to_euler", "m3dreg_euler_to_matrix",
    "m3dreg_scan_upload", "m3dreg_scan_size", "m3dreg_scan_clear", "m3dreg_icp_pair", "m3dreg_icp_iteration_host",
    "m3dreg_export_last_grid", "m3dreg_export_last_nn", "m3dreg_sweep_zero", "m3dreg_sweep_accumulate",
    "m3dreg_sweep_solve", "m3dreg_icp_begin", "m3dreg_icp_step", "m3dreg_icp_end", "m3dreg_icp_copy_neq", "m3dreg_icp_set_neq_out",
    "m3dreg_set_profiling", "m3dreg_get_stage_ms", "m3dreg_set_pruning", "m3dreg_get_nn_evaluations", "m3dreg_get_nn_fallbacks",
    "m3dreg_get_grid_phase_ns", "m3dreg_slam_sweep", "m3dreg_slam_copy_neq", "m3dreg_slam_plan", "m3dreg_slam_plan_measured", "m3dreg_nccl_get_unique_id",
    "m3dreg_nccl_init", "m3dreg_nccl_attach",
    "m3dreg_remove_noise_host", "m3dreg_downsample_host", "m3dreg_classify_host", "m3dreg_find_best_yaw_host",
]
#: every symbol include/m3dreg_node.h declares
NODE_EXPORTS = [
    "m3dreg_pcd_write_binary", "m3dreg_pcd_read", "m3dreg_model_create", "m3dreg_model_destroy", "m3dreg_model_load", "m3dreg_model_save",
    "m3dreg_model_set_algorithm_name", "m3dreg_model_set_dataset_path", "m3dreg_model_get_dataset_path", "m3dreg_model_set_affine",
    "m3dreg_model_get_affine", "m3dreg_model_set_cloud_name", "m3dreg_model_get_cloud_name", "m3dreg_model_scan_count", "m3dreg_model_scan_id",
    "m3dreg_model_full_cloud_path", "m3dreg_node_default_params", "m3dreg_node_create", "m3dreg_node_destroy",
    "m3dreg_node_register_single_scan", "m3dreg_node_scan_count", "m3dreg_node_get_pose", "m3dreg_node_scan_size", "m3dreg_node_get_scan",
    "m3dreg_node_scan_id", "m3dreg_node_metascan", "m3dreg_node_register_all", "m3dreg_node_load_map", "m3dreg_node_set_initial_pose",
]


class RegParams(C.Structure):
    """m3dreg_reg_params (include/m3dreg.h) = hot-path subset of the reference's ROS params."""
    _fields_ = [("search_radius", C.c_float), ("bucket_size", C.c_float), ("bbox_extension", C.c_float),
                ("max_inner", C.c_int32), ("max_outer", C.c_int32), ("obs_threshold", C.c_int32),
                ("weight", C.c_float * 4), ("dof", C.c_int32), ("mode", C.c_int32)]


class IcpStats(C.Structure):
    _fields_ = [("iterations_run", C.c_int32), ("last_status", C.c_int32), ("n_obs_last", C.c_int64),
                ("n_buckets_last", C.c_int64), ("x_last", C.c_double * 6), ("device_ms", C.c_float)]


class SlamParams(C.Structure):
    """m3dreg_slam_params (include/m3dreg.h)."""
    _fields_ = [("reg", RegParams), ("distance_threshold", C.c_float), ("first_optimised", C.c_int32)]


class SweepStats(C.Structure):
    _fields_ = [("n_pairs", C.c_int64), ("n_pairs_mine", C.c_int64), ("points_all", C.c_int64), ("points_mine", C.c_int64),
                ("accumulate_ms", C.c_float), ("allreduce_ms", C.c_float), ("rank", C.c_int32), ("world", C.c_int32)]


def slam_plan(poses, sizes, distance_threshold: float = 10.0, first_optimised: int = 0, world: int = 1):
    """Pairs of a sweep and the rank owning each (m3dreg_slam_plan: pure host code, no GPU needed)."""
    poses = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, 16)
    sizes = np.ascontiguousarray(sizes, dtype=np.int32)
    n = len(poses)
    cnt = lib().m3dreg_slam_plan(_p(poses), C.c_int(n), _p(sizes), C.c_float(distance_threshold), C.c_int(first_optimised), C.c_int(world),
                                 None, None, None, C.c_int(0))
    if cnt < 0:
        raise M3dRegError(cnt, "m3dreg_slam_plan")
    pi, pj, ow = (np.zeros(max(cnt, 1), dtype=np.int32) for _ in range(3))
    rc = lib().m3dreg_slam_plan(_p(poses), C.c_int(n), _p(sizes), C.c_float(distance_threshold), C.c_int(first_optimised), C.c_int(world),
                                _p(pi), _p(pj), _p(ow), C.c_int(len(pi)))
    if rc < 0:
        raise M3dRegError(rc, "m3dreg_slam_plan")
    return pi[:cnt], pj[:cnt], ow[:cnt]


def slam_plan_measured(poses, sizes, cost_per_pair, distance_threshold: float = 10.0, first_optimised: int = 0, world: int = 1):
    """The plan of a sweep balanced by measured device time per pair of every scan's group (m3dreg_slam_plan_measured)."""
    poses = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, 16)
    sizes = np.ascontiguousarray(sizes, dtype=np.int32)
    cpp = np.ascontiguousarray(cost_per_pair, dtype=np.float64)
    n = len(poses)
    pi, pj = slam_plan(poses, sizes, distance_threshold, first_optimised, 1)[:2]
    ow = np.zeros(max(len(pi), 1), dtype=np.int32)
    pi2, pj2 = np.zeros_like(ow), np.zeros_like(ow)
    rc = lib().m3dreg_slam_plan_measured(_p(poses), C.c_int(n), _p(sizes), C.c_float(distance_threshold), C.c_int(first_optimised), C.c_int(world),
                                         _p(cpp), _p(pi2), _p(pj2), _p(ow), C.c_int(len(ow)))
    if rc < 0:
        raise M3dRegError(rc, "m3dreg_slam_plan_measured")
    return pi2[:rc], pj2[:rc], ow[:rc]


def nccl_unique_id() -> np.ndarray:
    """128-byte ncclUniqueId (rank 0 creates it, every rank passes it to Context.nccl_init)."""
    out = np.zeros(128, dtype=np.uint8)
    _check(lib().m3dreg_nccl_get_unique_id(_p(out)), "m3dreg_nccl_get_unique_id")
    return out


def default_params(radius: float = 0.5, bucket: float | None = None, dof: int = 6, mode: int = MODE_ICP) -> RegParams:
    """Reference defaults (gpu_6dslam/include/gpu6DSLAM.h:179-210)."""
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


def sources() -> list[str]:
    return [os.path.join(CSRC, "m3dreg.cu")]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libm3dreg.so, in-tree (nvcc cross-compiles without a GPU)."""
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "m3dreg.h"), os.path.join(ROOT, "include", "m3dreg_node.h")]
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps)
    if force or stale:
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + sources()
        subprocess.check_call(cmd)
    return LIB_PATH


class M3dRegError(RuntimeError):
    def __init__(self, status: int, where: str):
        self.status = status
        try:
            msg = lib().m3dreg_status_string(status).decode()
        except Exception:  # pragma: no cover
            msg = "?"
        super().__init__(f"{where}: status {status} ({msg})")


_lib = None


def lib() -> C.CDLL:
    """The CUDA library.  Fails loudly when it has not been built — there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the product has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.m3dreg_status_string.restype = C.c_char_p
        L.m3dreg_get_stream.restype = C.c_void_p
        L.m3dreg_launch_count.restype = C.c_int64
        L.m3dreg_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        for name in EXPORTS + NODE_EXPORTS:
            getattr(L, name)
        L.m3dreg_model_create.restype = C.c_void_p
        _lib = L
    return _lib


def _p(x):
    """void* of a numpy array / torch tensor / raw int address / None."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        if x.dtype.names and "normal_x" in x.dtype.names:
            assert x.dtype.itemsize == 40, "point array lost its 40-byte layout (np.concatenate re-packs it)"
        return C.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(int(x))


def _check(st: int, where: str, allow=()):
    if st != 0 and st not in allow:
        raise M3dRegError(st, where)
    return st


class Context:
    """Owns one m3dreg_ctx (one per device / per rank)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _check(lib().m3dreg_create(C.byref(self._h), int(device)), "m3dreg_create")
        self.device = device

    def close(self):
        if self._h:
            lib().m3dreg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # -- lifecycle -------------------------------------------------------------------------------------
    def warm_up(self):
        _check(lib().m3dreg_warm_up(self._h), "m3dreg_warm_up")

    def set_stream(self, cuda_stream: int | None):
        _check(lib().m3dreg_set_stream(self._h, C.c_void_p(cuda_stream or 0)), "m3dreg_set_stream")

    def synchronize(self):
        _check(lib().m3dreg_synchronize(self._h), "m3dreg_synchronize")

    def nn_evaluations(self, reset: bool = True) -> int:
        v = C.c_uint64(0)
        _check(lib().m3dreg_get_nn_evaluations(self._h, C.byref(v), C.c_int(1 if reset else 0)), "m3dreg_get_nn_evaluations")
        return int(v.value)

    def nn_fallbacks(self, reset: bool = True) -> int:
        """Queries handed to the per-thread search by the warp-shared NN kernel (counted while profiling is on)."""
        v = C.c_uint64(0)
        _check(lib().m3dreg_get_nn_fallbacks(self._h, C.byref(v), C.c_int(1 if reset else 0)), "m3dreg_get_nn_fallbacks")
        return int(v.value)

    def grid_phase_ns(self) -> np.ndarray:
        """%globaltimer stamps of the last k_grid_build launch (profiling on): see m3dreg_get_grid_phase_ns."""
        out = np.zeros(32, dtype=np.uint64)
        _check(lib().m3dreg_get_grid_phase_ns(self._h, _p(out)), "m3dreg_get_grid_phase_ns")
        return out

    def set_pruning(self, enabled: bool):
        _check(lib().m3dreg_set_pruning(self._h, C.c_int(1 if enabled else 0)), "m3dreg_set_pruning")

    @property
    def launch_count(self) -> int:
        return int(lib().m3dreg_launch_count(self._h))

    # -- pre-registration steps (host clouds, as CCudaWrapper::removeNoiseNaive / downsampling / classify / findBestYaw) ----
    def remove_noise(self, cloud: np.ndarray, resolution: float, ext: float, threshold: int):
        """-> (filtered cloud, markers): a point stays iff its bucket holds more than `threshold` points."""
        n = len(cloud)
        out = np.zeros(n, dtype=POINT_DTYPE); markers = np.zeros(n, dtype=np.uint8); kept = C.c_int(0)
        _check(lib().m3dreg_remove_noise_host(self._h, _p(cloud), C.c_int(n), C.c_float(resolution), C.c_float(ext), C.c_int(threshold),
                                              _p(out), C.byref(kept), _p(markers)), "m3dreg_remove_noise_host")
        return out[:kept.value].copy(), markers

    def downsample(self, cloud: np.ndarray, resolution: float, ext: float):
        """-> (one point per occupied bucket, markers)."""
        n = len(cloud)
        out = np.zeros(n, dtype=POINT_DTYPE); markers = np.zeros(n, dtype=np.uint8); kept = C.c_int(0)
        _check(lib().m3dreg_downsample_host(self._h, _p(cloud), C.c_int(n), C.c_float(resolution), C.c_float(ext),
                                            _p(out), C.byref(kept), _p(markers)), "m3dreg_downsample_host")
        return out[:kept.value].copy(), markers

    def classify(self, cloud: np.ndarray, radius: float, curvature_threshold: float, ground_z: float, plane_points: int, ext: float,
                 max_inner: int, max_outer: int, viewpoint=(0.0, 0.0, 0.0), want_debug: bool = False):
        """Normals + plane/edge/ceiling/ground labels written into a COPY of `cloud` (returned).  want_debug also returns the
        reference's d_mean (3 floats per sorted position) and the sorted table those positions refer to."""
        out = np.ascontiguousarray(cloud).copy()
        n = len(out)
        mean = np.zeros((n, 3), dtype=np.float32) if want_debug else None
        table = np.zeros(n, dtype=HASH_DTYPE) if want_debug else None
        _check(lib().m3dreg_classify_host(self._h, _p(out), C.c_int(n), C.c_float(radius), C.c_float(curvature_threshold), C.c_float(ground_z),
                                          C.c_int(plane_points), C.c_float(ext), C.c_int(max_inner), C.c_int(max_outer),
                                          C.c_float(viewpoint[0]), C.c_float(viewpoint[1]), C.c_float(viewpoint[2]), _p(mean), _p(table)),
               "m3dreg_classify_host")
        return (out, mean, table) if want_debug else out

    def find_best_yaw(self, first: np.ndarray, second: np.ndarray, second_transform=None, first_transform_inverse=None, bucket: float = 1.0,
                      ext: float = 1.0, radius: float = 1.0, max_inner: int = 100, max_outer: int = 100,
                      angle_start: float = -30.0, angle_finish: float = 30.0, angle_step: float = 0.5):
        """-> (best angle in degrees, its match count, matches per angle)."""
        def m34(m):
            return None if m is None else np.ascontiguousarray(np.asarray(m, dtype=np.float32).reshape(-1)[:12])
        a, b = m34(second_transform), m34(first_transform_inverse)
        cap = int((angle_finish - angle_start) / angle_step) + 8
        counts = np.full(cap, -1, dtype=np.int32); best = C.c_float(0.0); best_n = C.c_int(0)
        _check(lib().m3dreg_find_best_yaw_host(self._h, _p(first), C.c_int(len(first)), _p(second), C.c_int(len(second)), _p(a), _p(b),
                                               C.c_float(bucket), C.c_float(ext), C.c_float(radius), C.c_int(max_inner), C.c_int(max_outer),
                                               C.c_float(angle_start), C.c_float(angle_finish), C.c_float(angle_step),
                                               C.byref(best), C.byref(best_n), _p(counts), C.c_int(cap)), "m3dreg_find_best_yaw_host")
        return float(best.value), int(best_n.value), counts[counts >= 0].copy()

    # -- stage level (device pointers: torch CUDA tensors or raw addresses) ------------------------------
    def calculate_grid_params(self, d_cloud, n, rx, ry=None, rz=None, ext=1.0) -> np.ndarray:
        ry = rx if ry is None else ry
        rz = rx if rz is None else rz
        out = np.zeros(1, dtype=GRID_PARAMS_DTYPE)
        _check(lib().m3dreg_calculate_grid_params(self._h, _p(d_cloud), C.c_int(n), C.c_float(rx), C.c_float(ry),
                                                  C.c_float(rz), C.c_float(ext), _p(out)), "m3dreg_calculate_grid_params")
        return out

    def calculate_grid(self, d_cloud, n, gp: np.ndarray, d_buckets, d_table):
        _check(lib().m3dreg_calculate_grid(self._h, _p(d_cloud), C.c_int(n), _p(gp), _p(d_buckets), _p(d_table)),
               "m3dreg_calculate_grid")

    def nn_search(self, d_first, n1, d_second, n2, d_table, d_buckets, gp, radius, max_inner, max_outer, d_nn):
        _check(lib().m3dreg_nn_search(self._h, _p(d_first), C.c_int(n1), _p(d_second), C.c_int(n2), _p(d_table),
                                      _p(d_buckets), _p(gp), C.c_float(radius), C.c_int(max_inner), C.c_int(max_outer),
                                      _p(d_nn)), "m3dreg_nn_search")

    def transform(self, d_in, d_out, n, m3x4):
        m = np.ascontiguousarray(m3x4, dtype=np.float32).reshape(-1)[:12].copy()
        _check(lib().m3dreg_transform(self._h, _p(d_in), _p(d_out), C.c_int(n), _p(m)), "m3dreg_transform")

    def normal_equations(self, d_obs, n_obs, pose6, dof=6):
        p = np.asarray(pose6, dtype=np.float64).copy()
        N = np.zeros(dof * dof)
        b = np.zeros(dof)
        _check(lib().m3dreg_normal_equations(self._h, _p(d_obs), C.c_int(n_obs), _p(p), C.c_int(dof), _p(N), _p(b)),
               "m3dreg_normal_equations")
        return N.reshape(dof, dof).T.copy(), b

    def solve_chol(self, N, b):
        dof = len(b)
        A = np.ascontiguousarray(np.asarray(N, dtype=np.float64).T).reshape(-1)  # column-major buffer
        bb = np.asarray(b, dtype=np.float64).copy()
        x = np.zeros(dof)
        st = lib().m3dreg_solve_chol(self._h, _p(A), _p(bb), C.c_int(dof), _p(x))
        _check(st, "m3dreg_solve_chol", allow=(E_NOT_SPD,))
        return st, x

    def solve_observations(self, d_obs, n_obs, pose6, dof=6):
        p = np.asarray(pose6, dtype=np.float64).copy()
        x = np.zeros(dof)
        st = lib().m3dreg_solve_observations(self._h, _p(d_obs), C.c_int(n_obs), _p(p), C.c_int(dof), _p(x))
        _check(st, "m3dreg_solve_observations", allow=(E_NOT_SPD,))
        return st, x

    # -- CCudaWrapper level (host buffers) -----------------------------------------------------------------
    def semantic_nn_host(self, first, second, radius, bucket, ext=1.0, max_inner=100, max_outer=100, nn_out=None):
        nn = np.empty(len(second), dtype=np.int32) if nn_out is None else nn_out
        _check(lib().m3dreg_semantic_nn_host(self._h, _p(first), C.c_int(len(first)), _p(second), C.c_int(len(second)),
                                             C.c_float(radius), C.c_float(bucket), C.c_float(ext), C.c_int(max_inner),
                                             C.c_int(max_outer), _p(nn)), "m3dreg_semantic_nn_host")
        return nn

    def register_ls_host(self, obs, pose6, dof=6):
        p = np.asarray(pose6, dtype=np.float64).copy()
        x = np.zeros(6)
        st = lib().m3dreg_register_ls_host(self._h, _p(obs), C.c_int(len(obs)), _p(p), C.c_int(dof), _p(x))
        _check(st, "m3dreg_register_ls_host", allow=(E_NOT_SPD,))
        return st, p, x[:dof]

    # -- scan store + fused loops ------------------------------------------------------------------------------
    def scan_upload(self, slot: int, pts, n: int | None = None, on_device: bool = False):
        n = len(pts) if n is None else n
        _check(lib().m3dreg_scan_upload(self._h, C.c_int(slot), _p(pts), C.c_int(n), C.c_int(1 if on_device else 0)),
               "m3dreg_scan_upload")

    def scan_size(self, slot: int) -> int:
        return int(lib().m3dreg_scan_size(self._h, C.c_int(slot)))

    def scan_clear(self):
        _check(lib().m3dreg_scan_clear(self._h), "m3dreg_scan_clear")

    def icp_pair(self, first_slot, second_slot, pose_first, pose_second, params: RegParams, iterations: int):
        pf = np.ascontiguousarray(pose_first, dtype=np.float32).reshape(16).copy()
        ps = np.ascontiguousarray(pose_second, dtype=np.float32).reshape(16).copy()
        st = IcpStats()
        _check(lib().m3dreg_icp_pair(self._h, C.c_int(first_slot), C.c_int(second_slot), _p(pf), _p(ps), C.byref(params),
                                     C.c_int(iterations), C.byref(st)), "m3dreg_icp_pair")
        return pf.reshape(4, 4), st

    def icp_begin(self, first_slot, second_slot, pose_first, pose_second, params: RegParams):
        pf = np.ascontiguousarray(pose_first, dtype=np.float32).reshape(16).copy()
        ps = np.ascontiguousarray(pose_second, dtype=np.float32).reshape(16).copy()
        self._params = params   # keep alive
        _check(lib().m3dreg_icp_begin(self._h, C.c_int(first_slot), C.c_int(second_slot), _p(pf), _p(ps), C.byref(params)),
               "m3dreg_icp_begin")

    def icp_step(self, iterations: int = 1):
        _check(lib().m3dreg_icp_step(self._h, C.c_int(iterations)), "m3dreg_icp_step")

    def icp_end(self):
        pf = np.zeros(16, dtype=np.float32)
        st = IcpStats()
        _check(lib().m3dreg_icp_end(self._h, _p(pf), C.byref(st)), "m3dreg_icp_end")
        return pf.reshape(4, 4), st

    def icp_copy_neq(self, d_dst):
        _check(lib().m3dreg_icp_copy_neq(self._h, _p(d_dst)), "m3dreg_icp_copy_neq")

    def icp_set_neq_out(self, d_dst):
        """Following fused iterations also write their 28-double block to d_dst (device tensor / pointer; None: off)."""
        ptr = C.c_void_p(0) if d_dst is None else _p(d_dst)
        _check(lib().m3dreg_icp_set_neq_out(self._h, ptr), "m3dreg_icp_set_neq_out")

    def set_profiling(self, enabled: bool):
        _check(lib().m3dreg_set_profiling(self._h, C.c_int(1 if enabled else 0)), "m3dreg_set_profiling")

    def get_stage_ms(self):
        ms = np.zeros(4, dtype=np.float32)
        it = C.c_int(0)
        _check(lib().m3dreg_get_stage_ms(self._h, _p(ms), C.byref(it)), "m3dreg_get_stage_ms")
        return ms, int(it.value)

    def icp_iteration_host(self, first_local, second_global, pose_first, params: RegParams, nn_out=None,
                           pose_inout: np.ndarray | None = None):
        pf = pose_inout if pose_inout is not None else np.ascontiguousarray(pose_first, dtype=np.float32).reshape(16).copy()
        st = IcpStats()
        _check(lib().m3dreg_icp_iteration_host(self._h, _p(first_local), C.c_int(len(first_local)), _p(second_global),
                                               C.c_int(len(second_global)), _p(pf), C.byref(params), _p(nn_out),
                                               C.byref(st)), "m3dreg_icp_iteration_host")
        return pf.reshape(4, 4), st

    def export_last_grid(self, n_first: int):
        gp = np.zeros(1, dtype=GRID_PARAMS_DTYPE)
        _check(lib().m3dreg_export_last_grid(self._h, _p(gp), None, C.c_int(0), None, C.c_int64(0)), "m3dreg_export_last_grid")
        nb = int(gp["number_of_buckets"][0])
        table = np.zeros(n_first, dtype=HASH_DTYPE)
        buckets = np.zeros(nb, dtype=BUCKET_DTYPE)
        _check(lib().m3dreg_export_last_grid(self._h, _p(gp), _p(table), C.c_int(n_first), _p(buckets), C.c_int64(nb)),
               "m3dreg_export_last_grid")
        return gp, table, buckets

    def export_last_nn(self, n_second: int) -> np.ndarray:
        nn = np.zeros(n_second, dtype=np.int32)
        _check(lib().m3dreg_export_last_nn(self._h, _p(nn), C.c_int(n_second)), "m3dreg_export_last_nn")
        return nn

    # -- multi-scan sweep ----------------------------------------------------------------------------------------
    def sweep_zero(self, d_neq, n_scans):
        _check(lib().m3dreg_sweep_zero(self._h, _p(d_neq), C.c_int(n_scans)), "m3dreg_sweep_zero")

    def sweep_accumulate(self, pair_i, pair_j, poses, params: RegParams, d_neq):
        pi = np.ascontiguousarray(pair_i, dtype=np.int32)
        pj = np.ascontiguousarray(pair_j, dtype=np.int32)
        ps = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, 16)
        _check(lib().m3dreg_sweep_accumulate(self._h, C.c_int(len(pi)), _p(pi), _p(pj), _p(ps), C.c_int(len(ps)),
                                             C.byref(params), _p(d_neq)), "m3dreg_sweep_accumulate")

    def nccl_init(self, unique_id, rank: int, world: int):
        """Create this rank's NCCL communicator from a 128-byte id (m3dreg_nccl_init)."""
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        assert uid.size == 128
        _check(lib().m3dreg_nccl_init(self._h, _p(uid), C.c_int(rank), C.c_int(world)), "m3dreg_nccl_init")

    def nccl_attach(self, comm, rank: int, world: int):
        """Use an existing ncclComm_t (raw address) of the host application; world == 1 detaches."""
        _check(lib().m3dreg_nccl_attach(self._h, C.c_void_p(int(comm)) if comm else None, C.c_int(rank), C.c_int(world)), "m3dreg_nccl_attach")

    def slam_sweep(self, poses, params: RegParams, distance_threshold: float = 10.0, first_optimised: int = 0):
        """One registerAll sweep over the uploaded scans 0..n-1 (m3dreg_slam_sweep): gate, partition, accumulate, NCCL
        all-reduce of the normal-equation blocks, solve.  Returns (poses [n,4,4], status [n], SweepStats)."""
        ps = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, 16).copy()
        n = len(ps)
        sp = SlamParams()
        C.memmove(C.byref(sp.reg), C.byref(params), C.sizeof(RegParams))
        sp.distance_threshold = distance_threshold
        sp.first_optimised = first_optimised
        status = np.zeros(n, dtype=np.int32)
        st = SweepStats()
        _check(lib().m3dreg_slam_sweep(self._h, C.c_int(n), _p(ps), C.byref(sp), _p(status), C.byref(st)), "m3dreg_slam_sweep")
        return ps.reshape(-1, 4, 4), status, st

    def slam_neq(self, n_scans: int) -> np.ndarray:
        """The (all-reduced) normal-equation blocks of the last slam_sweep, [n_scans, 28] doubles."""
        out = np.zeros((n_scans, 28), dtype=np.float64)
        _check(lib().m3dreg_slam_copy_neq(self._h, _p(out), C.c_int(n_scans)), "m3dreg_slam_copy_neq")
        return out

    def sweep_solve(self, d_neq, poses, params: RegParams, begin=0, end=None):
        ps = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, 16).copy()
        n = len(ps)
        end = n if end is None else end
        status = np.zeros(n, dtype=np.int32)
        _check(lib().m3dreg_sweep_solve(self._h, _p(d_neq), C.c_int(n), C.c_int(begin), C.c_int(end), _p(ps),
                                        C.byref(params), _p(status)), "m3dreg_sweep_solve")
        return ps.reshape(-1, 4, 4), status


def matrix4_to_euler(m):
    m = np.ascontiguousarray(m, dtype=np.float32).reshape(16)
    o = np.zeros(3, dtype=np.float32)
    t = np.zeros(3, dtype=np.float32)
    lib().m3dreg_matrix4_to_euler(_p(m), _p(o), _p(t))
    return o, t


def euler_to_matrix(omfika, xyz):
    o = np.ascontiguousarray(omfika, dtype=np.float32)
    t = np.ascontiguousarray(xyz, dtype=np.float32)
    m = np.zeros(16, dtype=np.float32)
    lib().m3dreg_euler_to_matrix(_p(o), _p(t), _p(m))
    return m.reshape(4, 4)


from .wrapper import CCudaWrapper, Observations  # noqa: E402,F401


# ---------------------------------------------------------------------------------------------------------------------
# include/m3dreg_node.h: persistence formats and the per-scan driver (class gpu6DSLAM without ROS / PCL / Eigen / Boost)
# ---------------------------------------------------------------------------------------------------------------------
def pcd_write_binary(path: str, cloud: np.ndarray):
    """pcl::io::savePCDFileBinary for the 40-byte point type (packed 38-byte records, see formats_host.inl)."""
    _check(lib().m3dreg_pcd_write_binary(str(path).encode(), _p(np.ascontiguousarray(cloud)), C.c_int(len(cloud))), "m3dreg_pcd_write_binary")


def pcd_read(path: str) -> np.ndarray:
    n = C.c_int(0)
    _check(lib().m3dreg_pcd_read(str(path).encode(), None, C.c_int(0), C.byref(n)), "m3dreg_pcd_read")
    out = np.zeros(n.value, dtype=POINT_DTYPE)
    if n.value:
        _check(lib().m3dreg_pcd_read(str(path).encode(), _p(out), C.c_int(n.value), C.byref(n)), "m3dreg_pcd_read")
    return out


class Model:
    """class data_model (include/data_model.hpp): the XML pose model."""

    def __init__(self):
        self._h = C.c_void_p(lib().m3dreg_model_create())

    def close(self):
        if self._h:
            lib().m3dreg_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def load(self, path):
        return lib().m3dreg_model_load(self._h, str(path).encode()) == 0

    def save(self, path):
        _check(lib().m3dreg_model_save(self._h, str(path).encode()), "m3dreg_model_save")

    def set_algorithm_name(self, name):
        lib().m3dreg_model_set_algorithm_name(self._h, name.encode())

    def set_dataset_path(self, path):
        lib().m3dreg_model_set_dataset_path(self._h, path.encode())

    def _str(self, fn, *args):
        buf = C.create_string_buffer(4096)
        st = fn(self._h, *args, buf, C.c_int(4096))
        return None if st < 0 else buf.value.decode()

    def dataset_path(self):
        return self._str(lib().m3dreg_model_get_dataset_path)

    def set_affine(self, scan_id, m):
        lib().m3dreg_model_set_affine(self._h, scan_id.encode(), _p(np.ascontiguousarray(m, dtype=np.float32).reshape(16)))

    def affine(self, scan_id):
        m = np.zeros(16, dtype=np.float32)
        return m.reshape(4, 4) if lib().m3dreg_model_get_affine(self._h, scan_id.encode(), _p(m)) == 0 else None

    def set_cloud_name(self, scan_id, fn):
        lib().m3dreg_model_set_cloud_name(self._h, scan_id.encode(), fn.encode())

    def cloud_name(self, scan_id):
        return self._str(lib().m3dreg_model_get_cloud_name, scan_id.encode())

    def scan_ids(self):
        n = lib().m3dreg_model_scan_count(self._h)
        return [self._str(lib().m3dreg_model_scan_id, C.c_int(k)) for k in range(max(n, 0))]

    def full_cloud_path(self, scan_id):
        return self._str(lib().m3dreg_model_full_cloud_path, scan_id.encode())


class NodeParams(C.Structure):
    """m3dreg_node_params (include/m3dreg_node.h) = the public parameter members of class gpu6DSLAM."""
    _fields_ = [("noise_removal_resolution", C.c_float), ("noise_removal_number_of_points_in_bucket_threshold", C.c_int32),
                ("noise_removal_bounding_box_extension", C.c_float), ("downsampling_resolution", C.c_float),
                ("semantic_classification_normal_vectors_search_radius", C.c_float), ("semantic_classification_curvature_threshold", C.c_float),
                ("semantic_classification_ground_Z_coordinate_threshold", C.c_float),
                ("semantic_classification_number_of_points_needed_for_plane_threshold", C.c_int32),
                ("semantic_classification_max_number_considered_in_INNER_bucket", C.c_int32),
                ("semantic_classification_max_number_considered_in_OUTER_bucket", C.c_int32),
                ("semantic_classification_bounding_box_extension", C.c_float),
                ("slam_registerLastArrivedScan_distance_threshold", C.c_float), ("slam_registerAll_distance_threshold", C.c_float),
                ("slam_number_of_observations_threshold", C.c_int32),
                ("slam_search_radius_step", C.c_float * 3), ("slam_bucket_size_step", C.c_float * 3),
                ("slam_registerLastArrivedScan_number_of_iterations_step", C.c_int32 * 3), ("slam_registerAll_number_of_iterations_step", C.c_int32 * 3),
                ("slam_search_radius_register_all", C.c_float), ("slam_bucket_size_step_register_all", C.c_float),
                ("slam_bounding_box_extension", C.c_float), ("slam_max_number_considered_in_INNER_bucket", C.c_int32),
                ("slam_max_number_considered_in_OUTER_bucket", C.c_int32), ("slam_observation_weight", C.c_float * 4),
                ("findBestYaw_start_angle", C.c_float), ("findBestYaw_finish_angle", C.c_float), ("findBestYaw_step_angle", C.c_float),
                ("findBestYaw_bucket_size", C.c_float), ("findBestYaw_bounding_box_extension", C.c_float), ("findBestYaw_search_radius", C.c_float),
                ("findBestYaw_max_number_considered_in_INNER_bucket", C.c_int32), ("findBestYaw_max_number_considered_in_OUTER_bucket", C.c_int32),
                ("viewpoint", C.c_float * 3), ("cutoff_z_min", C.c_float), ("cutoff_z_max", C.c_float), ("cutoff_xy2_min", C.c_float),
                ("number_of_last_scans_in_sweeps", C.c_int32), ("dof", C.c_int32), ("use_find_best_yaw", C.c_int32), ("write_files", C.c_int32)]


class NodeScanStats(C.Structure):
    _fields_ = [("n_raw", C.c_int32), ("n_after_cutoff", C.c_int32), ("n_after_noise_removal", C.c_int32), ("n_after_downsampling", C.c_int32),
                ("pair_iterations", C.c_int32), ("pair_last_status", C.c_int32), ("sweeps", C.c_int32), ("sweep_solved_last", C.c_int32),
                ("yaw_deg", C.c_float), ("preprocess_ms", C.c_float), ("register_ms", C.c_float)]


def node_default_params() -> NodeParams:
    p = NodeParams()
    lib().m3dreg_node_default_params(C.byref(p))
    return p


class Node:
    """class gpu6DSLAM (include/gpu6DSLAM.h) replayed without ROS: registerSingleScan, getMetascan, registerAll, loadmapfromfile."""

    def __init__(self, ctx: "Context", params: NodeParams | None = None, root_folder: str | None = None):
        self._h = C.c_void_p()
        self.ctx = ctx
        _check(lib().m3dreg_node_create(C.byref(self._h), ctx._h, C.byref(params) if params is not None else None,
                                        root_folder.encode() if root_folder else None), "m3dreg_node_create")

    def close(self):
        if self._h:
            lib().m3dreg_node_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def register_single_scan(self, cloud: np.ndarray, mtf: np.ndarray, iso_time: str) -> NodeScanStats:
        st = NodeScanStats()
        _check(lib().m3dreg_node_register_single_scan(self._h, _p(np.ascontiguousarray(cloud)), C.c_int(len(cloud)),
                                                      _p(np.ascontiguousarray(mtf, dtype=np.float32).reshape(16)), iso_time.encode(), C.byref(st)),
               "m3dreg_node_register_single_scan")
        return st

    def __len__(self):
        return int(lib().m3dreg_node_scan_count(self._h))

    def pose(self, i: int):
        """-> (registered, tf) row-major 4x4"""
        r = np.zeros(16, dtype=np.float32); t = np.zeros(16, dtype=np.float32)
        _check(lib().m3dreg_node_get_pose(self._h, C.c_int(i), _p(r), _p(t)), "m3dreg_node_get_pose")
        return r.reshape(4, 4), t.reshape(4, 4)

    def scan(self, i: int) -> np.ndarray:
        n = _check(lib().m3dreg_node_scan_size(self._h, C.c_int(i)), "m3dreg_node_scan_size", allow=range(1, 1 << 31))
        out = np.zeros(n, dtype=POINT_DTYPE)
        _check(lib().m3dreg_node_get_scan(self._h, C.c_int(i), _p(out), C.c_int(n)), "m3dreg_node_get_scan")
        return out

    def scan_id(self, i: int) -> str:
        buf = C.create_string_buffer(512)
        lib().m3dreg_node_scan_id(self._h, C.c_int(i), buf, C.c_int(512))
        return buf.value.decode()

    def metascan(self) -> np.ndarray:
        n = C.c_int(0)
        _check(lib().m3dreg_node_metascan(self._h, None, C.c_int(0), C.byref(n)), "m3dreg_node_metascan")
        out = np.zeros(n.value, dtype=POINT_DTYPE)
        if n.value:
            _check(lib().m3dreg_node_metascan(self._h, _p(out), C.c_int(n.value), C.byref(n)), "m3dreg_node_metascan")
        return out

    def register_all(self) -> int:
        solved = C.c_int(0)
        _check(lib().m3dreg_node_register_all(self._h, C.byref(solved)), "m3dreg_node_register_all")
        return int(solved.value)

    def load_map(self, xml_path: str):
        _check(lib().m3dreg_node_load_map(self._h, str(xml_path).encode()), "m3dreg_node_load_map")

    def set_initial_pose(self, pose: np.ndarray):
        _check(lib().m3dreg_node_set_initial_pose(self._h, _p(np.ascontiguousarray(pose, dtype=np.float32).reshape(16))), "m3dreg_node_set_initial_pose")
