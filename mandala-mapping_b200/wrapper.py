"""Python mirror of the reference's L1 call surface, ``class CCudaWrapper``
(gpu_6dslam/gpu_6dslam/include/cudaWrapper.h:26-105, src/cudaWrapper.cpp), forwarding to the C ABI.

Same method names, argument order/meaning and error behaviour as the reference so the parity tests read like
calls the reference's own host code (src/gpu6DSLAM.cpp:313,405-406) would make.  The C++ equivalent for the
real m3d pipeline is ``include/cuda_wrapper_shim.hpp`` (repository root).
"""
from __future__ import annotations

import numpy as np


class Observations:
    """observations_t (include/cudaWrapper.h:14-24)."""

    def __init__(self, vobs_nn=None, om=0.0, fi=0.0, ka=0.0, tx=0.0, ty=0.0, tz=0.0):
        from . import OBS_DTYPE
        self.vobs_nn = np.zeros(0, dtype=OBS_DTYPE) if vobs_nn is None else vobs_nn
        self.m_pose = np.eye(4, dtype=np.float32)
        self.om, self.fi, self.ka = float(om), float(fi), float(ka)
        self.tx, self.ty, self.tz = float(tx), float(ty), float(tz)


class CudaError(RuntimeError):
    """Stands in for thrust::system_error thrown by throw_on_cuda_error (src/cudaWrapper.cpp:650-660)."""

    def __init__(self, code, where):
        self.code = code
        super().__init__(f"{where}: CUDA/m3dreg status {code}")


class CCudaWrapper:
    def __init__(self):
        self._ctx = None
        self.cuda_device = 0
        self.threads = 0
        self.threadsNV = 0

    # -- src/cudaWrapper.cpp:36-44 ------------------------------------------------------------------------
    def warmUpGPU(self, cudaDevice: int = 0):
        from . import Context
        if self._ctx is None or self.cuda_device != cudaDevice:
            if self._ctx is not None:
                self._ctx.close()
            self._ctx = Context(cudaDevice)
            self.cuda_device = cudaDevice
        self._ctx.warm_up()
        self.threads, self.threadsNV = 1024, 256   # cudaWrapper.cpp:71-91 (kept for API compatibility; unused)

    def getNumberOfAvailableThreads(self, cudaDevice: int = 0):
        return 1024

    @property
    def context(self):
        if self._ctx is None:
            self.warmUpGPU(self.cuda_device)
        return self._ctx

    # -- src/cudaWrapper.cpp:344-424 ----------------------------------------------------------------------
    def semanticNearestNeighbourhoodSearch(self, first_point_cloud, second_point_cloud, search_radius, bucket_size,
                                           bounding_box_extension, max_number_considered_in_INNER_bucket,
                                           max_number_considered_in_OUTER_bucket, nearest_neighbour_indexes):
        """Fills ``nearest_neighbour_indexes`` (int32 array, one entry per point of the second cloud) in place."""
        if len(nearest_neighbour_indexes) != len(second_point_cloud):
            return   # the reference silently returns (cudaWrapper.cpp:354)
        from . import M3dRegError
        try:
            self.context.semantic_nn_host(first_point_cloud, second_point_cloud, search_radius, bucket_size,
                                          bounding_box_extension, int(max_number_considered_in_INNER_bucket),
                                          int(max_number_considered_in_OUTER_bucket), nn_out=nearest_neighbour_indexes)
        except M3dRegError as e:
            self.throw_on_cuda_error(e.status, __file__, 0)

    # -- src/cudaWrapper.cpp:470-514 ----------------------------------------------------------------------
    @staticmethod
    def Matrix4ToEuler(m):
        from . import matrix4_to_euler
        return matrix4_to_euler(m)

    @staticmethod
    def EulerToMatrix(omfika, xyz):
        from . import euler_to_matrix
        return euler_to_matrix(omfika, xyz)

    # -- src/cudaWrapper.cpp:516-648 ----------------------------------------------------------------------
    def _register(self, obs: Observations, dof: int) -> bool:
        from . import M3dRegError, E_NOT_SPD
        try:
            st, pose6, _ = self.context.register_ls_host(obs.vobs_nn, [obs.tx, obs.ty, obs.tz, obs.om, obs.fi, obs.ka], dof)
        except M3dRegError as e:
            self.throw_on_cuda_error(e.status, __file__, 0)
        if st == E_NOT_SPD:
            print("problem with solving Ax=B")
            return False
        obs.tx, obs.ty, obs.tz, obs.om, obs.fi, obs.ka = (float(v) for v in pose6)
        return True

    def registerLS(self, obs: Observations) -> bool:
        return self._register(obs, 6)

    def registerLS_4DOF(self, obs: Observations) -> bool:
        return self._register(obs, 4)

    # -- src/cudaWrapper.cpp:650-660 ----------------------------------------------------------------------
    @staticmethod
    def throw_on_cuda_error(code, file, line):
        if code != 0:
            raise CudaError(code, f"{file}({line})")
