"""Runs the reference's own CUDA implementation (oracle/_ref/libnid_ref_gpu.so, compiled unmodified from
CudaPoints3d.cu and g2o/g2o/core/computeH.cu of the upstream tree by oracle/Makefile, plus CudaComputeHref.cu with
the two-line cudaMemset fix of SURVEY appendix B-2 applied at build time). TEST
INFRASTRUCTURE ONLY — needs a GPU; used to pin the CPU oracle and to time the "old GPU path".

The reference kernels have no `i < rows*cols` guard: they touch
thread_work * (SMs*2048/1024) * 1024 pixel slots. On a 148-SM B200 that is a multiple of 303104, so
images of 592 x 512 = 303104 pixels make the launch exact and the unmodified code memory-safe.
Its knot tables exist for 6/8/10/12/14 bins only (computeH.cu:99-134).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libnid_ref_gpu.so")
SAFE_ROWS, SAFE_COLS = 512, 592

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_L = None


def available() -> bool:
    return os.path.exists(SO)


def lib():
    global _L
    if _L is None:
        L = C.CDLL(SO)
        L.ref_managed_alloc.restype = C.c_void_p
        L.ref_managed_alloc.argtypes = [C.c_size_t]
        L.ref_managed_free.argtypes = [C.c_void_p]
        L.ref_Calculate3Dpoint.argtypes = [_dp, _dp, _dp, _dp, C.c_int, C.c_int]
        L.ref_CudaComputeH.argtypes = [C.c_int, _dp, _dp, _dp, _ip, _dp, _ip, _dp, _dp] + [C.c_int] * 5 + [_dp] * 4
        if hasattr(L, "ref_CudaComputeHref"):
            L.ref_CudaComputeHref.argtypes = [_dp, _dp, _dp, _dp] + [C.c_int] * 5 + [_dp, _ip, _ip, _dp]
        _L = L
    return _L


class Managed:
    """cudaMallocManaged buffer viewed as a numpy array (NID_pose_estimation.cpp:240-242)."""

    def __init__(self, n):
        self.ptr = lib().ref_managed_alloc(n * 8)
        if not self.ptr:
            raise RuntimeError("cudaMallocManaged failed")
        self.a = np.ctypeslib.as_array((C.c_double * n).from_address(self.ptr))

    def p(self):
        return C.cast(self.ptr, _dp)

    def free(self):
        if self.ptr:
            lib().ref_managed_free(self.ptr)
            self.ptr = None


def _d(a):
    return a.ctypes.data_as(_dp)


class RefGpu:
    """The reference's GPU hot path for one pair: Calculate3Dpoint once, then CudaComputeH per pose. The
    prepare arrays (bs_counter, bs_value, bs_index, Href) are supplied by the caller (the CPU oracle),
    because the reference's CudaComputeHref never zeroes its accumulators (appendix B-2)."""

    def __init__(self, pair, cell, bins):
        self.rows, self.cols, self.cell, self.bins = pair.rows, pair.cols, cell, bins
        n = self.rows * self.cols
        self.n = n
        self.im0 = Managed(n)
        self.im1 = Managed(n)
        self.pts = Managed(3 * n)
        self.im0.a[:] = pair.im0.reshape(-1)
        self.im1.a[:] = pair.im1.reshape(-1)
        self.intr = np.ascontiguousarray(pair.intr, dtype=np.float64)
        depth = np.ascontiguousarray(pair.depth0.reshape(-1), dtype=np.float64)
        twc0 = np.ascontiguousarray(pair.T_wc0, dtype=np.float64)
        lib().ref_Calculate3Dpoint(_d(depth), _d(twc0), self.pts.p(), _d(self.intr), self.rows, self.cols)

    def points3d(self):
        return np.array(self.pts.a)

    def compute_href(self, pose16):
        """The reference's CudaComputeHref (CudaComputeHref.cu:139-223) at the initial pose: returns
        (bs_counter, bs_value[4N], bs_index[N], Href) as the reference leaves them (NaN weights / index 0 for pixels
        without a sample, NaN Href for cells under 300 points); im0 is clamped in place like upstream does."""
        c2 = self.cell * self.cell
        bs_value = np.zeros(4 * self.n)
        bs_index = np.zeros(self.n, dtype=np.int32)
        bs_counter = np.zeros(c2, dtype=np.int32)
        Href = np.zeros(c2)  # `Href[i] -= ...` into the caller's zeros (NID_pose_estimation.cpp:238)
        pose = np.ascontiguousarray(pose16, dtype=np.float64)
        lib().ref_CudaComputeHref(self.im0.p(), self.pts.p(), _d(pose), _d(self.intr), self.bins, 3, self.cell, self.rows,
                                  self.cols, _d(bs_value), bs_index.ctypes.data_as(_ip), bs_counter.ctypes.data_as(_ip),
                                  _d(Href))
        return bs_counter, bs_value, bs_index, Href

    def set_prepare(self, bs_counter, bs_value, bs_index, Href):
        self.bs_counter = np.ascontiguousarray(bs_counter, dtype=np.int32)
        self.bs_value = np.ascontiguousarray(bs_value, dtype=np.float64)
        self.bs_index = np.ascontiguousarray(bs_index, dtype=np.int32)
        self.Href = np.ascontiguousarray(Href, dtype=np.float64)

    def compute_h(self, pose16, want_jac=True):
        c2 = self.cell * self.cell
        Ht, Hj = np.zeros(c2), np.zeros(c2)
        der = np.zeros(6 * c2)
        pose = np.ascontiguousarray(pose16, dtype=np.float64)
        lib().ref_CudaComputeH(int(want_jac), self.im0.p(), self.im1.p(), self.pts.p(),
                               self.bs_counter.ctypes.data_as(_ip), _d(self.bs_value),
                               self.bs_index.ctypes.data_as(_ip), _d(pose), _d(self.intr), self.bins, 3, self.cell,
                               self.rows, self.cols, _d(self.Href), _d(Ht), _d(Hj), _d(der))
        return Ht, Hj, der.reshape(c2, 6)

    def close(self):
        for m in (self.im0, self.im1, self.pts):
            m.free()
