"""ctypes binding of the CPU oracle (oracle/libnid_oracle.so). TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; the product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libnid_oracle.so")
    src = os.path.join(_HERE, "nid_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libnid_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_bspline.restype = C.c_double
        L.orc_bspline.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int]
        L.orc_bspline_der.restype = C.c_double
        L.orc_bspline_der.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int]
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [_u8p, _dp, _u8p, C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_quirks.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_points3d.argtypes = [C.c_void_p, _dp]
        L.orc_cell_points.restype = C.c_int
        L.orc_cell_points.argtypes = [C.c_void_p, C.c_int]
        L.orc_prepare.argtypes = [C.c_void_p, _dp, _ip, _dp]
        L.orc_ref_weights.argtypes = [C.c_void_p, _dp, _ip]
        L.orc_eval.argtypes = [C.c_void_p, _dp, C.c_int, _dp, _dp, _dp, _dp]
        L.orc_last_hist.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.orc_last_dhist.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.orc_pixels.argtypes = [C.c_void_p, _dp, _dp]
        L.orc_optimize.restype = C.c_int
        L.orc_optimize.argtypes = [C.c_void_p, _dp, C.c_int, C.c_double, _dp, _ip]
        L.orc_gn_system.argtypes = [C.c_void_p, _dp, C.c_double, _dp, _dp, _dp]
        L.orc_hard_nid.restype = C.c_double
        L.orc_hard_nid.argtypes = [_u8p, _dp, _u8p, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, C.c_int, _dp, C.c_int]
        L.orc_huber.argtypes = [C.c_double, C.c_double, _dp]
        L.orc_ldlt6_solve.restype = C.c_int
        L.orc_ldlt6_solve.argtypes = [_dp, _dp, _dp]
        for name, n in (("orc_se3_exp", 2), ("orc_se3_inverse", 2), ("orc_se3_to_mat16", 2),
                        ("orc_reference_perturbation", 2), ("orc_se3_mul", 3), ("orc_se3_from_Rt", 3),
                        ("orc_se3_map", 3)):
            getattr(L, name).argtypes = [_dp] * n
        _LIB = L
    return _LIB


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None:
        assert a.size == n, (a.size, n)
    return a


def bspline(index, order, u, bins):
    return lib().orc_bspline(index, order, float(u), bins)


def bspline_der(index, order, u, bins):
    return lib().orc_bspline_der(index, order, float(u), bins)


def se3_exp(upd6):
    out = np.zeros(7)
    lib().orc_se3_exp(_d(_f64(upd6, 6)), _d(out))
    return out


def se3_mul(a7, b7):
    out = np.zeros(7)
    lib().orc_se3_mul(_d(_f64(a7, 7)), _d(_f64(b7, 7)), _d(out))
    return out


def se3_inverse(a7):
    out = np.zeros(7)
    lib().orc_se3_inverse(_d(_f64(a7, 7)), _d(out))
    return out


def se3_to_mat16(a7):
    out = np.zeros(16)
    lib().orc_se3_to_mat16(_d(_f64(a7, 7)), _d(out))
    return out


def se3_from_Rt(R, t):
    out = np.zeros(7)
    lib().orc_se3_from_Rt(_d(_f64(R, 9)), _d(_f64(t, 3)), _d(out))
    return out


def se3_from_mat16(m16):
    M = np.asarray(m16, dtype=np.float64).reshape(4, 4).T
    return se3_from_Rt(M[:3, :3].reshape(9), M[:3, 3])


def se3_map(a7, p):
    out = np.zeros(3)
    lib().orc_se3_map(_d(_f64(a7, 7)), _d(_f64(p, 3)), _d(out))
    return out


def reference_perturbation(T_wc1_mat16):
    out = np.zeros(7)
    lib().orc_reference_perturbation(_d(_f64(T_wc1_mat16, 16)), _d(out))
    return out


def huber(chi2, delta):
    out = np.zeros(3)
    lib().orc_huber(float(chi2), float(delta), _d(out))
    return out


def ldlt6_solve(H, b):
    x = np.zeros(6)
    ok = lib().orc_ldlt6_solve(_d(_f64(H, 36)), _d(_f64(b, 6)), _d(x))
    return ok, x


class Problem:
    """One frame pair with cell x cell unary NID edges (CPU mode of NID_pose_estimation)."""

    def __init__(self, im0, depth, im1, T_wc0, intr, cell, bins, threads=1):
        self.rows, self.cols = im0.shape
        self.cell, self.bins = cell, bins
        self.n = self.rows * self.cols
        self._im0 = np.ascontiguousarray(im0, dtype=np.uint8)
        self._im1 = np.ascontiguousarray(im1, dtype=np.uint8)
        self._depth = _f64(depth, self.n)
        self._h = lib().orc_create(self._im0.ctypes.data_as(_u8p), _d(self._depth), self._im1.ctypes.data_as(_u8p),
                                   self.rows, self.cols, _d(_f64(T_wc0, 16)), _d(_f64(np.asarray(intr)[:4], 4)),
                                   cell, bins, threads)

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_quirks(self, jac_bound_gpu=0, warp_with_matrix=0):
        lib().orc_set_quirks(self._h, jac_bound_gpu, warp_with_matrix)

    def points3d(self):
        out = np.zeros(3 * self.n)
        lib().orc_points3d(self._h, _d(out))
        return out

    def prepare(self, pose7):
        nc = np.zeros(self.cell * self.cell, dtype=np.int32)
        href = np.zeros(self.cell * self.cell)
        lib().orc_prepare(self._h, _d(_f64(pose7, 7)), nc.ctypes.data_as(_ip), _d(href))
        return nc, href

    def ref_weights(self):
        bv = np.zeros(4 * self.n)
        bi = np.zeros(self.n, dtype=np.int32)
        lib().orc_ref_weights(self._h, _d(bv), bi.ctypes.data_as(_ip))
        return bv, bi

    def eval(self, pose7, want_jac=True):
        c2 = self.cell * self.cell
        Ht, Hj, err, J = np.zeros(c2), np.zeros(c2), np.zeros(c2), np.full(6 * c2, np.nan)
        lib().orc_eval(self._h, _d(_f64(pose7, 7)), int(want_jac), _d(Ht), _d(Hj), _d(err), _d(J))
        return Ht, Hj, err, J.reshape(c2, 6)

    def last_hist(self, c):
        pt = np.zeros(self.bins)
        pj = np.zeros(self.bins * self.bins)
        lib().orc_last_hist(self._h, c, _d(pt), _d(pj))
        return pt, pj.reshape(self.bins, self.bins)

    def last_dhist(self, c):
        dpt = np.zeros(self.bins * 6)
        dpj = np.zeros(self.bins * self.bins * 6)
        lib().orc_last_dhist(self._h, c, _d(dpt), _d(dpj))
        return dpt.reshape(self.bins, 6), dpj.reshape(self.bins, self.bins, 6)

    def pixels(self, pose7):
        out = np.zeros(8 * self.n)
        lib().orc_pixels(self._h, _d(_f64(pose7, 7)), _d(out))
        return out.reshape(self.n, 8)

    def gn_system(self, pose7, delta):
        chi2 = C.c_double(0)
        H = np.zeros(36)
        b = np.zeros(6)
        lib().orc_gn_system(self._h, _d(_f64(pose7, 7)), float(delta), C.byref(chi2), _d(H), _d(b))
        return chi2.value, H.reshape(6, 6), b

    def optimize(self, pose7, max_iters=10, delta=np.sqrt(0.95)):
        pose = _f64(pose7, 7).copy()
        trace = np.zeros(10 * max_iters)
        counts = np.zeros(2, dtype=np.int32)
        its = lib().orc_optimize(self._h, _d(pose), max_iters, float(delta), _d(trace), counts.ctypes.data_as(_ip))
        return pose, its, trace.reshape(max_iters, 10)[:its], counts


def hard_nid(im0, depth, im1, T_wc0, T_cw1_mat16, intr, cell, bins, threads=1):
    rows, cols = im0.shape
    im0 = np.ascontiguousarray(im0, dtype=np.uint8)
    im1 = np.ascontiguousarray(im1, dtype=np.uint8)
    cells = np.zeros(cell * cell)
    total = lib().orc_hard_nid(im0.ctypes.data_as(_u8p), _d(_f64(depth, rows * cols)), im1.ctypes.data_as(_u8p),
                               rows, cols, _d(_f64(T_wc0, 16)), _d(_f64(T_cw1_mat16, 16)),
                               _d(_f64(np.asarray(intr)[:4], 4)), cell, bins, _d(cells), threads)
    return total, cells
