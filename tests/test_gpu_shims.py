"""The three exact-signature entry points of the reference (Calculate3Dpoint, CudaComputeHref,
g2o::CudaComputeH), driven the way NID_pose_estimation.cpp:232-276 and the LM loop
(optimization_algorithm_levenberg.cpp:78-115) drive them, checked against the oracle."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    return a.ctypes.data_as(_dp)


def test_reference_call_sequence(nid, orc, make_pair):
    p = make_pair(1000, 240, 320)
    L = nid.lib()
    rows, cols, cell, bins = p.rows, p.cols, 4, 16
    N = rows * cols
    depth = p.depth0.reshape(-1).copy()
    Twc0 = p.T_wc0.copy()
    intr = p.intr.copy()
    points = np.zeros(3 * N)  # the reference uses managed memory; any host pointer works here
    im0 = p.im0.astype(np.float64).reshape(-1).copy()
    im0[:7] = 255.0           # exercise the in-place clamp (CudaComputeHref.cu:102-105)
    im0_before = im0.copy()
    im1 = p.im1.astype(np.float64).reshape(-1).copy()
    pose0 = orc.reference_perturbation(p.T_wc1)
    M0 = orc.se3_to_mat16(pose0)

    L.nid_shim_Calculate3Dpoint(_d(depth), _d(Twc0), _d(points), _d(intr), rows, cols)
    P = orc.Problem((im0_before.reshape(rows, cols)).astype(np.uint8), p.depth0, p.im1, p.T_wc0, p.intr, cell, bins)
    P.set_quirks(0, 1)
    exp_pts = P.points3d()
    assert np.array_equal(points, exp_pts, equal_nan=True)

    bs_value = np.zeros(4 * N)
    bs_index = np.zeros(N, dtype=np.int32)
    bs_counter = np.zeros(cell * cell, dtype=np.int32)
    Href = np.zeros(cell * cell)
    L.nid_shim_CudaComputeHref(_d(im0), _d(points), _d(M0), _d(intr), bins, 3, cell, rows, cols, _d(bs_value),
                               bs_index.ctypes.data_as(_ip), bs_counter.ctypes.data_as(_ip), _d(Href))
    nco, hrefo = P.prepare(pose0)
    assert np.array_equal(bs_counter, nco)
    np.testing.assert_allclose(Href, hrefo, rtol=1e-12)
    bvo, bio = P.ref_weights()
    inb = ~np.isnan(bs_value[0::4])
    assert inb.sum() == nco.sum()
    np.testing.assert_allclose(bs_value.reshape(N, 4)[inb], bvo.reshape(N, 4)[inb], rtol=1e-13, atol=1e-16)
    assert np.array_equal(bs_index[inb], bio[inb])
    # in-place clamp of the reference image where it was used
    used255 = inb & (im0_before >= 255)
    assert np.all(im0[used255] == 254.999)

    for calc_der in (1, 0):
        Ht = np.zeros(cell * cell)
        Hj = np.zeros(cell * cell)
        der = np.full(6 * cell * cell, -7.0)
        L.nid_shim_CudaComputeH(calc_der, _d(im0), _d(im1), _d(points), bs_counter.ctypes.data_as(_ip), _d(bs_value),
                                bs_index.ctypes.data_as(_ip), _d(M0), _d(intr), bins, 3, cell, rows, cols, _d(Href),
                                None, None, _d(Ht), _d(Hj), _d(der))
        Hto, Hjo, erro, Jo = P.eval(pose0, True)
        np.testing.assert_allclose(Ht, Hto, rtol=1e-11)
        np.testing.assert_allclose(Hj, Hjo, rtol=1e-11)
        if calc_der:
            scale = np.abs(Jo).max(axis=1, keepdims=True)
            assert np.max(np.abs(der.reshape(-1, 6) - Jo) / scale) < 1e-8
        else:
            assert np.all(der == -7.0)  # der must not be touched (sparse_optimizer.cpp:413,423)
    L.nid_shim_reset()
