"""The product's host-side pose algebra, Huber and 6x6 solve (nid-pose-estimation_b200/host/nid_host_math.hpp: what the LM
driver and the apps run on the host) against the oracle's independently written restatement of se3quat.h,
robust_kernel_impl.cpp and linear_solver_dense.h. Compiled here with g++ behind a tiny C wrapper; no GPU involved.
(The GPU tests compare whole LM trajectories; this pins the host pieces one by one in the CPU tier.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WRAPPER = r'''
#include "nid_host_math.hpp"
using namespace nidhost;
static Pose7 P(const double* v) { Pose7 p; for (int i = 0; i < 3; i++) p.t[i] = v[i]; for (int i = 0; i < 4; i++) p.q[i] = v[3 + i]; return p; }
static void O(const Pose7& p, double* v) { for (int i = 0; i < 3; i++) v[i] = p.t[i]; for (int i = 0; i < 4; i++) v[3 + i] = p.q[i]; }
extern "C" {
void hm_exp(const double* xi, double* out7) { O(pose_exp(xi), out7); }
void hm_mul(const double* a, const double* b, double* out7) { O(pose_mul(P(a), P(b)), out7); }
void hm_inverse(const double* a, double* out7) { O(pose_inverse(P(a)), out7); }
void hm_to_mat16(const double* a, double* m) { pose_to_mat16(P(a), m); }
void hm_from_mat16(const double* m, double* out7) { O(pose_from_mat16(m), out7); }
void hm_perturb(const double* Twc1, double* out7) { O(reference_perturbation(Twc1), out7); }
void hm_huber(double e2, double delta, double* rho3) { huber(e2, delta, rho3); }
int hm_ldlt(const double* H, const double* b, double* x) { return ldlt6_solve(H, b, x) ? 1 : 0; }
}
'''


@pytest.fixture(scope="module")
def hm(tmp_path_factory):
    d = tmp_path_factory.mktemp("hostmath")
    src = d / "wrap.cpp"
    src.write_text(WRAPPER)
    so = d / "libhm.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "nid-pose-estimation_b200", "host"),
                           "-o", str(so), str(src)])
    L = C.CDLL(str(so))
    L.hm_huber.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_double)]
    L.hm_ldlt.restype = C.c_int
    return L


def _p(a):
    return np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(C.POINTER(C.c_double))


def _call7(f, *args):
    out = np.zeros(7)
    f(*[_p(a) for a in args], _p(out))
    return out


def test_pose_algebra_equals_the_oracle(hm, orc, synth):
    rng = np.random.default_rng(11)
    for k in range(200):
        scale = [1e-9, 1e-4, 0.05, 1.0, 3.0][k % 5]  # tiny angles take the series branch of exp (se3quat.h:236-243)
        xi = rng.uniform(-1, 1, 6) * scale
        a = _call7(hm.hm_exp, xi)
        ao = orc.se3_exp(xi)
        np.testing.assert_allclose(a, ao, rtol=0, atol=1e-15)
        b = orc.se3_exp(rng.uniform(-1, 1, 6))
        np.testing.assert_allclose(_call7(hm.hm_mul, a, b), orc.se3_mul(ao, b), rtol=0, atol=1e-15)
        np.testing.assert_allclose(_call7(hm.hm_inverse, b), orc.se3_inverse(b), rtol=0, atol=1e-15)
        m = np.zeros(16)
        hm.hm_to_mat16(_p(b), _p(m))
        np.testing.assert_allclose(m, orc.se3_to_mat16(b), rtol=0, atol=1e-15)
        np.testing.assert_allclose(_call7(hm.hm_from_mat16, m), orc.se3_from_mat16(m), rtol=0, atol=1e-15)
    p = synth.make_pair(1000, 48, 64)
    np.testing.assert_allclose(_call7(hm.hm_perturb, p.T_wc1), orc.reference_perturbation(p.T_wc1), rtol=0, atol=1e-15)


def test_huber_and_ldlt_equal_the_oracle(hm, orc):
    rng = np.random.default_rng(12)
    delta = np.sqrt(0.95)
    dsqr = float(np.float32(delta * delta))  # robust_kernel_impl.h:84 keeps delta^2 as a float
    for e2 in [0.0, 1e-12, 0.5, dsqr, np.nextafter(dsqr, 2.0), 0.95, 0.9500001, 1.0, 7.3, 1e6]:
        rho = np.zeros(3)
        hm.hm_huber(float(e2), float(delta), _p(rho))
        np.testing.assert_array_equal(rho, np.asarray(orc.huber(e2, delta), dtype=np.float64))
    for k in range(100):
        A = rng.normal(size=(6, 6))
        H = A @ A.T + (1e-3 if k % 2 else 10.0) * np.eye(6)  # positive definite, well and badly conditioned
        if k % 10 == 9:
            H[2, 2] = -abs(H[2, 2])  # an indefinite system: the solve must refuse (linear_solver_dense.h:104-110)
        b = rng.normal(size=6)
        x = np.zeros(6)
        ok = hm.hm_ldlt(_p(H.reshape(-1)), _p(b), _p(x))
        oko, xo = orc.ldlt6_solve(H, b)
        assert bool(ok) == bool(oko)
        if ok:
            np.testing.assert_allclose(x, xo, rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(H @ x, b, rtol=1e-8, atol=1e-8)
