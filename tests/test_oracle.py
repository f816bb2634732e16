"""Known-answer and self-consistency tests of the CPU oracle. The reference ships
no tests for this path; these pins plus tests/golden/ (outputs of the reference's own CUDA code) are
what the oracle stands on."""
import numpy as np
import pytest


@pytest.mark.parametrize("bins", [6, 8, 10, 12, 14, 16, 32])
def test_bspline_partition_of_unity(orc, bins):
    for I in np.concatenate([np.linspace(0, 254.999, 301), [0, 51, 102, 254.999]]):
        u = I * (bins - 3) / 255.0
        k = int(np.floor(u))
        w = [orc.bspline(k + m, 4, u, bins) for m in range(4)]
        dw = [orc.bspline_der(k + m, 4, u, bins) for m in range(4)]
        assert abs(sum(w) - 1) < 4e-16 * 4
        assert abs(sum(dw)) < 2e-15
        assert min(w) >= 0


def test_bspline_known_answers(orc):
    # SURVEY A-6, B = 8
    w = [orc.bspline(1 + m, 4, 1.0, 8) for m in range(4)]
    dw = [orc.bspline_der(1 + m, 4, 1.0, 8) for m in range(4)]
    np.testing.assert_allclose(w, [1 / 4, 7 / 12, 1 / 6, 0], atol=1e-15)
    np.testing.assert_allclose(dw, [-3 / 4, 1 / 4, 1 / 2, 0], atol=1e-15)
    w = [orc.bspline(2 + m, 4, 2.0, 8) for m in range(4)]
    dw = [orc.bspline_der(2 + m, 4, 2.0, 8) for m in range(4)]
    np.testing.assert_allclose(w, [1 / 6, 2 / 3, 1 / 6, 0], atol=1e-15)
    np.testing.assert_allclose(dw, [-1 / 2, 0, 1 / 2, 0], atol=1e-15)
    # the reference's quirk at u == 0: derivative reported as 0 (true value [-3, 3, 0, 0])
    assert [orc.bspline(m, 4, 0.0, 8) for m in range(4)] == [1, 0, 0, 0]
    assert [orc.bspline_der(m, 4, 0.0, 8) for m in range(4)] == [0, 0, 0, 0]


def test_bspline_derivative_matches_finite_difference(orc):
    for bins in (8, 16):
        for u in (0.3, 1.7, 2.2, bins - 3 - 0.4):
            k = int(np.floor(u))
            h = 1e-6
            for m in range(4):
                fd = (orc.bspline(k + m, 4, u + h, bins) - orc.bspline(k + m, 4, u - h, bins)) / (2 * h)
                assert abs(fd - orc.bspline_der(k + m, 4, u, bins)) < 1e-8


def test_huber_boundary(orc):
    d = np.sqrt(0.95)
    dsqr = float(np.float32(d * d))  # robust_kernel_impl.h:84 declares `float dsqr`
    r = orc.huber(dsqr, d)
    assert r[0] == dsqr and r[1] == 1.0
    e = 0.96
    r = orc.huber(e, d)
    assert r[0] == pytest.approx(2 * np.sqrt(e) * d - dsqr, rel=1e-15)
    assert r[1] == pytest.approx(d / np.sqrt(e), rel=1e-15)


def test_se3_exp_small_angle_branch(orc):
    # theta < 1e-5 : R = I + Omega + Omega^2 ; V = R (se3quat.h:237-243)
    xi = np.array([3e-6, -2e-6, 1e-6, 0.1, -0.2, 0.3])
    p = orc.se3_exp(xi)
    M = orc.se3_to_mat16(p).reshape(4, 4).T
    Om = np.array([[0, -xi[2], xi[1]], [xi[2], 0, -xi[0]], [-xi[1], xi[0], 0]])
    R = np.eye(3) + Om + Om @ Om
    np.testing.assert_allclose(M[:3, 3], R @ xi[3:], atol=1e-15)
    np.testing.assert_allclose(M[:3, :3], R, atol=1e-10)  # quaternion normalisation re-orthogonalises


def test_se3_exp_matches_matrix_exponential(orc):
    from scipy.linalg import expm
    rng = np.random.default_rng(0)
    for _ in range(5):
        xi = rng.normal(size=6) * 0.3
        Om = np.array([[0, -xi[2], xi[1]], [xi[2], 0, -xi[0]], [-xi[1], xi[0], 0]])
        A = np.zeros((4, 4))
        A[:3, :3] = Om
        A[:3, 3] = xi[3:]
        M = orc.se3_to_mat16(orc.se3_exp(xi)).reshape(4, 4).T
        np.testing.assert_allclose(M, expm(A), atol=1e-12)


def test_se3_group_ops(orc):
    rng = np.random.default_rng(1)
    a = orc.se3_exp(rng.normal(size=6) * 0.2)
    b = orc.se3_exp(rng.normal(size=6) * 0.2)
    Ma = orc.se3_to_mat16(a).reshape(4, 4).T
    Mb = orc.se3_to_mat16(b).reshape(4, 4).T
    np.testing.assert_allclose(orc.se3_to_mat16(orc.se3_mul(a, b)).reshape(4, 4).T, Ma @ Mb, atol=1e-14)
    np.testing.assert_allclose(orc.se3_to_mat16(orc.se3_inverse(a)).reshape(4, 4).T, np.linalg.inv(Ma), atol=1e-14)
    p = rng.normal(size=3)
    np.testing.assert_allclose(orc.se3_map(a, p), Ma[:3, :3] @ p + Ma[:3, 3], atol=1e-14)
    assert orc.se3_mul(a, b)[6] >= 0  # normalizeRotation keeps w >= 0


def test_ldlt6(orc):
    rng = np.random.default_rng(2)
    A = rng.normal(size=(6, 6))
    H = A @ A.T + 1e-3 * np.eye(6)
    b = rng.normal(size=6)
    ok, x = orc.ldlt6_solve(H, b)
    assert ok == 1
    np.testing.assert_allclose(x, np.linalg.solve(H, b), rtol=1e-9)
    H2 = H.copy()
    H2[2, 2] = -5.0
    ok, _ = orc.ldlt6_solve(H2, b)
    assert ok == 0  # not positive: linear_solver_dense.h:106 returns false


def test_reference_perturbation(orc, make_pair):
    p = make_pair()
    pose = orc.reference_perturbation(p.T_wc1)
    M = orc.se3_to_mat16(pose).reshape(4, 4).T
    Twc = p.T_wc1.reshape(4, 4).T
    Tcw = np.linalg.inv(Twc)
    a = 0.005 * np.pi
    c, s = np.cos(a), np.sin(a)
    Rx = np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    Ry = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    Rz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    np.testing.assert_allclose(M[:3, :3], Rx @ Ry @ Rz @ Tcw[:3, :3], atol=1e-12)
    np.testing.assert_allclose(M[:3, 3], Tcw[:3, 3] + [0.01, -0.02, -0.02], atol=1e-12)


def test_hard_binned_nid_of_an_image_with_itself_is_zero(orc, make_pair, synth):
    # NID_standard_property.cpp:473-481: identical images at the identity pose -> every cell's NID is 0
    # up to rounding of (2Hj - Hr - Hc)/Hj with Hj == Hr == Hc
    p = make_pair()
    Tcw0 = synth.mat16_inverse(p.T_wc0)
    total, cells = orc.hard_nid(p.im0, p.depth0, p.im0, p.T_wc0, Tcw0, p.intr, 4, 8)
    assert np.all(np.abs(cells) < 1e-2)  # depth quantisation moves a few samples across bin borders
    # with exact (unquantised) geometry the joint histogram is diagonal: use a constant-depth plane
    depth = np.full_like(p.depth0, 2.0)
    I4 = np.eye(4).T.reshape(16)
    total, cells = orc.hard_nid(p.im0, depth, p.im0, I4, I4, p.intr, 4, 8)
    assert total < 1e-12


def test_oracle_jacobian_against_finite_differences_interior_cells(orc, make_pair):
    """Sanity only (SURVEY A-6): the analytic Jacobian ignores samples entering/leaving the image, so it
    is compared on interior cells of a smooth scene."""
    p = make_pair(1000, 240, 320)
    pose0 = orc.reference_perturbation(p.T_wc1)
    P = orc.Problem(p.im0, p.depth0, p.im1, p.T_wc0, p.intr, 4, 16)
    P.prepare(pose0)
    _, _, _, J = P.eval(pose0, True)
    h = 1e-5
    Jfd = np.zeros_like(J)
    for a in range(6):
        d = np.zeros(6)
        d[a] = h
        ep = P.eval(orc.se3_mul(orc.se3_exp(d), pose0), False)[2]
        em = P.eval(orc.se3_mul(orc.se3_exp(-d), pose0), False)[2]
        Jfd[:, a] = (ep - em) / (2 * h)
    for c in (5, 6, 9, 10):
        scale = np.max(np.abs(Jfd[c]))
        assert np.max(np.abs(J[c] - Jfd[c])) < 0.08 * scale


def test_inactive_cells_and_counts(orc, make_pair):
    p = make_pair(1001, 120, 160, invalid_depth_frac=0.2)
    pose0 = orc.reference_perturbation(p.T_wc1)
    P = orc.Problem(p.im0, p.depth0, p.im1, p.T_wc0, p.intr, 8, 10)  # 15x20 = 300 px per cell, 20% invalid
    nc, href = P.prepare(pose0)
    assert np.all(nc < 300) and np.all(np.isnan(href))
    Ht, Hj, err, J = P.eval(pose0, True)
    assert np.all(np.isnan(Ht)) and np.all(np.isnan(J))


def test_lm_decreases_cost(orc, make_pair):
    p = make_pair(1000, 240, 320)
    pose0 = orc.reference_perturbation(p.T_wc1)
    P = orc.Problem(p.im0, p.depth0, p.im1, p.T_wc0, p.intr, 4, 16, threads=4)
    P.prepare(pose0)
    chi0, H, b = P.gn_system(pose0, np.sqrt(0.95))
    pose, its, trace, counts = P.optimize(pose0, 10)
    assert 1 <= its <= 10 and counts[0] == its
    assert trace[-1, 0] < chi0
    assert np.all(np.diff(trace[:, 0]) <= 1e-12)  # accepted steps only ever lower the robust cost


@pytest.mark.parametrize("bins", [6, 7, 8, 10, 12, 14, 16, 32, 40])
def test_uniform_basis_fold_equals_clamped_basis(orc, bins):
    """the sorted kernels evaluate the *uniform* cubic B-spline per pixel and fold a 3x3 block at either
    end once per task row (values) / class table (derivatives). Check the identity N = A U and N' = A U' against the
    oracle's Cox-de Boor recursion (types_six_dof_expmap.cpp:738-800) over the whole domain, for every supported B."""
    NS = bins - 3
    A = np.eye(bins)
    A[0, 0] = 6.0
    A[1, 0], A[1, 1] = -6.0, 1.5
    A[2, 0], A[2, 1], A[2, 2] = 1.0, -0.5, 1.0
    A[bins - 1, bins - 1] = 6.0
    A[bins - 2, bins - 1], A[bins - 2, bins - 2] = -6.0, 1.5
    A[bins - 3, bins - 1], A[bins - 3, bins - 2], A[bins - 3, bins - 3] = 1.0, -0.5, 1.0
    rng = np.random.default_rng(7)
    us = np.concatenate([rng.uniform(0, NS, 400), rng.uniform(0, 2, 100), rng.uniform(NS - 2, NS, 100),
                         [1e-9, 0.5, 1.0 - 1e-12, NS - 1e-9, 2.0 + 1e-7]])
    for u in us:
        k = min(int(u), NS - 1)
        f = u - k
        U = np.zeros(bins)
        dU = np.zeros(bins)
        U[k:k + 4] = [(1 - f) ** 3 / 6, (3 * f ** 3 - 6 * f ** 2 + 4) / 6, (-3 * f ** 3 + 3 * f ** 2 + 3 * f + 1) / 6, f ** 3 / 6]
        dU[k:k + 4] = [-(1 - f) ** 2 / 2, (3 * f ** 2 - 4 * f) / 2, (-3 * f ** 2 + 2 * f + 1) / 2, f ** 2 / 2]
        N = np.array([orc.bspline(j, 4, u, bins) for j in range(bins)])
        dN = np.array([orc.bspline_der(j, 4, u, bins) for j in range(bins)])
        np.testing.assert_allclose(A @ U, N, rtol=0, atol=2e-14)
        np.testing.assert_allclose(A @ dU, dN, rtol=0, atol=2e-13)
