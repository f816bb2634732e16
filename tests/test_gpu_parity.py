"""GPU parity: every result of the CUDA path, obtained through the C-ABI, against the CPU oracle on the
same seeded inputs. Tolerances: north_star asks for 1e-5 relative on cost and gradient; the checks here
are much tighter because both sides are fp64 and only differ by summation order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DELTA = np.sqrt(0.95)


NATURAL, SORTED = 1, 2  # nid_set_option("path", ...): both kernel families are held to the same parity bar


def _setup(nid, orc, p, cell, bins, matrix_warp=True, path=0, **ctx_kw):
    pose0 = orc.reference_perturbation(p.T_wc1)
    P = orc.Problem(p.im0, p.depth0, p.im1, p.T_wc0, p.intr, cell, bins, threads=4)
    if matrix_warp:
        P.set_quirks(0, 1)  # warp with the 4x4 like the CUDA code does: bit-identical (u, v)
    ctx = nid.Context(p.rows, p.cols, cell, bins, **ctx_kw)
    ctx.set_option("path", path)
    ctx.set_option("keep_hist", 1)
    ctx.set_pair(0, p.depth0, p.im0, p.im1, p.T_wc0, p.intr)
    return P, ctx, pose0


def _jrel(J, Jo):
    """max over cells of |dJ| / max|J_cell|"""
    act = ~np.isnan(Jo[:, 0])
    assert np.array_equal(np.isnan(J[:, 0]), ~act)
    scale = np.max(np.abs(Jo[act]), axis=1, keepdims=True)
    return np.max(np.abs(J[act] - Jo[act]) / scale)


def test_points3d(nid, orc, make_pair):
    p = make_pair(1002, 120, 160, invalid_depth_frac=0.05)
    P, ctx, pose0 = _setup(nid, orc, p, 4, 16)
    got = ctx.points3d(0)
    exp = P.points3d()
    assert np.array_equal(np.isnan(got), np.isnan(exp))
    m = ~np.isnan(exp)
    assert np.array_equal(got[m], exp[m])  # same operation order, no contraction: bit-exact


@pytest.mark.parametrize("path", [NATURAL, SORTED])
@pytest.mark.parametrize("cell,bins,rows,cols", [(1, 8, 120, 160), (4, 16, 240, 320), (16, 10, 480, 640), (4, 32, 240, 320), (3, 12, 125, 170),
                                                 (2, 9, 120, 160), (2, 21, 120, 160), (2, 40, 240, 320)])
def test_prepare_eval_parity(nid, orc, make_pair, cell, bins, rows, cols, path):
    p = make_pair(1000, rows, cols)
    P, ctx, pose0 = _setup(nid, orc, p, cell, bins, path=path)
    M0 = orc.se3_to_mat16(pose0)
    nc, href = ctx.prepare(0, M0)
    nco, hrefo = P.prepare(pose0)
    assert np.array_equal(nc, nco)
    act = ~np.isnan(hrefo)
    assert np.array_equal(np.isnan(href), ~act)
    np.testing.assert_allclose(href[act], hrefo[act], rtol=1e-12)
    # evaluate at the prepare pose and at a different one
    for pose in (pose0, orc.se3_mul(orc.se3_exp(np.array([0.002, -0.001, 0.0015, 0.004, -0.003, 0.002])), pose0)):
        M = orc.se3_to_mat16(pose)
        Ht, Hj, J = ctx.eval(0, M, True)
        Hto, Hjo, erro, Jo = P.eval(pose, True)
        np.testing.assert_allclose(Ht[act], Hto[act], rtol=1e-11)
        np.testing.assert_allclose(Hj[act], Hjo[act], rtol=1e-11)
        assert np.all(np.isnan(Ht[~act])) and np.all(np.isnan(Hj[~act]))
        assert _jrel(J, Jo) < 1e-8
        err = (2 * Hj - href - Ht) / Hj
        np.testing.assert_allclose(err[act], erro[act], rtol=1e-10)
        # cost-only flavour gives the same entropies and leaves der alone
        Ht2, Hj2, J2 = ctx.eval(0, M, False)
        if path == SORTED:  # fixed summation order: bit-identical
            assert np.array_equal(Ht2[act], Ht[act]) and np.array_equal(Hj2[act], Hj[act])
        np.testing.assert_allclose(Ht2[act], Ht[act], rtol=1e-13)
        np.testing.assert_allclose(Hj2[act], Hj[act], rtol=1e-13)
        assert np.all(np.isnan(J2))


@pytest.mark.parametrize("path", [NATURAL, SORTED])
def test_histograms(nid, orc, make_pair, path):
    p = make_pair(1000, 240, 320)
    P, ctx, pose0 = _setup(nid, orc, p, 4, 16, path=path)
    M0 = orc.se3_to_mat16(pose0)
    ctx.prepare(0, M0)
    P.prepare(pose0)
    ctx.eval(0, M0, True)
    P.eval(pose0, True)
    for c in (0, 5, 15):
        pt, pj = ctx.debug_hist(0, c)
        pto, pjo = P.last_hist(c)
        np.testing.assert_allclose(pt, pto, rtol=1e-11, atol=1e-16)
        np.testing.assert_allclose(pj, pjo, rtol=1e-11, atol=1e-16)
        assert abs(pt.sum() - 1.0) < 1e-3  # few samples leave the image


def test_quaternion_vs_matrix_warp_is_within_tolerance(nid, orc, make_pair):
    """The reference CPU edge warps with the quaternion (se3quat.h:217-220), its CUDA code and this
    library with the 4x4: the two differ by rounding only; north_star tolerance 1e-5 relative."""
    p = make_pair(1000, 240, 320)
    P, ctx, pose0 = _setup(nid, orc, p, 4, 16, matrix_warp=False)
    M0 = orc.se3_to_mat16(pose0)
    nc, href = ctx.prepare(0, M0)
    nco, hrefo = P.prepare(pose0)
    assert np.array_equal(nc, nco)
    Ht, Hj, J = ctx.eval(0, M0, True)
    Hto, Hjo, erro, Jo = P.eval(pose0, True)
    np.testing.assert_allclose(Ht, Hto, rtol=1e-9)
    np.testing.assert_allclose(Hj, Hjo, rtol=1e-9)
    assert _jrel(J, Jo) < 1e-5


def test_warp_sample_per_pixel(nid, orc, make_pair):
    p = make_pair(1003, 120, 160, invalid_depth_frac=0.03)
    P, ctx, pose0 = _setup(nid, orc, p, 4, 16)
    M0 = orc.se3_to_mat16(pose0)
    got = ctx.warp_sample_f64(0, M0)
    exp = P.pixels(pose0)
    assert np.array_equal(np.isnan(got[:, 0]), np.isnan(exp[:, 0]))
    m = ~np.isnan(exp[:, 0])
    assert np.array_equal(got[m, 0], exp[m, 0]) and np.array_equal(got[m, 1], exp[m, 1])  # u, v bit-exact
    assert np.array_equal(got[m, 5], exp[m, 5]) and np.array_equal(got[m, 6], exp[m, 6])  # validity
    np.testing.assert_allclose(got[m, 2:5], exp[m, 2:5], rtol=1e-12, atol=1e-10)
    # packed float4 flavour
    g4 = ctx.warp_sample(0, M0)
    np.testing.assert_allclose(g4[m, 0], exp[m, 2], rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(g4[m, 1:3], exp[m, 3:5], rtol=1e-6, atol=1e-5)
    assert np.array_equal(g4[m, 3], exp[m, 5] + 2 * exp[m, 6])
    assert np.all(g4[~m] == 0)


@pytest.mark.parametrize("path", [NATURAL, SORTED])
def test_out_of_bounds_and_inactive_cells(nid, orc, make_pair, path):
    """A pose that pushes a band of cells out of the image (n_c < 300 -> NaN) and leaves others partly in."""
    p = make_pair(1004, 240, 320, invalid_depth_frac=0.1)
    P, ctx, pose0 = _setup(nid, orc, p, 8, 10, path=path)
    shift = orc.se3_mul(orc.se3_exp(np.array([0, 0.12, 0, 0, 0, 0])), pose0)  # ~60 px sideways
    M = orc.se3_to_mat16(shift)
    nc, href = ctx.prepare(0, M)
    nco, hrefo = P.prepare(shift)
    assert np.array_equal(nc, nco)
    assert np.any(nc < 300) and np.any(nc >= 300)
    assert np.array_equal(np.isnan(href), np.isnan(hrefo))
    # evaluate somewhere else: points that were out of bounds at prepare now contribute to P_t only
    pose = orc.se3_mul(orc.se3_exp(np.array([0, -0.03, 0.01, 0.01, 0, 0])), shift)
    Ht, Hj, J = ctx.eval(0, orc.se3_to_mat16(pose), True)
    Hto, Hjo, erro, Jo = P.eval(pose, True)
    act = ~np.isnan(hrefo)
    np.testing.assert_allclose(Ht[act], Hto[act], rtol=1e-11)
    np.testing.assert_allclose(Hj[act], Hjo[act], rtol=1e-11)
    assert np.all(np.isnan(Ht[~act]))
    assert _jrel(J, Jo) < 1e-8


def test_eval_jobs_matches_single_evals(nid, orc, make_pair):
    pa = make_pair(1000, 120, 160)
    pb = make_pair(1001, 120, 160)
    ctx = nid.Context(120, 160, 4, 16, n_pairs=2, max_jobs=6)
    poses0 = []
    for i, p in enumerate((pa, pb)):
        ctx.set_pair(i, p.depth0, p.im0, p.im1, p.T_wc0, p.intr)
        pose0 = orc.reference_perturbation(p.T_wc1)
        ctx.prepare(i, orc.se3_to_mat16(pose0))
        poses0.append(pose0)
    rng = np.random.default_rng(5)
    job_pair = [0, 1, 1, 0, 1, 0]
    poses = [orc.se3_to_mat16(orc.se3_mul(orc.se3_exp(rng.normal(size=6) * 2e-3), poses0[j])) for j in job_pair]
    Ht, Hj, J = ctx.eval_jobs(np.array(poses), job_pair, True)
    for k, (j, M) in enumerate(zip(job_pair, poses)):
        ht, hj, jj = ctx.eval(j, M, True)
        # different strip counts re-associate the sums: equal to rounding, not bit-equal
        np.testing.assert_allclose(Ht[k], ht, rtol=1e-12)
        np.testing.assert_allclose(Hj[k], hj, rtol=1e-12)
        np.testing.assert_allclose(J[k], jj, rtol=1e-8, atol=1e-11)


def test_run_to_run_determinism(nid, orc, make_pair):
    """The sorted path has no floating-point atomics and merges partials in a fixed order: results are
    bit-identical run to run, across contexts, and whether 1 or several jobs are in flight."""
    p = make_pair(1000, 240, 320)
    P, ctx, pose0 = _setup(nid, orc, p, 4, 16, path=SORTED, max_jobs=3)
    M0 = orc.se3_to_mat16(pose0)
    ctx.prepare(0, M0)
    a = ctx.eval(0, M0, True)
    for _ in range(3):
        b = ctx.eval(0, M0, True)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    Ht, Hj, J = ctx.eval_jobs(np.tile(M0, 3), [0, 0, 0], True)
    for k in range(3):
        assert np.array_equal(Ht[k], a[0]) and np.array_equal(Hj[k], a[1]) and np.array_equal(J[k], a[2])
    P2, ctx2, _ = _setup(nid, orc, p, 4, 16, path=SORTED)
    ctx2.prepare(0, M0)
    for x, y in zip(a, ctx2.eval(0, M0, True)):
        assert np.array_equal(x, y)
    # the natural-order path uses shared-memory fp64 atomics: equal to rounding only
    P3, ctx3, _ = _setup(nid, orc, p, 4, 16, path=NATURAL)
    ctx3.prepare(0, M0)
    for x, y in zip(a, ctx3.eval(0, M0, True)):
        np.testing.assert_allclose(x, y, rtol=1e-9, atol=1e-12)


def test_gn_block_and_chi2(nid, orc, make_pair):
    p = make_pair(1000, 240, 320)
    P, ctx, pose0 = _setup(nid, orc, p, 4, 16)
    M0 = orc.se3_to_mat16(pose0)
    ctx.prepare(0, M0)
    P.prepare(pose0)
    chi2, H, b, err, J = ctx.eval_gn(0, M0, DELTA)
    chi2o, Ho, bo = P.gn_system(pose0, DELTA)
    assert chi2 == pytest.approx(chi2o, rel=1e-10)
    np.testing.assert_allclose(H, Ho, rtol=1e-7, atol=1e-9 * np.abs(Ho).max())
    np.testing.assert_allclose(b, bo, rtol=1e-7, atol=1e-9 * np.abs(bo).max())


@pytest.mark.parametrize("cell,bins,rows,cols", [(4, 16, 240, 320), (16, 10, 480, 640), (4, 16, 480, 640)])
def test_lm_solve_matches_oracle(nid, orc, make_pair, cell, bins, rows, cols):
    """Converged pose within 1e-4 rad / 1e-4 x scene depth of the reference CPU path (north_star);
    in practice the whole LM trajectory coincides."""
    p = make_pair(1000, rows, cols)
    P, ctx, pose0 = _setup(nid, orc, p, cell, bins)
    ctx.prepare(0, orc.se3_to_mat16(pose0))
    P.prepare(pose0)
    pose, trace, stats = ctx.solve(0, pose0, 10, DELTA)
    poseo, its, traceo, counts = P.optimize(pose0, 10, DELTA)
    assert stats[0] == its and stats[1] == counts[0]
    assert np.array_equal(trace[:, 2], traceo[:, 2])           # same number of LM trials per iteration
    np.testing.assert_allclose(trace[:, 0], traceo[:, 0], rtol=1e-7)  # chi2 per iteration
    depth = float(np.median(p.depth0))
    assert np.max(np.abs(pose[:3] - poseo[:3])) < 1e-4 * depth
    assert 2 * np.max(np.abs(pose[3:6] - poseo[3:6])) < 1e-4


def test_solve_jobs_lockstep_equals_sequential(nid, orc, make_pair):
    pairs = [make_pair(1000 + i, 120, 160) for i in range(3)]
    ctx = nid.Context(120, 160, 4, 16, n_pairs=3, max_jobs=3)
    poses0 = []
    for i, p in enumerate(pairs):
        ctx.set_pair(i, p.depth0, p.im0, p.im1, p.T_wc0, p.intr)
        pose0 = orc.reference_perturbation(p.T_wc1)
        ctx.prepare(i, orc.se3_to_mat16(pose0))
        poses0.append(pose0)
    out, stats = ctx.solve_jobs(np.array(poses0), [0, 1, 2], 10, DELTA)
    for i in range(3):
        pose, trace, st = ctx.solve(i, poses0[i], 10, DELTA)
        assert np.array_equal(st, stats[i])
        np.testing.assert_allclose(out[i], pose, rtol=0, atol=1e-9)


def test_hard_binned_nid(nid, orc, make_pair, synth):
    p = make_pair(1000, 240, 320)
    ctx = nid.Context(240, 320, 8, 8, max_jobs=4)
    ctx.set_pair(0, p.depth0, p.im0, p.im1, p.T_wc0, p.intr)
    gt = synth.mat16_inverse(p.T_wc1)
    pose_gt = orc.se3_from_mat16(gt)
    poses = [gt] + [orc.se3_to_mat16(orc.se3_mul(orc.se3_exp(d), pose_gt))
                    for d in (np.array([0.01, 0, 0, 0, 0, 0]), np.array([0, 0, 0, 0.02, 0, 0]), np.array([0, 0.2, 0, 0, 0, 0]))]
    total, cells = ctx.hard_eval_jobs(np.array(poses))
    for k, M in enumerate(poses):
        to, co = orc.hard_nid(p.im0, p.depth0, p.im1, p.T_wc0, M, p.intr, 8, 8)
        np.testing.assert_allclose(cells[k], co, rtol=1e-12, atol=1e-15)
        assert total[k] == pytest.approx(to, rel=1e-12)
    assert total[0] < total[1] and total[0] < total[2]  # the cost surface has its minimum at the true pose


def test_f64_image_entry_point_and_rejection(nid, orc, make_pair):
    p = make_pair(1000, 120, 160)
    ctx = nid.Context(120, 160, 4, 16)
    ctx.set_pair_f64(0, p.depth0, p.im0.astype(np.float64), p.im1.astype(np.float64), p.T_wc0, p.intr)
    pose0 = orc.reference_perturbation(p.T_wc1)
    M0 = orc.se3_to_mat16(pose0)
    nc, href = ctx.prepare(0, M0)
    ctx2 = nid.Context(120, 160, 4, 16)
    ctx2.set_pair(0, p.depth0, p.im0, p.im1, p.T_wc0, p.intr)
    nc2, href2 = ctx2.prepare(0, M0)
    assert np.array_equal(nc, nc2) and np.array_equal(href, href2)
    bad = p.im1.astype(np.float64) + 0.25
    with pytest.raises(nid.NidError, match="8-bit"):
        ctx.set_pair_f64(0, p.depth0, p.im0.astype(np.float64), bad, p.T_wc0, p.intr)


def test_state_errors(nid, orc, make_pair):
    p = make_pair(1000, 120, 160)
    ctx = nid.Context(120, 160, 4, 16)
    M = np.eye(4).reshape(16)
    with pytest.raises(nid.NidError, match="not set"):
        ctx.prepare(0, M)
    ctx.set_pair(0, p.depth0, p.im0, p.im1, p.T_wc0, p.intr)
    with pytest.raises(nid.NidError, match="not prepared"):
        ctx.eval(0, M, True)
    with pytest.raises(nid.NidError):
        ctx.eval_jobs(np.tile(M, 2), [0, 0], True)  # more jobs than max_jobs


def _saturate(p, lo=110, hi=190):
    """A copy of the pair whose images are stretched so that large plateaus sit at exactly 0 and 255."""
    import dataclasses
    def stretch(im):
        v = (im.astype(np.float64) - lo) * 255.0 / (hi - lo)
        return np.clip(np.rint(v), 0, 255).astype(np.uint8)
    return dataclasses.replace(p, im0=stretch(p.im0), im1=stretch(p.im1))


@pytest.mark.parametrize("kind", ["plain", "saturated", "integer_aligned"])
def test_warp_sample_jobs_batched_kernel1(nid, orc, make_pair, synth, kind):
    """The batched kernel-1 launch (fast composed warp + packed gather, bench.py's HBM probe) against the oracle's
    per-pixel record: validity flags identical, I_c / g_x / g_y to float precision, for several poses of two
    pairs in one launch; on saturated plateaus and integer-aligned warps the decisions must be the reference's."""
    import dataclasses
    pa = make_pair(1003, 120, 160, invalid_depth_frac=0.03)
    pb = make_pair(1004, 120, 160)
    if kind == "saturated":
        pa, pb = _saturate(pa), _saturate(pb)
    poses = []
    if kind == "integer_aligned":
        pb = dataclasses.replace(pb, im1=pb.im0.copy())  # (first row / column keep their depth: taps at -1.0 are defined)
    ctx = nid.Context(pa.rows, pa.cols, 4, 16, n_pairs=2, max_jobs=4)
    probs = []
    for i, p in enumerate((pa, pb)):
        ctx.set_pair(i, p.depth0, p.im0, p.im1, p.T_wc0, p.intr)
        P = orc.Problem(p.im0, p.depth0, p.im1, p.T_wc0, p.intr, 4, 16, threads=4)
        P.set_quirks(0, 1)
        probs.append(P)
    base = [orc.reference_perturbation(pa.T_wc1), orc.reference_perturbation(pb.T_wc1)]
    if kind == "integer_aligned":
        base[1] = orc.se3_from_mat16(synth.mat16_inverse(pb.T_wc0))
    job_pair = [0, 1, 1, 0]
    xi = np.array([0.002, -0.001, 0.0015, 0.004, -0.003, 0.002])
    poses = [base[0], base[1], orc.se3_mul(orc.se3_exp(xi), base[1]), orc.se3_mul(orc.se3_exp(-xi), base[0])]
    got = ctx.warp_sample_jobs(np.stack([orc.se3_to_mat16(q) for q in poses]), job_pair)
    for j, (pr, pose) in enumerate(zip(job_pair, poses)):
        exp = probs[pr].pixels(pose)
        m = ~np.isnan(exp[:, 0])
        assert np.array_equal(got[j, m, 3], exp[m, 5] + 2 * exp[m, 6])
        np.testing.assert_allclose(got[j, m, 0], exp[m, 2], rtol=1e-6, atol=1e-5)
        np.testing.assert_allclose(got[j, m, 1:3], exp[m, 3:5], rtol=1e-6, atol=1e-5)
        assert np.all(got[j, ~m] == 0)


@pytest.mark.parametrize("path", [NATURAL, SORTED])
def test_saturated_plateaus(nid, orc, make_pair, path):
    """The reference clamps `>= 255 -> 254.999` after the bilinear sample and returns a zero B-spline
    derivative at exactly 0: on saturated plateaus the result depends on the last bit of the bilinear
    weights. The sorted path re-evaluates such pixels with the reference's exact sequence."""
    p = _saturate(make_pair(1000, 240, 320))
    assert (p.im1 == 255).mean() > 0.05 and (p.im1 == 0).mean() > 0.05
    P, ctx, pose0 = _setup(nid, orc, p, 4, 16, path=path)
    M0 = orc.se3_to_mat16(pose0)
    nc, href = ctx.prepare(0, M0)
    nco, hrefo = P.prepare(pose0)
    assert np.array_equal(nc, nco)
    pose = orc.se3_mul(orc.se3_exp(np.array([0.002, -0.001, 0.0015, 0.004, -0.003, 0.002])), pose0)
    Ht, Hj, J = ctx.eval(0, orc.se3_to_mat16(pose), True)
    Hto, Hjo, erro, Jo = P.eval(pose, True)
    np.testing.assert_allclose(Ht, Hto, rtol=1e-11)
    np.testing.assert_allclose(Hj, Hjo, rtol=1e-11)
    assert _jrel(J, Jo) < 1e-8


@pytest.mark.parametrize("path", [NATURAL, SORTED])
def test_integer_aligned_warp(nid, orc, make_pair, synth, path):
    """T_cw1 = T_wc0^-1: every pixel maps onto (almost exactly) itself, so u and v sit on integers, where
    the in-bounds tests, the (int) truncation and the first-row/column gradient quirk all switch. The
    decisions must be the reference's."""
    p = make_pair(1000, 120, 160)
    import dataclasses
    # (points with u == 0 or v == 0 exactly, whose gradient taps index column / row -1 upstream, are covered by
    # tests/test_gpu_api_state.py::test_first_row_and_column_points_at_identity)
    p = dataclasses.replace(p, im1=p.im0.copy())
    pose_id = orc.se3_from_mat16(synth.mat16_inverse(p.T_wc0))
    P = orc.Problem(p.im0, p.depth0, p.im1, p.T_wc0, p.intr, 2, 16, threads=4)
    P.set_quirks(0, 1)
    ctx = nid.Context(p.rows, p.cols, 2, 16)
    ctx.set_option("path", path)
    ctx.set_pair(0, p.depth0, p.im0, p.im1, p.T_wc0, p.intr)
    M = orc.se3_to_mat16(pose_id)
    nc, href = ctx.prepare(0, M)
    nco, hrefo = P.prepare(pose_id)
    assert np.array_equal(nc, nco)
    Ht, Hj, J = ctx.eval(0, M, True)
    Hto, Hjo, erro, Jo = P.eval(pose_id, True)
    np.testing.assert_allclose(Ht, Hto, rtol=1e-11)
    np.testing.assert_allclose(Hj, Hjo, rtol=1e-11)
    assert _jrel(J, Jo) < 1e-8


@pytest.mark.parametrize("task_px", [8, 16, 32, 128, 256])
def test_task_length_does_not_change_results(nid, orc, make_pair, task_px):
    p = make_pair(1000, 240, 320)
    P, ctx, pose0 = _setup(nid, orc, p, 4, 16, path=SORTED)
    M0 = orc.se3_to_mat16(pose0)
    ctx.prepare(0, M0)
    a = ctx.eval(0, M0, True)
    ctx.set_option("task_px", 64)
    ctx.set_option("task_px", task_px)  # a different task length invalidates the pixel store
    with pytest.raises(nid.NidError, match="not prepared"):
        ctx.eval(0, M0, True)
    ctx.prepare(0, M0)
    b = ctx.eval(0, M0, True)
    np.testing.assert_allclose(b[0], a[0], rtol=1e-12)
    np.testing.assert_allclose(b[1], a[1], rtol=1e-12)
    np.testing.assert_allclose(b[2], a[2], rtol=1e-8, atol=1e-11)


def test_full_size_c3_1280x960_32_bins_gamma(nid, orc, make_pair):
    """BASELINE config 3 at its full size: 1280x960, 4x4 cells, 32-bin histograms, strong gamma change."""
    p = make_pair(1000, 960, 1280, gamma=0.45)
    P = orc.Problem(p.im0, p.depth0, p.im1, p.T_wc0, p.intr, 4, 32, threads=8)
    P.set_quirks(0, 1)
    pose0 = orc.reference_perturbation(p.T_wc1)
    ctx = nid.Context(960, 1280, 4, 32)
    ctx.set_pair(0, p.depth0, p.im0, p.im1, p.T_wc0, p.intr)
    nc, href = ctx.prepare(0, orc.se3_to_mat16(pose0))
    nco, hrefo = P.prepare(pose0)
    assert np.array_equal(nc, nco)
    np.testing.assert_allclose(href, hrefo, rtol=1e-10)
    xi = np.array([0.002, -0.001, 0.0015, 0.004, -0.003, 0.002])
    pose = orc.se3_mul(orc.se3_exp(xi), pose0)
    Ht, Hj, J = ctx.eval(0, orc.se3_to_mat16(pose), True)
    Hto, Hjo, erro, Jo = P.eval(pose, True)
    np.testing.assert_allclose(Ht, Hto, rtol=1e-10)
    np.testing.assert_allclose(Hj, Hjo, rtol=1e-10)
    assert _jrel(J, Jo) < 1e-5  # north_star's bar, stated; measured orders of magnitude below
    assert _jrel(J, Jo) < 1e-8


def test_full_size_c2_batch_properties(nid, orc, make_pair):
    """BASELINE config 4 shape (a batch of C2 pairs) through size-independent properties: every job of a batch equals
    the same evaluation submitted alone, bit for bit, whatever the batch composition; identical poses on
    identical pairs give identical results; the cost-only call returns the entropies of the cost+Jacobian call."""
    pairs = [make_pair(1000 + i, 480, 640) for i in range(2)]
    n_slots, n_jobs = 6, 40
    ctx = nid.Context(480, 640, 4, 16, n_pairs=n_slots, max_jobs=n_jobs)
    pose0 = [orc.reference_perturbation(p.T_wc1) for p in pairs]
    for s in range(n_slots):
        p = pairs[s % 2]
        ctx.set_pair(s, p.depth0, p.im0, p.im1, p.T_wc0, p.intr)
        ctx.prepare(s, orc.se3_to_mat16(pose0[s % 2]))
    rng = np.random.default_rng(9)
    job_pair = rng.integers(0, n_slots, size=n_jobs).astype(np.int32)
    xis = rng.uniform(-1, 1, size=(n_jobs, 6)) * 3e-3
    poses = np.stack([orc.se3_to_mat16(orc.se3_mul(orc.se3_exp(xis[j]), pose0[job_pair[j] % 2])) for j in range(n_jobs)])
    Ht, Hj, J = ctx.eval_jobs(poses, job_pair, True)
    Ht2, Hj2, _ = ctx.eval_jobs(poses, job_pair, False)
    assert np.array_equal(Ht, Ht2) and np.array_equal(Hj, Hj2)
    for j in (0, 7, 23, 39):
        a, b, c = ctx.eval(int(job_pair[j]), poses[j], True)
        assert np.array_equal(a, Ht[j]) and np.array_equal(b, Hj[j]) and np.array_equal(c, J[j])
    # slots s and s+2 hold the same pair: same pose => same bits
    same = np.stack([poses[0]] * 3)
    Ha, Hb, Jc = ctx.eval_jobs(same, np.array([job_pair[0] % 2, job_pair[0] % 2 + 2, job_pair[0] % 2 + 4], dtype=np.int32), True)
    assert np.array_equal(Ha[0], Ha[1]) and np.array_equal(Ha[0], Ha[2]) and np.array_equal(Jc[0], Jc[2])
    # and one job against the oracle at full size
    P = orc.Problem(pairs[0].im0, pairs[0].depth0, pairs[0].im1, pairs[0].T_wc0, pairs[0].intr, 4, 16, threads=8)
    P.set_quirks(0, 1)
    P.prepare(pose0[0])
    j = int(np.where(job_pair % 2 == 0)[0][0])
    Hto, Hjo, _, Jo = P.eval(orc.se3_mul(orc.se3_exp(xis[j]), pose0[0]), True)
    np.testing.assert_allclose(Ht[j], Hto, rtol=1e-10)
    np.testing.assert_allclose(Hj[j], Hjo, rtol=1e-10)
    assert _jrel(J[j], Jo) < 1e-8


def test_set_target_reuses_the_reference_frame(nid, orc, make_pair):
    """Tracking against a key frame: the second target replaces only im1 of the pair slot; results equal
    those of a context that was given the (reference, second target) pair from scratch, bit for bit, and evaluating
    before the new prepare is refused."""
    import dataclasses
    pa = make_pair(1000, 240, 320)
    pb = make_pair(1001, 240, 320)
    pose0 = orc.reference_perturbation(pa.T_wc1)
    M0 = orc.se3_to_mat16(pose0)
    ctx = nid.Context(pa.rows, pa.cols, 4, 16)
    ctx.set_pair(0, pa.depth0, pa.im0, pa.im1, pa.T_wc0, pa.intr)
    ctx.prepare(0, M0)
    ctx.eval(0, M0, True)
    ctx.set_target(0, pb.im1)
    with pytest.raises(nid.NidError):
        ctx.eval(0, M0, True)
    nc, href = ctx.prepare(0, M0)
    got = ctx.eval(0, M0, True)
    ref = nid.Context(pa.rows, pa.cols, 4, 16)
    ref.set_pair(0, pa.depth0, pa.im0, pb.im1, pa.T_wc0, pa.intr)
    nc2, href2 = ref.prepare(0, M0)
    exp = ref.eval(0, M0, True)
    assert np.array_equal(nc, nc2) and np.array_equal(href, href2)
    for a, b in zip(got, exp):
        assert np.array_equal(a, b, equal_nan=True)
    P = orc.Problem(pa.im0, pa.depth0, pb.im1, pa.T_wc0, pa.intr, 4, 16, threads=4)
    P.set_quirks(0, 1)
    P.prepare(pose0)
    Hto, Hjo, erro, Jo = P.eval(pose0, True)
    np.testing.assert_allclose(got[0], Hto, rtol=1e-11)
    np.testing.assert_allclose(got[1], Hjo, rtol=1e-11)
    assert _jrel(got[2], Jo) < 1e-8


def test_more_jobs_than_one_geometry_table(nid, orc, make_pair):
    """The pixel kernels take the per-job geometry as a by-value table of 96 entries; a larger batch is issued in
    chunks. Jobs on either side of the chunk boundary must equal the same evaluation submitted alone, bit for bit;
    the batched kernel-1 entry point and the LM driver (two half-batches) cross the same boundary."""
    p = make_pair(1000, 120, 160)
    n_jobs = 100
    ctx = nid.Context(p.rows, p.cols, 2, 16, n_pairs=2, max_jobs=n_jobs)
    pose0 = orc.reference_perturbation(p.T_wc1)
    for s in range(2):
        ctx.set_pair(s, p.depth0, p.im0, p.im1, p.T_wc0, p.intr)
        ctx.prepare(s, orc.se3_to_mat16(pose0))
    rng = np.random.default_rng(3)
    job_pair = (np.arange(n_jobs) % 2).astype(np.int32)
    xis = rng.uniform(-1, 1, size=(n_jobs, 6)) * 3e-3
    poses = np.stack([orc.se3_to_mat16(orc.se3_mul(orc.se3_exp(xis[j]), pose0)) for j in range(n_jobs)])
    Ht, Hj, J = ctx.eval_jobs(poses, job_pair, True)
    for j in (0, 95, 96, 99):
        a, b, c = ctx.eval(int(job_pair[j]), poses[j], True)
        assert np.array_equal(a, Ht[j]) and np.array_equal(b, Hj[j]) and np.array_equal(c, J[j])
    ws = ctx.warp_sample_jobs(poses, job_pair)
    one = ctx.warp_sample_jobs(poses[97:98], job_pair[97:98])
    assert np.array_equal(ws[97], one[0])
    p7 = np.stack([pose0] * n_jobs)
    out, stats = ctx.solve_jobs(p7, job_pair, 3)
    solo, _, st1 = ctx.solve(1, pose0, 3)
    assert np.array_equal(out[99], solo) and stats[99, 0] == st1[0]
