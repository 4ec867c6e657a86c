"""Pin the CPU oracle (oracle/scvx_oracle.cpp).  The reference has no tests / golden vectors for this path
("parity unpinned"), so the oracle is pinned by independent means:
  * the survey's scratch anchors (50-digit / scipy cross-checked numbers),
  * a numpy complex-step restatement in another language (oracle/py_restatement.py),
  * scipy's natural cubic spline (Interpolations.jl Cubic(Line(OnGrid())) == natural spline),
  * central finite differences, structural zeros, and the committed golden vectors.
"""
import os

import numpy as np
import pytest

from oracle import oracle, py_restatement as pr
from successiveconvexification_b200 import workloads
from successiveconvexification_b200.defns import ProbInfo
from successiveconvexification_b200.first_round import linear_points

from conftest import GOLDEN, assert_parity, parity_report


def _inp(prob, i, sigma=1.0):
    pts = linear_points(prob)
    return np.concatenate([pts[i].state, pts[i].control, pts[i + 1].control, [sigma]])


def test_survey_anchor_constants(prob_aero):
    info = ProbInfo(prob_aero)
    assert info.g0 == pytest.approx(0.00982, rel=1e-15)
    assert info.a == pytest.approx(0.345, rel=1e-15)
    assert info.sos == pytest.approx(0.352, rel=1e-15)
    assert np.diag(info.jB) == pytest.approx([1.0979889007543397e-6, 3.1406851161804359e-5, 3.1406851161804359e-5], rel=1e-14)
    assert prob_aero.mdry == pytest.approx(0.9989245357326789, rel=1e-15)
    assert prob_aero.Tmin == pytest.approx(0.007098954830500773, rel=1e-15)
    assert info.aero.force_scalar == pytest.approx(1.514738404677512e-8, rel=1e-15)
    assert info.aero.length_scalar == pytest.approx(1e-3, rel=1e-15)
    inp = _inp(prob_aero, 0)
    expect = [1, 1, 1, 0.1, -0.1, -0.2, 0, 0.85065080835204, 0, 0, 0.5257311121191336, 0, 0, 0,
              0.00982, 0, 0, 0.0098197887788179, 0, 0, 1]
    assert inp == pytest.approx(expect, rel=1e-13, abs=1e-300)


def test_survey_anchor_spline(oracle_tables):
    tb = oracle_tables
    v, g = oracle.spline_eval(tb.drag, tb.geom, -0.93, 0.63)
    assert v == pytest.approx(-34.69784799236458, rel=1e-13)
    assert g == pytest.approx([-453.5140056352759, -141.0370757488106], rel=1e-12)
    assert oracle.spline_eval(tb.lift, tb.geom, -0.93, 0.63)[0] == pytest.approx(-132.1065535448402, rel=1e-13)
    assert oracle.spline_eval(tb.drag, tb.geom, -1.0, 0.6352465845169858)[0] == pytest.approx(-4.608211225266031, rel=1e-13)
    assert oracle.spline_eval(tb.drag, tb.geom, 0.5, 1.2)[0] == pytest.approx(-802.6218855614263, rel=1e-13)
    v, g = oracle.spline_eval(tb.drag, tb.geom, 1.2, 2.0)          # Flat extrapolation, strictly outside
    assert v == pytest.approx(-17.848844109060312, rel=1e-13)
    assert list(g) == [0.0, 0.0]


def test_survey_anchor_linearisation(prob_aero, prob_exo, oracle_tables):
    info, inp = ProbInfo(prob_aero), _inp(prob_aero, 0)
    f = oracle.rhs(info, oracle_tables, inp[:14], inp[14:17], 1.0)
    assert f[[4, 5]] == pytest.approx([-0.00542833127563264, 0.00878333744873471], rel=1e-12)
    assert f[0] == pytest.approx(-0.0033879, rel=1e-12)
    fe = oracle.rhs(ProbInfo(prob_exo), None, inp[:14], inp[14:17], 1.0)
    assert fe[[4, 5]] == pytest.approx([-0.00542836249219041, 0.00878327501561917], rel=1e-12)
    blk = oracle.linearize_interval(info, oracle_tables, inp, 1 / 51, 10, 0)
    D = blk[:, 1:22]
    assert blk[:7, 0] == pytest.approx([0.9999335713026599, 0.9979851084317743, 0.9961661637871958, 0.1,
                                        -0.10010629003522521, -0.19982748208816042, 0], rel=1e-13, abs=1e-300)
    assert D[4, 0] == pytest.approx(-8.6408005837847326e-5, rel=1e-12)
    assert D[1, 4] == pytest.approx(1.9607835053252230e-2, rel=1e-13)
    assert D[0, 14] == pytest.approx(-3.3823529411764709e-3, rel=1e-13)
    assert D[12, 16] == pytest.approx(1.3301518874883955, rel=1e-13)
    assert D[12, 19] == pytest.approx(1.3301518874883955, rel=1e-13)
    assert D[1, 20] == pytest.approx(-2.0689474410495513e-3, rel=1e-12)
    assert blk[1, 22] == pytest.approx(2.1232802717809918e-3, rel=1e-11)
    assert np.linalg.norm(D) == pytest.approx(4.640907526203618, rel=1e-13)
    blk = oracle.linearize_interval(info, oracle_tables, inp, 1 / 51, 10, 1)       # TEXTBOOK twin
    assert blk[[1, 2, 4, 5], 0] == pytest.approx([0.9980381721904007, 0.9960801198517665, -0.10010643593460077,
                                                  -0.19982777383116213], rel=1e-13)
    assert blk[1, 21] == pytest.approx(-1.9628712867821442e-3, rel=1e-12)
    assert np.linalg.norm(blk[:, 1:22]) == pytest.approx(4.591237975482222, rel=1e-13)


def test_prefilter_interpolates_and_matches_scipy(prob_aero, oracle_tables):
    from scipy.interpolate import CubicSpline
    tb = oracle_tables
    s = np.asarray(prob_aero.aero.drag_itrp.samples)
    n1, n2 = s.shape
    xs = -1.0 + np.arange(n1) / 90.0
    ys = np.arange(n2) * 0.025
    # exact interpolation at grid points
    for (i, j) in [(0, 0), (5, 7), (90, 30), (180, 60), (179, 1)]:
        v, _ = oracle.spline_eval(tb.drag, tb.geom, xs[i], ys[j])
        assert v == pytest.approx(s[i, j], rel=1e-12, abs=1e-9 * np.abs(s).max())
    # tensor-product natural cubic spline
    rng = np.random.default_rng(7)
    scale = np.abs(s).max()
    for _ in range(40):
        x, y = rng.uniform(-1, 1), rng.uniform(0, 1.5)
        col = CubicSpline(xs, s, axis=0, bc_type="natural")(x)          # n2 values
        ref = CubicSpline(ys, col, bc_type="natural")(y)
        dref_dy = CubicSpline(ys, col, bc_type="natural")(y, 1)
        v, g = oracle.spline_eval(tb.drag, tb.geom, x, y)
        assert abs(v - ref) <= 1e-12 * scale
        assert abs(g[1] - dref_dy) <= 1e-10 * scale / 0.025
    # numpy dense-solve prefilter of the independent restatement
    assert np.abs(pr.prefilter(s) - tb.drag).max() <= 1e-13 * np.abs(tb.drag).max()


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("aero", [True, False])
def test_oracle_vs_complex_step_restatement(prob_aero, prob_exo, oracle_tables, mode, aero):
    prob = prob_aero if aero else prob_exo
    info = ProbInfo(prob)
    P = pr.probinfo_dict(info)
    T = pr.tables_dict(prob_aero.aero) if aero else None
    tb = oracle_tables if aero else None
    rng = np.random.default_rng(11)
    for trial, i in enumerate([0, 17, 49]):
        # LITERAL stage increments are not scaled by the sub-step (dynamics.jl:126-128): the discrete map is
        # ill-conditioned for large sigma, so the literal twin is exercised near the reference's sigma = 1
        inp = _inp(prob, i, sigma=(1.0 + 4.0 * trial) if mode == 1 else (1.0 + 0.25 * trial))
        if trial > 0:     # leave the |dp| >= 0.95 branch, spin, gimbal
            inp[7:11] += rng.normal(0, 0.2, 4); inp[7:11] /= np.linalg.norm(inp[7:11])
            inp[11:14] += rng.normal(0, 0.05, 3); inp[15:17] += rng.normal(0, 0.003, 2)
        e, D, z = pr.linearize_interval(P, T, inp, 1 / 51, 10, mode)
        blk = oracle.linearize_interval(info, tb, inp, 1 / 51, 10, mode)
        ref = np.concatenate([e[None], D.T, z[None]])[None]             # (1, 23, 14)
        assert_parity(blk.T[None], ref)


@pytest.mark.parametrize("mode", [0, 1])
def test_oracle_vs_50_digit_arithmetic(prob_aero, oracle_tables, mode):
    """The FP64 oracle against the same map evaluated and differentiated at 50 digits (oracle/mp_restatement.py):
    the residual is pure FP64 rounding of the oracle, and it sets the floor of the parity metric (conftest)."""
    from oracle import mp_restatement as mpr
    info = ProbInfo(prob_aero)
    P = mpr.probinfo_mp(info)
    tb = oracle_tables
    T = dict(drag=tb.drag, lift=tb.lift, geom=tb.geom)
    rng = np.random.default_rng(5)
    inp = _inp(prob_aero, 17, sigma=1.0 if mode == 0 else 6.0)
    inp[7:11] += rng.normal(0, 0.2, 4); inp[7:11] /= np.linalg.norm(inp[7:11])        # lift branch, spin, gimbal
    inp[11:14] += rng.normal(0, 0.05, 3); inp[15:17] += rng.normal(0, 0.003, 2)
    e, D = mpr.linearize_interval(P, T, inp, 1 / 51, 10, mode)
    blk = oracle.linearize_interval(info, tb, inp, 1 / 51, 10, mode)                  # (14, 23) column = block column
    e64 = np.array([float(v) for v in e])
    D64 = np.array([[float(v) for v in row] for row in D])
    assert np.abs(blk[:, 0] - e64).max() <= 4e-16 * np.abs(e64).max()
    got = blk[:, 1:22]
    # per part, relative to the part's largest entry: FP64 evaluation of the oracle is good to a few 1e-16 of that
    for lo, hi in ((0, 14), (14, 17), (17, 20), (20, 21)):
        part, ref = got[:, lo:hi], D64[:, lo:hi]
        assert np.abs(part - ref).max() <= 2e-14 * np.abs(ref).max()
    # and it passes the parity metric used for the GPU with a wide margin
    z = e64 - D64 @ inp
    ref_blk = np.concatenate([e64[None], D64.T, z[None]])[None]
    rep = parity_report(blk.T[None], ref_blk)
    assert max(rep.values()) <= 1e-11, rep


def test_oracle_vs_finite_differences(prob_aero, oracle_tables):
    info = ProbInfo(prob_aero)
    rng = np.random.default_rng(3)
    inp = _inp(prob_aero, 10, sigma=3.0)
    inp[7:11] += rng.normal(0, 0.2, 4); inp[7:11] /= np.linalg.norm(inp[7:11])
    inp[11:14] = rng.normal(0, 0.05, 3)
    blk = oracle.linearize_interval(info, oracle_tables, inp, 1 / 51, 10, 0)
    D = blk[:, 1:22]
    for c in range(21):
        h = 1e-6 * max(1.0, abs(inp[c]))
        ip, im = inp.copy(), inp.copy()
        ip[c] += h; im[c] -= h
        fd = (oracle.rk4(info, oracle_tables, ip, 1 / 51) - oracle.rk4(info, oracle_tables, im, 1 / 51)) / (2 * h)
        assert np.abs(fd - D[:, c]).max() <= 2e-6 * max(1.0, np.abs(D[:, c]).max())


def test_structure(prob_aero, oracle_tables):
    """Nothing depends on position: D[:, r] = [0; I3; 0] exactly; z closes the affine model."""
    info = ProbInfo(prob_aero)
    X, U, sigma, params = workloads.monte_carlo_batch(prob_aero, 4, 5, 99)
    blocks, err, tlb, _ = oracle.linearize_batch(params, oracle_tables, X, U, sigma, 0.2, 10, 1)
    D = blocks[:, :, 1:22, :]                                              # (B, ni, 21 cols, 14 rows)
    expect = np.zeros((3, 14)); expect[[0, 1, 2], [1, 2, 3]] = 1.0
    assert np.array_equal(D[:, :, 1:4, :], np.broadcast_to(expect, D[:, :, 1:4, :].shape))
    inp = np.concatenate([X[:, :-1], U[:, :-1], U[:, 1:], np.broadcast_to(sigma[:, None, None], X[:, :-1, :1].shape)], axis=-1)
    z = blocks[:, :, 0, :] - np.einsum("bicr,bic->bir", D, inp)
    assert np.abs(z - blocks[:, :, 22, :]).max() <= 1e-13 * np.abs(D).max()
    assert np.abs(err - (blocks[:, :, 0, :] - X[:, 1:])).max() == 0.0
    nu = np.linalg.norm(U, axis=-1)
    assert np.abs(tlb[..., 3] - (info.Tmin - nu)).max() <= 1e-18
    assert np.abs(tlb[..., :3] + U / nu[..., None]).max() <= 1e-16


def test_golden_vectors(prob_aero, prob_exo, oracle_tables):
    """The oracle reproduces the committed golden vectors (regression pin of the checker itself)."""
    g = np.load(os.path.join(GOLDEN, "golden_c2.npz"))
    X, U, sigma, dt = workloads.sample_trajectory(prob_aero)
    for aname, prob, tb in (("aero", prob_aero, oracle_tables), ("exo", prob_exo, None)):
        for mname, mode in (("literal", 0), ("textbook", 1)):
            blocks, _, _, _ = oracle.linearize_batch(ProbInfo(prob), tb, X, U, sigma, dt, 10, mode, False, False)
            assert_parity(blocks[0], g[f"{aname}_{mname}"], tol=1e-13)
    gm = np.load(os.path.join(GOLDEN, "golden_mc.npz"))
    Xm, Um, sm, Pm = workloads.monte_carlo_batch(prob_aero, 5, 6, 4242, sweep=True, sigma_range=(0.8, 1.5))
    blocks, err, tlb, _ = oracle.linearize_batch(Pm, oracle_tables, Xm, Um, sm, 1.0 / 6.0)
    assert_parity(blocks, gm["blocks"], tol=1e-13)
    assert np.abs(err - gm["lin_err"]).max() <= 1e-15
    assert np.abs(tlb - gm["tlb"]).max() <= 1e-15


def test_predict_matches_block_endpoint(prob_aero, oracle_tables):
    X, U, sigma, params = workloads.monte_carlo_batch(prob_aero, 3, 4, 5)
    blocks, _, _, _ = oracle.linearize_batch(params, oracle_tables, X, U, sigma, 0.25, 10, 0, False, False)
    end, _ = oracle.predict_batch(params, oracle_tables, X, U, sigma, 0.25)
    assert np.array_equal(end, blocks[:, :, 0, :])


def test_binary128_evaluation_and_branch_signatures(prob_aero, oracle_tables):
    """The oracle's IEEE binary128 evaluation (same operation sequence, 113-bit arithmetic) is the yardstick of the
    conditioning-aware parity checks: at the reference's sigma = 1 the FP64 oracle sits within rounding of it; the
    distance grows with sigma under the LITERAL stage rule (dynamics.jl:126-128) and stays small under TEXTBOOK."""
    from conftest import parity_metric_per_interval
    X, U, sigma, dt = workloads.sample_trajectory(prob_aero)
    info = ProbInfo(prob_aero)
    b64, s64 = oracle.linearize_batch_ex(info, oracle_tables, X, U, sigma, dt, 10, 0, precision=0)
    bq, sq = oracle.linearize_batch_ex(info, oracle_tables, X, U, sigma, dt, 10, 0, precision=1)
    plain, _, _, _ = oracle.linearize_batch(info, oracle_tables, X, U, sigma, dt, 10, 0, False, False)
    assert np.array_equal(b64, plain)                       # the _ex entry point is the same FP64 arithmetic
    # the sample trajectory flies exactly tail-first (cos(aoa) = -1): rounding decides on which side of the clamp bound a
    # few intervals land (the signature differs, the value does not — the clamped quantity has zero gradient there)
    assert (s64 != sq).sum() <= 8 and len(set(s64.reshape(-1).tolist())) > 1
    assert parity_metric_per_interval(b64, bq).max() <= 1e-12
    # 50-digit central differences (mp_restatement) and binary128 forward mode agree far below FP64 resolution
    from oracle import mp_restatement as mpr
    inp = np.concatenate([X[0, 0], U[0, 0], U[0, 1], sigma[:1]])
    T = dict(drag=oracle_tables.drag, lift=oracle_tables.lift, geom=oracle_tables.geom)
    e, D = mpr.linearize_interval(mpr.probinfo_mp(info), T, inp, dt, 10, 0)
    D64 = np.array([[float(v) for v in row] for row in D])
    assert np.abs(D64.T - bq[0, 0, 1:22]).max() <= 1e-15 * np.abs(D64).max()
    # conditioning: LITERAL degrades with sigma, TEXTBOOK does not
    Xm, Um, sm, P = workloads.monte_carlo_batch(prob_aero, 4, 12, 1003, sigma_range=(12.0, 15.0))
    lit64, a = oracle.linearize_batch_ex(P, oracle_tables, Xm, Um, sm, 0.2, 10, 0, precision=0)
    litq, b = oracle.linearize_batch_ex(P, oracle_tables, Xm, Um, sm, 0.2, 10, 0, precision=1)
    txt64, _ = oracle.linearize_batch_ex(P, oracle_tables, Xm, Um, sm, 0.2, 10, 1, precision=0)
    txtq, _ = oracle.linearize_batch_ex(P, oracle_tables, Xm, Um, sm, 0.2, 10, 1, precision=1)
    same = (a == b).reshape(-1)
    assert parity_metric_per_interval(txt64, txtq).max() <= 1e-10
    assert parity_metric_per_interval(lit64, litq)[same].max() > 1e-10     # no FP64 implementation holds 1e-10 here


def test_fins_extension_is_consistent(prob_aero, oracle_tables):
    """SURVEY.md §8f-4 oracle extension (no reference consumer): (a) with zero fin commands and a zero torque table it is
    the 3-control model; (b) its forward-mode Jacobian agrees with central differences of its own endpoint map."""
    import copy
    X, U, sigma, P = workloads.monte_carlo_batch(prob_aero, 4, 5, 21, sigma_range=(0.8, 1.5))
    U5 = np.concatenate([U, np.zeros(U.shape[:2] + (2,))], axis=-1)
    tb0 = copy.copy(oracle_tables)
    tb0.trq = np.zeros_like(oracle_tables.trq)
    b5, _ = oracle.linearize_batch_fins(P, tb0, X, U5, sigma, 0.2, 10, 1)
    b3, _, _, _ = oracle.linearize_batch(P, oracle_tables, X, U, sigma, 0.2, 10, 1)
    cols5 = list(range(1, 15)) + [15, 16, 17, 20, 21, 22, 25]
    assert np.abs(b5[:, :, 0] - b3[:, :, 0]).max() <= 1e-15
    assert np.abs(b5[:, :, cols5] - b3[:, :, 1:22]).max() <= 1e-13 * np.abs(b3[:, :, 1:22]).max()
    # (b) central differences, with fins and torque active
    rng = np.random.default_rng(3)
    U5 = np.concatenate([U, rng.normal(0, 0.002, U.shape[:2] + (2,))], axis=-1)
    base, _ = oracle.linearize_batch_fins(P, oracle_tables, X[:1, :2], U5[:1, :2], sigma[:1], 0.2, 10, 1)
    D = base[0, 0, 1:26]                                   # rows = columns of D
    inp = np.concatenate([X[0, 0], U5[0, 0], U5[0, 1], sigma[:1]])
    for j in (0, 5, 8, 12, 17, 18, 23, 24):
        h = 1e-6 * max(1.0, abs(inp[j]))
        ends = []
        for sgn in (+1, -1):
            v = inp.copy(); v[j] += sgn * h
            Xp = np.stack([v[:14], X[0, 1]])[None]; Up = np.stack([v[14:19], v[19:24]])[None]
            e, _ = oracle.linearize_batch_fins(P, oracle_tables, Xp, Up, v[24:25], 0.2, 10, 1)
            ends.append(e[0, 0, 0])
        fd = (ends[0] - ends[1]) / (2 * h)
        assert np.abs(fd - D[j]).max() <= 1e-6 * max(1.0, np.abs(D[j]).max()), j
