"""Host-side logic: problem records, sample problems, initial guess, workloads, sharding rule."""
import numpy as np
import pytest

from successiveconvexification_b200 import first_round, sample_problems as sp, sharding, workloads
from successiveconvexification_b200.defns import (CProbInfo, DescentProblem, ExoatmosphericData, LinPoint, LinRes,
                                                  ProbInfo)
from successiveconvexification_b200.dynamics import make_state


def test_descent_problem_defaults():
    p = DescentProblem()                      # master.jl:65-70
    assert (p.g, p.mdry, p.mwet, p.K, p.imax, p.wNu, p.tf_guess, p.sos) == (1.0, 1.0, 2.0, 50, 15, 1e5, 1.0, 5.0)
    assert isinstance(p.aero, ExoatmosphericData)
    info = ProbInfo(p)                        # master.jl:82
    assert info.a == p.alpha and info.g0 == p.g
    assert np.allclose(info.jBi @ info.jB, np.eye(3))


def test_normalize_problem_quirks(prob_aero):
    raw = sp.base_prob()
    n = sp.normalize_problem(raw)
    assert n.rIi == pytest.approx([1.0, 1.0, 0.1])
    assert n.vIf == pytest.approx(n.vIi)                        # sample_problems.jl:15 copies vIi
    assert n.rFB == pytest.approx(raw.rFB)                      # scaled by 1/Ut = 1 (sample_problems.jl:16)
    assert n.nuTol == DescentProblem().nuTol                    # not forwarded
    assert n.rTB == pytest.approx([-0.00426114, 0, 0])
    assert prob_aero.aero.force_scalar == pytest.approx(1 / (1000.0 * 66018.0))


def test_linear_points_and_make_state(prob_aero):
    pts = first_round.linear_points(prob_aero)
    assert len(pts) == prob_aero.K + 1
    assert pts[0].state[0] == prob_aero.mwet and pts[-1].state[0] == pytest.approx(prob_aero.mdry)
    for p in pts[::10]:
        assert np.linalg.norm(p.state[7:11]) == pytest.approx(1.0, rel=1e-15)
        assert p.control == pytest.approx([p.state[0] * prob_aero.g, 0, 0])
    s = make_state(pts[0], pts[1], 1.0)
    assert s.shape == (21,) and s[20] == 1.0 and np.array_equal(s[17:20], pts[1].control)
    X, U, sigma, dt = workloads.sample_trajectory(prob_aero)
    assert X.shape == (1, 51, 14) and dt == 1 / 51
    Xb, Ub = workloads.linear_points_batch(prob_aero, prob_aero.K, prob_aero.rIi[None], prob_aero.vIi[None])
    assert np.abs(Xb - X).max() <= 1e-15 and np.abs(Ub - U).max() <= 1e-18


def test_rotation_between():
    q = first_round.rotation_between([1, 0, 0], [0.1, 0.2, 0.0])
    assert q == pytest.approx([0.85065080835204, 0, 0, 0.5257311121191336], rel=1e-13)
    q = first_round.rotation_between([1, 0, 0], [-2.0, 0, 0])      # antiparallel
    assert np.linalg.norm(q) == pytest.approx(1.0) and abs(q[0]) < 1e-12


def test_probinfo_c_layout(prob_aero):
    info = ProbInfo(prob_aero)
    arr = workloads.probinfo_array(info, 3)
    assert arr.dtype.itemsize == 248
    assert arr["jB"][1].reshape(3, 3, order="F") == pytest.approx(info.jB)
    assert arr["aero_kind"].tolist() == [1, 1, 1]
    c = info.to_c()
    assert isinstance(c, CProbInfo) and c.Tmin == prob_aero.Tmin


def test_workloads_reproducible_and_shaped(prob_aero):
    a = workloads.monte_carlo_batch(prob_aero, 50, 8, 1003, shard=2)
    b = workloads.monte_carlo_batch(prob_aero, 50, 8, 1003, shard=2)
    c = workloads.monte_carlo_batch(prob_aero, 50, 8, 1003, shard=3)
    assert all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3]))
    assert not np.array_equal(a[0], c[0])
    X, U, sigma, P = a
    assert X.shape == (8, 51, 14) and U.shape == (8, 51, 3) and sigma.shape == (8,) and P.shape == (1,)
    nu = np.linalg.norm(U, axis=-1)
    assert nu.min() >= prob_aero.Tmin * (1 - 1e-12) and nu.max() <= prob_aero.Tmax * (1 + 1e-12)
    assert np.linalg.norm(X[..., 7:11], axis=-1) == pytest.approx(1.0, rel=1e-12)
    Xs, Us, ss, Ps = workloads.monte_carlo_batch(prob_aero, 20, 5, 1002, sweep=True)
    assert Ps.shape == (5,) and len(set(Ps["a"])) == 5


def test_shard_range_partitions():
    for B in (0, 1, 7, 262144):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(B, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def test_linres_views():
    D = np.asfortranarray(np.arange(14 * 21, dtype=float).reshape(14, 21))
    r = LinRes(np.zeros(14), D)
    assert r.A.shape == (14, 14) and r.Bm.shape == (14, 3) and r.Bp.shape == (14, 3) and r.Sigma.shape == (14,)
    assert LinPoint(range(14), [1, 2, 3]).state.dtype == np.float64


# ---- fixed-pattern sparse SOCP rows: the pattern is host-only code of the library --------------
@pytest.mark.parametrize("K", [1, 2, 3, 50])
def test_socp_pattern_matches_reference_assembly(K):
    """CSC pattern from the library == nonzero pattern of the reference-style triplet assembly (rocketland.jl:117-133,
    194-201) with a fully dense derivative."""
    from oracle import socp_assembly
    from successiveconvexification_b200 import rocketland
    nr, nc, colptr, rowind = rocketland.socp_pattern(K + 1)
    assert (nr, nc, len(rowind)) == (15 * K + 1, 31 * (K + 1) + 1, 325 * K + 3) == rocketland.socp_dims(K + 1)
    assert colptr[0] == 0 and colptr[-1] == len(rowind) and np.all(np.diff(colptr) >= 0)
    for j in range(nc):
        r = rowind[colptr[j]:colptr[j + 1]]
        assert np.all(np.diff(r) > 0), "rows must ascend strictly inside a column"
    rng = np.random.default_rng(K)
    D = rng.uniform(0.5, 1.5, (K, 14, 21))
    M, _ = socp_assembly.assemble_dense(D, rng.normal(size=(K, 14)), rng.normal(size=(K + 1, 14)),
                                        rng.uniform(0.5, 1.5, (K + 1, 3)), 0.1)
    v = socp_assembly.variable_index(K)
    first = v["dxv"][0, 0]
    assert first == 17 * (K + 1)
    assert not M[:, :first].any(), "xv / uv do not appear in these rows"
    local = M[:, first:]
    assert local.shape[1] == nc
    pat = np.zeros_like(local, dtype=bool)
    for j in range(nc):
        pat[rowind[colptr[j]:colptr[j + 1]], j] = True
    assert np.array_equal(pat, local != 0)
    cols = rocketland.variable_columns(K)
    for name in ("dxv", "duv", "nuv"):
        assert np.array_equal(cols[name] + first, v[name])
    assert cols["dsig"] + first == v["dsig"]


def test_socp_entry_points_reject_bad_arguments():
    from successiveconvexification_b200 import rocketland, _lib
    with pytest.raises(_lib.ScvxError):
        rocketland.socp_dims(1)


def test_fin_table_loader(tmp_path):
    """aero/fin.csv layout (aero/AeroTable.jl:94-112): header lift,drag,mach,aoa; Mach varies fastest."""
    from successiveconvexification_b200 import aerodynamics
    mach = 0.01 + 0.025 * np.arange(4)
    defl = 0.1 * np.arange(3)
    rows = [(10 * j + i, 100 * j + i, mach[i], defl[j]) for j in range(3) for i in range(4)]
    p = tmp_path / "fin.csv"
    p.write_text("lift,drag,mach,aoa\n" + "\n".join(",".join(repr(float(v)) for v in r) for r in rows) + "\n")
    lift, drag, axes = aerodynamics.load_fin_table(p, n_mach=4, n_defl=3)
    assert lift.shape == (4, 3) and lift[2, 1] == 12.0 and drag[3, 2] == 203.0
    assert axes == pytest.approx((0.01, 0.025, 0.0, 0.1))
