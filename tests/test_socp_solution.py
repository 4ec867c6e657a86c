"""SOCP-solution parity (BASELINE.json north_star: "the resulting SOCP solutions must match to within the solver
tolerance").  The reference solves its SCvx subproblem with Mosek through MathOptInterface (rocketland.jl:53-219, 271-283);
neither exists here, so the full K=50 subproblem is restated in conic standard form (oracle/socp_model.py) and solved
by a small interior-point method (oracle/socp_solver.py) — both test infrastructure.  The trajectory-dependent rows come
from the PRODUCT's fixed-pattern sparse writer (scvx_socp_pattern / scvx_socp_values_batch)."""
import numpy as np
import pytest

from oracle import oracle, socp_assembly, socp_model, socp_solver
from successiveconvexification_b200 import rocketland, workloads
from successiveconvexification_b200.defns import ProbInfo

SOLVER_TOL = 1e-9          # relative gap / residual tolerance of the test solver (Mosek's default is 1e-8)


def _vals_from_dense(M, K, pattern):
    nr, nc, colptr, rowind = pattern
    first = socp_assembly.variable_index(K)["dxv"][0, 0]
    Ml = M[:, first:]
    return np.concatenate([Ml[rowind[colptr[j]:colptr[j + 1]], j] for j in range(nc)])


def _solve(prob, K, X, U, pattern, vals, const, rk=100.0):
    mdl = socp_model.build(prob, K, X, U, pattern, vals, const, rk=rk)
    r = socp_solver.solve(mdl["c"], mdl["A"], mdl["b"], mdl["G"], mdl["h"], mdl["l"], mdl["q"], tol=SOLVER_TOL)
    return r, mdl


def _read_back(r, mdl, K):
    """What solve_step reads back (rocketland.jl:278-283): x, u, dsigma, nu, dx."""
    ix, x = mdl["idx"], r["x"]
    return {"x": x[ix["xv"]], "u": x[ix["uv"]], "dsig": x[ix["dsig"]], "nu": x[ix["nuv"]], "dx": x[ix["dxv"]]}


def test_solver_on_known_problems():
    from scipy.optimize import linprog
    rng = np.random.default_rng(0)
    n, p = 12, 5
    A = rng.normal(size=(p, n)); x0 = rng.uniform(0.5, 2, n); b = A @ x0; c = rng.uniform(0.1, 1, n)
    r = socp_solver.solve(c, A, b, -np.eye(n), np.zeros(n), n, [])
    ref = linprog(c, A_eq=A, b_eq=b, bounds=[(0, None)] * n, method="highs")
    assert r["status"] == "optimal" and abs(r["pcost"] - ref.fun) <= 1e-8 and np.abs(r["x"] - ref.x).max() <= 1e-7
    # min -x1  s.t. |(x1, x2)| <= 1, x2 = 1/2   ->  x1 = sqrt(3)/2
    G = np.array([[0, 0], [-1, 0], [0, -1.0]]); h = np.array([1.0, 0, 0])
    r = socp_solver.solve(np.array([-1.0, 0]), np.array([[0, 1.0]]), np.array([0.5]), G, h, 0, [3])
    assert r["status"] == "optimal" and r["x"] == pytest.approx([np.sqrt(0.75), 0.5], abs=1e-8)


def test_subproblem_of_the_sample_problem_solves_and_is_well_posed(prob_aero, oracle_tables):
    """The first SCvx subproblem of base_prob_aero_scaled (K=50, initial guess of linear_points, r_k = 100 as in
    create_initial rocketland.jl:38) from the ORACLE's matrices: the solver converges, the KKT residuals are at the
    tolerance, and a 1e-10 relative perturbation of the matrices (the parity bar) moves the solution far less than any
    solver tolerance — the computed amplification factor that ties matrix parity to solution parity."""
    K = prob_aero.K
    X, U, sigma, dt = workloads.sample_trajectory(prob_aero)
    blocks, err, tlb, _ = oracle.linearize_batch(ProbInfo(prob_aero), oracle_tables, X, U, sigma, dt, 10, 1)
    pattern = rocketland.socp_pattern(K + 1)

    def solve(bl):
        D = bl[0, :, 1:22, :].transpose(0, 2, 1)
        M, const = socp_assembly.assemble_dense(D, bl[0, :, 0, :], X[0], U[0], prob_aero.Tmin)
        return _solve(prob_aero, K, X[0], U[0], pattern, _vals_from_dense(M, K, pattern), const)
    r0, mdl = solve(blocks)
    assert r0["status"] == "optimal" and r0["pres"] <= SOLVER_TOL and r0["dres"] <= SOLVER_TOL
    sol = _read_back(r0, mdl, K)
    assert sol["x"][0, 0] == pytest.approx(prob_aero.mwet, abs=1e-8)                 # boundary condition rows
    assert sol["x"][7:11, K] == pytest.approx(prob_aero.qBIf, abs=1e-7)
    assert 0.0 < sol["dsig"] < 100.0
    rng = np.random.default_rng(1)
    r1, _ = solve(blocks * (1.0 + 1e-10 * rng.uniform(-1, 1, blocks.shape)))
    amp = np.abs(r1["x"] - r0["x"]).max() / (1e-10 * np.abs(r0["x"]).max())
    print(f"\n[SOCP sensitivity] 1e-10 relative data perturbation -> max |dx| = {np.abs(r1['x'] - r0['x']).max():.2e} "
          f"(amplification {amp:.1f})")
    assert np.abs(r1["x"] - r0["x"]).max() <= 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 0])
def test_socp_solutions_from_gpu_and_oracle_matrices_match(prob_aero, oracle_tables, mode):
    """Same subproblem assembled from (i) the CUDA path's matrices through scvx_socp_values_batch and (ii) the oracle's
    matrices: the read-back quantities of solve_step (x, u, dsigma, nu, dx) agree within the solver tolerance."""
    from successiveconvexification_b200 import dynamics as dyn
    K = prob_aero.K
    X, U, sigma, dt = workloads.sample_trajectory(prob_aero)
    cache = dyn.make_cache(prob_aero)
    blocks, err, tlb = dyn.linearize_batch(cache, X, U, sigma, dt, 10, mode)
    vals, const = rocketland.socp_values_batch(cache, blocks, err, tlb)
    pattern = rocketland.socp_pattern(K + 1)
    rg, mdl = _solve(prob_aero, K, X[0], U[0], pattern, vals[0], const[0])
    ref, rerr, rtlb, _ = oracle.linearize_batch(ProbInfo(prob_aero), oracle_tables, X, U, sigma, dt, 10, mode)
    D = ref[0, :, 1:22, :].transpose(0, 2, 1)
    M, cst = socp_assembly.assemble_dense(D, ref[0, :, 0, :], X[0], U[0], prob_aero.Tmin)
    ro, _ = _solve(prob_aero, K, X[0], U[0], pattern, _vals_from_dense(M, K, pattern), cst)
    assert rg["status"] == ro["status"] == "optimal"
    a, b = _read_back(rg, mdl, K), _read_back(ro, mdl, K)
    worst = max(float(np.abs(np.asarray(a[k]) - np.asarray(b[k])).max()) for k in a)
    print(f"\n[SOCP solution parity mode={mode}] max |x_gpu - x_oracle| over (x, u, dsigma, nu, dx) = {worst:.2e}; "
          f"objective {rg['pcost']:.12e} vs {ro['pcost']:.12e}")
    assert worst <= 1e-7                                     # Mosek's default relative tolerance is 1e-8 on an O(10) solution
    assert abs(rg["pcost"] - ro["pcost"]) <= 1e-7 * max(1.0, abs(ro["pcost"]))
