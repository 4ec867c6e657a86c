"""The SCvx subproblem of the reference in conic standard form — TEST INFRASTRUCTURE ONLY.

A numpy/scipy restatement of `Rocketland.build_model` (rocketland.jl:53-219) with the per-iteration data of
`solve_step` (rocketland.jl:245-269):

    minimise  -xv[1,K+1] + wNu*Jvnu + 0.5*Jtr + Jsig                                   (rocketland.jl:84-86)
    s.t.      dxv - xv + xbar = 0,  duv - uv + ubar = 0                                 (92-97, 245-248)
              |nuv| <= Jvnu,  |[dxv; duv]| <= Jtr,  |dsig| <= Jsig                       (100-102)
              boundary conditions on xv[:,1], xv[:,K+1], uv[2:3,K+1]                    (109-115)
              D_n [dxv_n; duv_n; duv_{n+1}; dsig] + nuv_{n+1} - dxv_{n+1} + lin_err_n = 0   (117-133, 251-258)
              mdry <= xv[1,k]; glide slope, tilt and rate cones; thrust cones            (137-192)
              H_n duv_n + (Tmin - |ubar_n|) <= 0                                         (194-201, 260-265)
              Jtr - r_k <= 0                                                             (215-216, 269)

in the form  min c'x  s.t.  A x = b,  G x + s = h,  s in R+^l x Q^{q_1} x ...  (oracle/socp_solver.py).  MOI's
`f(x) + const in Zeros / Nonpositives / Nonnegatives` become `A x = -const`, `G x + s = -const`.

The trajectory-dependent rows (dynamics equalities and thrust lower bound) are NOT rebuilt here: they are taken from
the fixed-pattern sparse rows of the product (`scvx_socp_pattern` + the value / constant arrays of
`scvx_socp_values_batch`), whose local column j is the reference's variable 17(K+1) + j.  Everything else of
`build_model` is trajectory-independent structure.

Variable order = the reference's creation order (rocketland.jl:71-81, 142, 155, 163, 184): xv, uv, dxv, duv, dsig,
nuv, Jvnu, Jtr, Jsig, gshelp(K), aoa_help(K), ang_sp_help(K), mtk(K+1).  The reference also creates `rK`
(rocketland.jl:215), a variable that enters no constraint and no cost; it is left out.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def variable_index(K: int):
    n = K + 1
    o = {}
    o["xv"] = np.arange(14 * n).reshape(n, 14).T                       # xv[j, k] = 14 k + j
    o["uv"] = 14 * n + np.arange(3 * n).reshape(n, 3).T
    o["local0"] = 17 * n                                               # first local column (dxv[0,0])
    o["dxv"] = 17 * n + np.arange(14 * n).reshape(n, 14).T
    o["duv"] = 31 * n + np.arange(3 * n).reshape(n, 3).T
    o["dsig"] = 34 * n
    o["nuv"] = 34 * n + 1 + np.arange(14 * n).reshape(n, 14).T
    p = 48 * n + 1
    o["Jvnu"], o["Jtr"], o["Jsig"] = p, p + 1, p + 2
    o["gshelp"] = p + 3 + np.arange(K)
    o["aoa_help"] = p + 3 + K + np.arange(K)
    o["ang_sp_help"] = p + 3 + 2 * K + np.arange(K)
    o["mtk"] = p + 3 + 3 * K + np.arange(n)
    o["n_vars"] = p + 3 + 3 * K + n
    return o


def build(prob, K, X, U, pattern, vals, const, rk=100.0):
    """prob: DescentProblem (normalised); X (K+1, 14), U (K+1, 3): the trajectory linearised about;
    pattern = (n_rows, n_cols, colptr, rowind) of scvx_socp_pattern; vals (nnz,), const (n_rows,) of one trajectory.
    -> dict(c, A, b, G, h, l, q, idx)."""
    n = K + 1
    ix = variable_index(K)
    nv = ix["n_vars"]
    nr, nc, colptr, rowind = pattern
    M = sp.csc_matrix((vals, rowind, colptr), shape=(nr, nc)).tocsr()    # local columns
    Mfull = sp.hstack([sp.csr_matrix((nr, ix["local0"])), M, sp.csr_matrix((nr, nv - ix["local0"] - nc))]).tocsr()

    tggs = np.tan(np.deg2rad(prob.gammaGs))
    sqcm = np.sqrt((1.0 - np.cos(np.deg2rad(prob.thetaMax))) / 2.0)
    delMax = np.cos(np.deg2rad(prob.deltaMax))

    c = np.zeros(nv)
    c[ix["xv"][0, K]] = -1.0
    c[ix["Jvnu"]], c[ix["Jtr"]], c[ix["Jsig"]] = prob.wNu, 0.5, 1.0

    rows, cols, data, b = [], [], [], []

    def eq(terms, rhs):
        r = len(b)
        for col, v in terms:
            rows.append(r); cols.append(int(col)); data.append(float(v))
        b.append(float(rhs))

    # state_base / control_base (rocketland.jl:92-97): dxv - xv + xbar = 0
    for k in range(n):
        for j in range(14):
            eq([(ix["dxv"][j, k], 1.0), (ix["xv"][j, k], -1.0)], -X[k, j])
    for k in range(n):
        for j in range(3):
            eq([(ix["duv"][j, k], 1.0), (ix["uv"][j, k], -1.0)], -U[k, j])
    # boundary conditions (rocketland.jl:109-115)
    xv, uv = ix["xv"], ix["uv"]
    bc_vars = [xv[0, 0], *xv[1:4, 0], *xv[4:7, 0], *xv[11:14, 0], *xv[1:4, K], *xv[4:7, K], *xv[7:11, K], *xv[11:14, K],
               uv[1, K], uv[2, K]]
    bc_vals = np.concatenate([[prob.mwet], prob.rIi, prob.vIi, prob.wBi, prob.rIf, prob.vIf, prob.qBIf, prob.wBf, [0.0, 0.0]])
    for v, val in zip(bc_vars, bc_vals):
        eq([(v, 1.0)], val)
    A1 = sp.csr_matrix((data, (rows, cols)), shape=(len(b), nv))
    b1 = np.array(b)
    # dynamics rows of the product: M x_local + lin_err = 0
    A2, b2 = Mfull[:14 * K], -np.asarray(const[:14 * K])
    rows, cols, data, b = [], [], [], []
    for k in range(K):                                                   # glide slope helper (rocketland.jl:142-144)
        eq([(ix["gshelp"][k], 1.0), (xv[1, k], -1.0 / tggs)], 0.0)
    for k in range(K):                                                   # tilt helper (155-156)
        eq([(ix["aoa_help"][k], 1.0)], sqcm)
    for k in range(K):                                                   # rate helper (163-164)
        eq([(ix["ang_sp_help"][k], 1.0)], prob.omMax)
    A3 = sp.csr_matrix((data, (rows, cols)), shape=(len(b), nv))
    b3 = np.array(b)
    A = sp.vstack([A1, A2, A3]).tocsr()
    bb = np.concatenate([b1, b2, b3])

    # ---- G x + s = h
    grow, gcol, gdat, h = [], [], [], []

    def ineq(terms, rhs):
        r = len(h)
        for col, v in terms:
            grow.append(r); gcol.append(int(col)); gdat.append(float(v))
        h.append(float(rhs))

    for k in range(1, n):                                                # mdry <= xv[1,k]  (137)
        ineq([(xv[0, k], -1.0)], -prob.mdry)
    for k in range(n):                                                   # mtk <= Tmax  (186)
        ineq([(ix["mtk"][k], 1.0)], prob.Tmax)
    for k in range(n):                                                   # mtk <= uv[1,k] / cos(deltaMax)  (188)
        ineq([(ix["mtk"][k], 1.0), (uv[0, k], -1.0 / delMax)], 0.0)
    n_lin_a = len(h)
    G_tlb, h_tlb = Mfull[14 * K:], -np.asarray(const[14 * K:])          # thrust lower bound rows of the product (194-201)
    ineq_rk = ([(ix["Jtr"], 1.0)], rk)                                   # Jtr - r_k <= 0  (216, 269)
    Ga = sp.csr_matrix((gdat, (grow, gcol)), shape=(n_lin_a, nv))
    Grk = sp.csr_matrix(([1.0], ([0], [ix["Jtr"]])), shape=(1, nv))
    l = n_lin_a + (K + 1) + 1
    h_lin = np.concatenate([np.array(h), h_tlb, [rk]])

    cones = [[ix["Jvnu"], *ix["nuv"].T.reshape(-1)],                                         # |nuv| <= Jvnu  (100)
             [ix["Jtr"], *ix["dxv"].T.reshape(-1), *ix["duv"].T.reshape(-1)],               # trust region (101)
             [ix["Jsig"], ix["dsig"]]]                                                       # (102)
    cones += [[ix["gshelp"][k], xv[2, k], xv[3, k]] for k in range(K)]                       # glide slope (146-148)
    cones += [[ix["aoa_help"][k], xv[9, k], xv[10, k]] for k in range(K)]                    # tilt: qbi[3:4] (158-160)
    cones += [[ix["ang_sp_help"][k], xv[11, k], xv[12, k], xv[13, k]] for k in range(K)]     # rate (165-167)
    cones += [[ix["mtk"][k], uv[0, k], uv[1, k], uv[2, k]] for k in range(n)]                # thrust (190-192)
    q = [len(cn) for cn in cones]
    flat = np.concatenate([np.asarray(cn, dtype=int) for cn in cones])
    Gc = sp.csr_matrix((-np.ones(flat.size), (np.arange(flat.size), flat)), shape=(flat.size, nv))
    G = sp.vstack([Ga, G_tlb, Grk, Gc]).tocsr()
    hh = np.concatenate([h_lin, np.zeros(flat.size)])
    return {"c": c, "A": A, "b": bb, "G": G, "h": hh, "l": l, "q": q, "idx": ix}
