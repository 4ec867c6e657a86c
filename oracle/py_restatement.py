"""Second, independent restatement of the hot path in numpy — TEST INFRASTRUCTURE ONLY.

Purpose: pin oracle/scvx_oracle.cpp (forward-mode dual numbers, C++) by a different language and a
different differentiation method.  Here the Jacobian D = d rk4(inp) / d inp is taken by the
COMPLEX-STEP method (perturb one input by i*1e-30, read the imaginary part), which is exact to
round-off for the analytic arithmetic the path executes and agrees with forward-mode AD about
branches and clamps because every comparison / floor is taken on the real part.

Follows, in this order: reference dynamics.jl:112-134 (rk4), 108-110 (current_control), 54-77
(dx_static), 29-52 (DCM, Omega); aerodynamics.jl:38-58 (aero_force), 11-36 (tables);
Interpolations.jl cubic B-spline semantics as written out in SURVEY.md §8a-7.
Pure Python loops: only for a handful of intervals.  PARITY UNPINNED by the reference (no tests).
"""
from __future__ import annotations

import numpy as np


# ------------------------------------------------------------------ Interpolations.jl restatement
def prefilter_line_matrix(n: int) -> np.ndarray:
    M = np.zeros((n + 2, n + 2))
    M[0, :3] = [1.0, -2.0, 1.0]
    for k in range(1, n + 1):
        M[k, k - 1:k + 2] = [1 / 6, 2 / 3, 1 / 6]
    M[n + 1, n - 1:] = [1.0, -2.0, 1.0]
    return M


def prefilter(samples: np.ndarray) -> np.ndarray:
    """(n1, n2) samples -> (n1+2, n2+2) cubic B-spline coefficients, Line(OnGrid()) boundary."""
    n1, n2 = samples.shape
    rhs = np.zeros((n1 + 2, n2))
    rhs[1:-1, :] = samples
    c1 = np.linalg.solve(prefilter_line_matrix(n1), rhs)             # along axis 0
    rhs2 = np.zeros((n2 + 2, n1 + 2))
    rhs2[1:-1, :] = c1.T
    return np.linalg.solve(prefilter_line_matrix(n2), rhs2).T        # along axis 1


def _clamp(x, lo, hi):
    if x.real > hi:
        return complex(hi)
    if x.real < lo:
        return complex(lo)
    return x


def _weights(d):
    o = 1.0 - d
    return [o ** 3 / 6.0, 2 / 3 - d * d + d ** 3 / 2.0, 2 / 3 - o * o + o ** 3 / 2.0, d ** 3 / 6.0]


def spline_eval(coef, geom, x, y):
    n1, n2, x0, dx, y0, dy = int(geom[0]), int(geom[1]), geom[2], geom[3], geom[4], geom[5]
    xi = _clamp((x - x0) / dx + 1.0, 1.0, float(n1))
    yi = _clamp((y - y0) / dy + 1.0, 1.0, float(n2))
    i = min(int(np.floor(xi.real)), n1 - 1)
    j = min(int(np.floor(yi.real)), n2 - 1)
    wx, wy = _weights(xi - i), _weights(yi - j)
    acc = 0.0
    for b in range(4):
        for a in range(4):
            acc = acc + wx[a] * wy[b] * coef[i - 1 + a, j - 1 + b]
    return acc


# ------------------------------------------------------------------ dynamics.jl restatement
def DCM(q):
    q0, q1, q2, q3 = q
    return np.array([[1 - 2 * (q2 ** 2 + q3 ** 2), 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2)],
                     [2 * (q1 * q2 + q0 * q3), 1 - 2 * (q1 ** 2 + q3 ** 2), 2 * (q2 * q3 - q0 * q1)],
                     [2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), 1 - 2 * (q1 ** 2 + q2 ** 2)]])


def Omega(w):
    return np.array([[0, -w[0], -w[1], -w[2]],
                     [w[0], 0, w[2], -w[1]],
                     [w[1], -w[2], 0, w[0]],
                     [w[2], w[1], -w[0], 0]])


def _norm(a):
    return np.sqrt(np.sum(a * a))        # NOT abs(): must stay analytic for the complex step


def aero_force(P, tables, bv, vel):
    nv = _norm(vel)
    dp = np.sum(bv * vel) / nv
    cos_aoa = _clamp(complex(dp / _norm(bv)), -1.0, 1.0)
    mach = nv / P["sos"]
    drag = spline_eval(tables["drag"], tables["geom"], cos_aoa, complex(mach)) * P["force_scalar"]
    if abs(dp.real) >= 0.95:
        return drag * vel / nv
    lift = spline_eval(tables["lift"], tables["geom"], cos_aoa, complex(mach)) * P["force_scalar"]
    trqd = np.cross(vel, bv)
    liftd = np.cross(-trqd, vel)
    liftd = liftd / _norm(liftd)
    return drag * vel / nv + lift * liftd


def dx_static(P, tables, x, u, mult):
    q, w, v = x[7:11], x[11:14], x[4:7]
    C = DCM(q)
    aerf = aero_force(P, tables, C @ np.array([1.0, 0.0, 0.0]), v) if P["aero_kind"] == 1 else np.zeros(3)
    acc = (C @ u + aerf) / x[0]
    rot_vel = 0.5 * (Omega(w) @ q)
    rot_acc = P["jBi"] @ (np.cross(P["rTB"], u) - np.cross(w, P["jB"] @ w))
    return np.concatenate([[-P["a"] * np.sqrt(np.sum(u * u))], v, [acc[0] - P["g0"], acc[1], acc[2]],
                           rot_vel, rot_acc]) * mult


def rk4(P, tables, inp, dt, npts=10, mode=0):
    inp = np.asarray(inp, dtype=complex)
    state = inp[:14].copy()
    su, eu = inp[14:17], inp[17:20]
    idt = dt / npts
    pcs = 1.0 / npts
    pca = 0.0
    s = 1.0 if mode == 0 else idt
    cc = lambda pc: (1.0 - pc) * su + pc * eu
    for _ in range(npts):
        ict, mct, ect = cc(pca), cc(pca + pcs / 2), cc(pca + pcs)
        k1 = dx_static(P, tables, state, ict, inp[20])
        k2 = dx_static(P, tables, state + s * k1 / 2, mct, inp[20])
        k3 = dx_static(P, tables, state + s * k2 / 2, mct, inp[20])
        k4 = dx_static(P, tables, state + s * k3, ect, inp[20])
        pca += pcs
        state = state + idt * (k1 / 6 + k2 / 3 + k3 / 3 + k4 / 6)
    return state


def linearize_interval(P, tables, inp, dt, npts=10, mode=0, h=1e-30):
    """-> (endpoint 14, D 14x21, z 14) by complex step."""
    inp = np.asarray(inp, dtype=np.float64)
    endpoint = rk4(P, tables, inp, dt, npts, mode).real
    D = np.zeros((14, 21))
    for c in range(21):
        pert = inp.astype(complex)
        pert[c] += 1j * h
        D[:, c] = rk4(P, tables, pert, dt, npts, mode).imag / h
    return endpoint, D, endpoint - D @ inp


def probinfo_dict(info) -> dict:
    """ProbInfo (successiveconvexification_b200.defns) -> plain dict used above."""
    fs = getattr(info.aero, "force_scalar", 0.0)
    return dict(a=info.a, g0=info.g0, sos=info.sos, jB=np.asarray(info.jB), jBi=np.asarray(info.jBi),
                rTB=np.asarray(info.rTB), force_scalar=fs, aero_kind=info.aero_kind)


def tables_dict(aero) -> dict:
    d = aero.drag_itrp
    return dict(drag=prefilter(np.asarray(aero.drag_itrp.samples)), lift=prefilter(np.asarray(aero.lift_itrp.samples)),
                geom=[d.samples.shape[0], d.samples.shape[1], d.cos0, d.dcos, d.mach0, d.dmach])
