"""Third restatement of the hot path, at 50 significant digits (mpmath) — TEST INFRASTRUCTURE ONLY.

Purpose: measure how far FP64 can resolve the entries of D at all.  The map rk4(inp) (reference dynamics.jl:112-134,
54-77, 29-52, 108-110; aerodynamics.jl:38-58; Interpolations.jl semantics as in SURVEY.md §8a-7) is evaluated in
50-digit arithmetic and differentiated by central differences with a 1e-25 step (error ~1e-48), i.e. to far more
digits than a double holds.  The FP64 oracle must agree with it to rounding; the residual is what defines the floor of
the parity metric in tests/conftest.py.  The spline coefficients are the FP64 prefiltered arrays (exactly
representable inputs), so only the evaluation is high precision.  Plain Python loops: one or two intervals.
PARITY UNPINNED by the reference (it has no tests for this path and cannot run here).
"""
import mpmath as mp

mp.mp.dps = 50
F = mp.mpf


def _dot(a, b):
    return sum(x * y for x, y in zip(a, b))


def _cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def _weights(d):
    o = 1 - d
    return [o ** 3 / 6, F(2) / 3 - d * d + d ** 3 / 2, F(2) / 3 - o * o + o ** 3 / 2, d ** 3 / 6]


def spline_eval(coef, geom, x, y):
    n1, n2 = int(geom[0]), int(geom[1])
    x0, dx, y0, dy = (F(float(g)) for g in geom[2:6])
    xi = min(max((x - x0) / dx + 1, F(1)), F(n1))
    yi = min(max((y - y0) / dy + 1, F(1)), F(n2))
    i = min(int(mp.floor(xi)), n1 - 1)
    j = min(int(mp.floor(yi)), n2 - 1)
    wx, wy = _weights(xi - i), _weights(yi - j)
    return sum(wx[a] * wy[b] * F(float(coef[i - 1 + a, j - 1 + b])) for a in range(4) for b in range(4))


def dcm(q):
    q0, q1, q2, q3 = q
    return [[1 - 2 * (q2 * q2 + q3 * q3), 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2)],
            [2 * (q1 * q2 + q0 * q3), 1 - 2 * (q1 * q1 + q3 * q3), 2 * (q2 * q3 - q0 * q1)],
            [2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), 1 - 2 * (q1 * q1 + q2 * q2)]]


def aero_force(P, tables, bv, v):
    nv = mp.sqrt(_dot(v, v))
    dp = _dot(bv, v) / nv
    ca = min(max(dp / mp.sqrt(_dot(bv, bv)), F(-1)), F(1))
    mach = nv / P["sos"]
    drag = spline_eval(tables["drag"], tables["geom"], ca, mach) * P["force_scalar"]
    out = [drag * c / nv for c in v]
    if abs(dp) >= F("0.95"):
        return out
    lift = spline_eval(tables["lift"], tables["geom"], ca, mach) * P["force_scalar"]
    ld = _cross([-c for c in _cross(v, bv)], v)
    nl = mp.sqrt(_dot(ld, ld))
    return [o + lift * c / nl for o, c in zip(out, ld)]


def dx_static(P, tables, x, u, mult):
    m, v, q, w = x[0], x[4:7], x[7:11], x[11:14]
    C = dcm(q)
    F_a = aero_force(P, tables, [C[0][0], C[1][0], C[2][0]], v) if P["aero_kind"] == 1 else [F(0)] * 3
    acc = [(_dot(C[r], u) + F_a[r]) / m for r in range(3)]
    qd = [(-w[0] * q[1] - w[1] * q[2] - w[2] * q[3]) / 2, (w[0] * q[0] + w[2] * q[2] - w[1] * q[3]) / 2,
          (w[1] * q[0] - w[2] * q[1] + w[0] * q[3]) / 2, (w[2] * q[0] + w[1] * q[1] - w[0] * q[2]) / 2]
    jw = [_dot(P["jB"][r], w) for r in range(3)]
    tq = [a - b for a, b in zip(_cross(P["rTB"], u), _cross(w, jw))]
    wd = [_dot(P["jBi"][r], tq) for r in range(3)]
    f = [-P["a"] * mp.sqrt(_dot(u, u))] + list(v) + [acc[0] - P["g0"], acc[1], acc[2]] + qd + wd
    return [c * mult for c in f]


def rk4(P, tables, inp, dt, npts=10, mode=0):
    x = list(inp[:14])
    su, eu, sig = inp[14:17], inp[17:20], inp[20]
    h = F(dt) / npts
    s = F(1) if mode == 0 else h
    pcs = F(1) / npts
    pca = F(0)
    cc = lambda pc: [(1 - pc) * a + pc * b for a, b in zip(su, eu)]
    ax = lambda base, k, c: [b + c * kk for b, kk in zip(base, k)]
    for _ in range(npts):
        k1 = dx_static(P, tables, x, cc(pca), sig)
        k2 = dx_static(P, tables, ax(x, k1, s / 2), cc(pca + pcs / 2), sig)
        k3 = dx_static(P, tables, ax(x, k2, s / 2), cc(pca + pcs / 2), sig)
        k4 = dx_static(P, tables, ax(x, k3, s), cc(pca + pcs), sig)
        pca += pcs
        x = [xx + h * (a / 6 + b / 3 + c / 3 + d / 6) for xx, a, b, c, d in zip(x, k1, k2, k3, k4)]
    return x


def linearize_interval(P, tables, inp, dt, npts=10, mode=0, step="1e-25"):
    """-> (endpoint, D) as nested lists of mpf; D[r][c] by central differences in 50-digit arithmetic."""
    inp = [F(float(v)) for v in inp]
    e = rk4(P, tables, inp, dt, npts, mode)
    hh = F(step)
    cols = []
    for c in range(21):
        up, dn = list(inp), list(inp)
        up[c] += hh
        dn[c] -= hh
        a, b = rk4(P, tables, up, dt, npts, mode), rk4(P, tables, dn, dt, npts, mode)
        cols.append([(p - q) / (2 * hh) for p, q in zip(a, b)])
    return e, [[cols[c][r] for c in range(21)] for r in range(14)]


def probinfo_mp(info) -> dict:
    import numpy as np
    conv = lambda a: [[F(float(v)) for v in row] for row in np.asarray(a)]
    return dict(a=F(float(info.a)), g0=F(float(info.g0)), sos=F(float(info.sos)), jB=conv(info.jB), jBi=conv(info.jBi),
                rTB=[F(float(v)) for v in info.rTB], force_scalar=F(float(getattr(info.aero, "force_scalar", 0.0))),
                aero_kind=info.aero_kind)
