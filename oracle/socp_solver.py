"""A small primal-dual interior-point solver for second-order cone programs — TEST INFRASTRUCTURE ONLY.

The reference solves its SCvx subproblem with Mosek through MathOptInterface (rocketland.jl:58-59, 271-276); neither is
available here (and Mosek needs a licence), so the SOCP-solution parity test (BASELINE.json north_star: "the resulting
SOCP solutions must match to within the solver tolerance") uses this solver on both data sets.  It is the textbook
path-following method with Nesterov-Todd scaling and Mehrotra's predictor-corrector for

    minimise  c'x   subject to   A x = b,   G x + s = h,   s in K = R+^l x Q^{q_1} x ... x Q^{q_N}

(the standard form of ECOS / CVXOPT `conelp`; Q^q = {(t, u): |u|_2 <= t}).  Dense cone blocks are avoided: the squared
NT scaling of a second-order cone is diagonal plus rank one, W^2 = beta^2 (2 w w' - J), and the rank-one part is carried
by one auxiliary unknown per cone, so every iteration is one sparse LU of the KKT matrix.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class Cones:
    def __init__(self, l: int, q):
        self.l, self.q = int(l), [int(v) for v in q]
        self.m = self.l + sum(self.q)
        self.degree = self.l + len(self.q)
        self.starts = np.concatenate([[self.l], self.l + np.cumsum(self.q)])[:-1].astype(int) if self.q else np.zeros(0, int)
        # vectorised segment bookkeeping for the second-order cones
        self.seg = np.repeat(np.arange(len(self.q)), self.q)                     # cone id of every SOC entry
        self.head = np.zeros(self.m - self.l, dtype=bool)
        self.head[(self.starts - self.l)] = True

    # ---- Jordan algebra on the product cone (SOC parts vectorised with reduceat)
    def _tail_dot(self, u, v):
        """per cone: u1'v1 (tails)"""
        uu, vv = u[self.l:], v[self.l:]
        p = np.where(self.head, 0.0, uu * vv)
        return np.add.reduceat(p, self.starts - self.l) if self.q else np.zeros(0)

    def identity(self):
        e = np.zeros(self.m)
        e[:self.l] = 1.0
        e[self.starts] = 1.0
        return e

    def det(self, u):
        return u[self.starts] ** 2 - self._tail_dot(u, u)

    def interior(self, u):
        ok = np.all(u[:self.l] > 0)
        if self.q:
            ok = ok and np.all(u[self.starts] > 0) and np.all(self.det(u) > 0)
        return bool(ok)

    def prod(self, u, v):
        out = np.empty(self.m)
        out[:self.l] = u[:self.l] * v[:self.l]
        if self.q:
            u0, v0 = u[self.starts], v[self.starts]
            full = np.add.reduceat(u[self.l:] * v[self.l:], self.starts - self.l)       # u'v per cone
            out[self.l:] = np.repeat(u0, self.q) * v[self.l:] + np.repeat(v0, self.q) * u[self.l:]
            out[self.starts] = full
        return out

    def inv_prod(self, lam, u):
        """x with lam o x = u."""
        out = np.empty(self.m)
        out[:self.l] = u[:self.l] / lam[:self.l]
        if self.q:
            l0, u0 = lam[self.starts], u[self.starts]
            d = self.det(lam)
            lu = self._tail_dot(lam, u)                                                  # lam1'u1
            x0 = (l0 * u0 - lu) / d
            c1 = (lu / l0 - u0) / d                                                      # coefficient of lam1 in x1
            out[self.l:] = np.repeat(c1, self.q) * lam[self.l:] + u[self.l:] / np.repeat(l0, self.q)
            out[self.starts] = x0
        return out

    def max_step(self, lam, d):
        """largest alpha with lam + alpha d in the cone (inf if unbounded)."""
        a = np.inf
        neg = d[:self.l] < 0
        if neg.any():
            a = min(a, float(np.min(-lam[:self.l][neg] / d[:self.l][neg])))
        if self.q:
            # per cone: (l0 + a d0)^2 - |l1 + a d1|^2 >= 0 and l0 + a d0 >= 0
            l0, d0 = lam[self.starts], d[self.starts]
            aa = d0 ** 2 - self._tail_dot(d, d)
            bb = 2.0 * (l0 * d0 - self._tail_dot(lam, d))
            cc = self.det(lam)
            for A2, B2, C2, L0, D0 in zip(aa, bb, cc, l0, d0):
                roots = []
                if abs(A2) < 1e-300:
                    if B2 < 0:
                        roots.append(-C2 / B2)
                else:
                    disc = B2 * B2 - 4 * A2 * C2
                    if disc >= 0:
                        sq = np.sqrt(disc)
                        roots += [(-B2 - sq) / (2 * A2), (-B2 + sq) / (2 * A2)]
                if D0 < 0:
                    roots.append(-L0 / D0)
                pos = [r for r in roots if r > 0]
                if pos:
                    a = min(a, min(pos))
        return a


class NTScaling:
    """Nesterov-Todd scaling W of (s, z):  lambda = W^{-T} s = W z."""

    def __init__(self, K: Cones, s, z):
        self.K = K
        l = K.l
        self.d = np.sqrt(s[:l] / z[:l])                       # LP part: W = diag(d)
        if K.q:
            ds, dz = np.sqrt(K.det(s)), np.sqrt(K.det(z))
            sb, zb = s[l:] / np.repeat(ds, K.q), z[l:] / np.repeat(dz, K.q)
            # gamma = sqrt((1 + zb'sb)/2)
            g = np.sqrt(0.5 * (1.0 + np.add.reduceat(sb * zb, K.starts - l)))
            Jz = np.where(K.head, zb, -zb)
            self.w = (sb + Jz) / np.repeat(2.0 * g, K.q)      # NT point (det = 1), per cone
            self.beta = np.sqrt(ds / dz)
        else:
            self.w, self.beta = np.zeros(0), np.zeros(0)

    def _apply(self, u, inverse):
        K, l = self.K, self.K.l
        out = np.empty(K.m)
        out[:l] = u[:l] / self.d if inverse else u[:l] * self.d
        if K.q:
            w, uu = self.w, u[l:]
            w0 = w[K.starts - l]
            wu_tail = np.add.reduceat(np.where(K.head, 0.0, w * uu), K.starts - l)
            u0 = uu[K.starts - l]
            sign = -1.0 if inverse else 1.0
            # W u / beta = [w0 u0 + sign w1'u1 ; u1 + (sign u0 + w1'u1/(1+w0)) w1]
            o0 = w0 * u0 + sign * wu_tail
            coef = sign * u0 + wu_tail / (1.0 + w0)
            o = uu + np.repeat(coef, K.q) * w
            o[K.starts - l] = o0
            scale = np.repeat(1.0 / self.beta if inverse else self.beta, K.q)
            out[l:] = o * scale
        return out

    def W(self, u):
        return self._apply(u, False)

    def Winv(self, u):
        return self._apply(u, True)


def _kkt_matrix(A, G, K: Cones, scal: NTScaling):
    """[[0 A' G' 0]; [A 0 0 0]; [G 0 -D U]; [0 0 U' I_N]] with  W^2 = -(-D) + ...: for a SOC,
    W^2 = beta^2 (2 w w' - J) = -beta^2 J' ... carried as  -W^2 dz = beta^2 J dz - u (u'dz), u = sqrt(2) beta w, with one
    auxiliary unknown t = u'dz per cone."""
    n, p, m, N = A.shape[1], A.shape[0], K.m, len(K.q)
    diag = np.empty(m)
    diag[:K.l] = -scal.d ** 2
    if N:
        b2 = np.repeat(scal.beta ** 2, K.q)
        diag[K.l:] = np.where(K.head, b2, -b2)                 # beta^2 J
        u = np.sqrt(2.0) * np.repeat(scal.beta, K.q) * scal.w
        U = sp.csr_matrix((-u, (np.arange(K.l, m), K.seg)), shape=(m, N))
        blocks = [[None, A.T, G.T, None], [A, None, None, None], [G, None, sp.diags(diag), U],
                  [None, None, U.T, sp.identity(N)]]
    else:
        blocks = [[None, A.T, G.T], [A, None, None], [G, None, sp.diags(diag)]]
    return sp.bmat(blocks, format="csc")


def solve(c, A, b, G, h, l, q, tol=1e-9, max_iter=80, verbose=False):
    """-> dict(x, s, y, z, status, iterations, gap, pres, dres, pcost, dcost)."""
    c, b, h = (np.asarray(v, dtype=float) for v in (c, b, h))
    A, G = sp.csr_matrix(A), sp.csr_matrix(G)
    K = Cones(l, q)
    n, p, m, N = A.shape[1], A.shape[0], K.m, len(K.q)
    assert G.shape == (m, n) and A.shape[1] == n
    e = K.identity()
    reg = 1e-9                       # static regularisation of the zero blocks (quasi-definite KKT matrix)

    def factor(scal):
        M = _kkt_matrix(A, G, K, scal).tolil()
        M.setdiag(M.diagonal() + np.concatenate([np.full(n, reg), np.full(p, -reg), np.zeros(m + N)]))
        lu = spla.splu(M.tocsc())
        Mc = M.tocsc()

        def kkt_solve(rx, ry, rz):
            rhs = np.concatenate([rx, ry, rz, np.zeros(N)])
            sol = lu.solve(rhs)
            for _ in range(2):                                 # iterative refinement against the regularisation
                sol = sol + lu.solve(rhs - Mc @ sol)
            return sol[:n], sol[n:n + p], sol[n + p:n + p + m]
        return kkt_solve

    # ---- initial point (CVXOPT conelp): W = I
    class _Id:
        d = np.ones(K.l); w = np.zeros(m - K.l); beta = np.ones(N)
    idscal = _Id()
    if N:
        idscal.w[K.starts - K.l] = 1.0
    kkt = factor(idscal)
    x, y, ms = kkt(np.zeros(n), b, h)            # minimise |s|^2 s.t. Ax = b, Gx + s = h   (z block returns -s)
    s = -ms
    _, y, z = kkt(-c, np.zeros(p), np.zeros(m))  # minimise |z|^2 s.t. G'z + A'y + c = 0

    def shift(v):
        a = K.max_step(v, e)                      # v + a e on the boundary?  use the classical shift instead
        # smallest t with v + t e in the interior
        t = 0.0
        if K.l:
            t = max(t, -float(v[:K.l].min()))
        if N:
            tails = np.sqrt(np.maximum(K._tail_dot(v, v), 0.0))
            t = max(t, float((tails - v[K.starts]).max()))
        return v + (1.0 + t) * e if t >= 0 or not K.interior(v) else v
    if not K.interior(s):
        s = shift(s)
    if not K.interior(z):
        z = shift(z)

    nrm_b, nrm_c, nrm_h = max(1.0, np.linalg.norm(b)), max(1.0, np.linalg.norm(c)), max(1.0, np.linalg.norm(h))
    status, it = "max_iter", 0
    for it in range(max_iter):
        rx = A.T @ y + G.T @ z + c
        ry = A @ x - b
        rz = G @ x + s - h
        gap = float(s @ z)
        pcost, dcost = float(c @ x), float(-(b @ y) - (h @ z))
        pres = max(np.linalg.norm(ry) / nrm_b, np.linalg.norm(rz) / nrm_h)
        dres = np.linalg.norm(rx) / nrm_c
        relgap = gap / max(1e-300, abs(pcost)) if pcost != 0 else gap
        if verbose:
            print(f"{it:3d} pcost {pcost: .9e} dcost {dcost: .9e} gap {gap:.2e} pres {pres:.2e} dres {dres:.2e}")
        if pres <= tol and dres <= tol and (gap <= tol or relgap <= tol):
            status = "optimal"
            break
        scal = NTScaling(K, s, z)
        lam = scal.W(z)
        kkt = factor(scal)
        mu = gap / K.degree

        def direction(ds, fac):
            # [0 A' G'; A 0 0; G 0 -W'W] [dx; dy; dz] = [-fac rx; -fac ry; -fac rz - W'(lam <> ds)],  Ds = W'(lam <> ds - W dz)
            t = K.inv_prod(lam, ds)
            dx, dy, dz = kkt(-fac * rx, -fac * ry, -fac * rz - scal.W(t))
            dsv = scal.W(t - scal.W(dz))
            return dx, dy, dz, dsv

        # predictor
        dxa, dya, dza, dsa = direction(-K.prod(lam, lam), 1.0)
        ts, tz = scal.Winv(dsa), scal.W(dza)
        alpha = min(1.0, K.max_step(lam, ts), K.max_step(lam, tz))
        sigma = (1.0 - alpha) ** 3
        # corrector
        ds = sigma * mu * e - K.prod(lam, lam) - K.prod(ts, tz)
        dx, dy, dz, dsv = direction(ds, 1.0 - sigma)
        ts, tz = scal.Winv(dsv), scal.W(dz)
        alpha = min(1.0, 0.99 * min(K.max_step(lam, ts), K.max_step(lam, tz)))
        x, y, z, s = x + alpha * dx, y + alpha * dy, z + alpha * dz, s + alpha * dsv
        if not (K.interior(s) and K.interior(z)):
            status = "left_cone"
            break
    return {"x": x, "s": s, "y": y, "z": z, "status": status, "iterations": it, "gap": gap, "pres": pres, "dres": dres,
            "pcost": pcost, "dcost": dcost}
