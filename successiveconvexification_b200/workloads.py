"""Synthetic Monte-Carlo / sweep batches of the named BASELINE configurations.

Host-side input construction only (numpy, PCG64 via `numpy.random.default_rng`): the same arrays feed
the CUDA path, the CPU checker used by the tests and — where available — the Julia reference.
"""
from __future__ import annotations

import ctypes

import numpy as np

from .defns import AERO_EXO, AERO_TABLE, AtmosphericData, CProbInfo, DescentProblem, ProbInfo

# numpy view of scvx_probinfo (include/scvx_b200.h)
PROBINFO_DTYPE = np.dtype([("a", "f8"), ("g0", "f8"), ("sos", "f8"), ("jB", "f8", (9,)), ("jBi", "f8", (9,)),
                           ("rTB", "f8", (3,)), ("rFB", "f8", (3,)), ("force_scalar", "f8"),
                           ("length_scalar", "f8"), ("Tmin", "f8"), ("aero_kind", "i4"), ("_pad", "i4")])
assert PROBINFO_DTYPE.itemsize == ctypes.sizeof(CProbInfo)


def probinfo_array(info: ProbInfo, n: int = 1) -> np.ndarray:
    """n identical records as a structured array (bit-identical to `info.to_c()`)."""
    rec = np.frombuffer(bytes(info.to_c()), dtype=PROBINFO_DTYPE)
    return np.repeat(rec, n)


def as_c_params(arr: np.ndarray):
    arr = np.ascontiguousarray(arr)
    return ctypes.cast(arr.ctypes.data, ctypes.POINTER(CProbInfo)), arr.shape[0], arr


def _quat_mul(a, b):
    """Hamilton product, scalar first, vectorised over leading axes."""
    aw, ax, ay, az = np.moveaxis(a, -1, 0)
    bw, bx, by, bz = np.moveaxis(b, -1, 0)
    return np.stack([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw], axis=-1)


def _rotation_between_x(v):
    """rotation_between([1,0,0], v) vectorised (Rotations.jl, as used at initial_solve.jl:121)."""
    nv = np.linalg.norm(v, axis=-1)
    w = nv + v[..., 0]
    axis = np.stack([np.zeros_like(nv), -v[..., 2], v[..., 1]], axis=-1)    # [1,0,0] x v
    anti = np.abs(w) < 100 * np.finfo(np.float64).eps
    axis[anti] = np.array([0.0, 0.0, 1.0])
    q = np.concatenate([w[..., None], axis], axis=-1)
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def linear_points_batch(prob: DescentProblem, K: int, rIi, vIi, mwet=None):
    """Vectorised `linear_points` (initial_solve.jl:113-129) for B perturbed initial conditions.
    -> X (B, K+1, 14), U (B, K+1, 3)."""
    rIi = np.asarray(rIi, dtype=np.float64)
    vIi = np.asarray(vIi, dtype=np.float64)
    B = rIi.shape[0]
    mwet = np.full(B, prob.mwet) if mwet is None else np.asarray(mwet, dtype=np.float64)
    k = np.arange(K + 1, dtype=np.float64)
    a = ((K - k) / K)[None, :, None]
    b = (k / K)[None, :, None]
    m = a[..., 0] * mwet[:, None] + b[..., 0] * prob.mdry
    r = a * rIi[:, None, :] + b * prob.rIf[None, None, :]
    v = a * vIi[:, None, :] + b * prob.vIf[None, None, :]
    q = _rotation_between_x(-v)
    X = np.concatenate([m[..., None], r, v, q, np.zeros((B, K + 1, 3))], axis=-1)
    U = np.zeros((B, K + 1, 3))
    U[..., 0] = m * prob.g
    return X, U


def monte_carlo_batch(prob: DescentProblem, K: int, B: int, seed: int, shard: int = 0, sweep: bool = False,
                      sigma_range=(1.0, 15.0)):
    """Configurations C3 / C5 (sweep=False) and C4 (sweep=True) of SURVEY.md §8d.

    Per trajectory: rIi += N(0, 0.05^2), vIi += N(0, 0.02^2), straight-line nodes; per node: r, v +=
    N(0, 0.01^2); attitude = initial-guess quaternion o random rotation of angle U(0, 25 deg) (both sides
    of the |dp| >= 0.95 branch); w ~ N(0, 0.05^2); u = m g [1,0,0] + N(0, 0.003^2) clipped to
    |u| in [Tmin, Tmax]; sigma ~ U(sigma_range) (C3-C5: U(1, 15); the LITERAL rk4 of dynamics.jl:126-128 is
    ill-conditioned for large sigma, so LITERAL parity cases use a range around the reference's sigma = 1).  sweep: per-trajectory mwet*U(.9,1.1), alpha*U(.8,1.2),
    Tmin,Tmax*U(.8,1.2).  `shard` selects an independent, reproducible stream (rank of a sharded run).
    -> X, U, sigma, params (structured PROBINFO array with 1 or B records)."""
    rng = np.random.default_rng([seed, shard])
    info = ProbInfo(prob)
    rIi = prob.rIi[None, :] + rng.normal(0.0, 0.05, (B, 3))
    vIi = prob.vIi[None, :] + rng.normal(0.0, 0.02, (B, 3))
    mwet = None
    params = probinfo_array(info, 1)
    tmin = np.full(B, prob.Tmin)
    tmax = np.full(B, prob.Tmax)
    if sweep:
        mwet = prob.mwet * rng.uniform(0.9, 1.1, B)
        params = probinfo_array(info, B)
        params["a"] = info.a * rng.uniform(0.8, 1.2, B)
        sc = rng.uniform(0.8, 1.2, B)
        tmin, tmax = tmin * sc, tmax * sc
        params["Tmin"] = tmin
    X, U = linear_points_batch(prob, K, rIi, vIi, mwet)
    n = K + 1
    X[..., 1:7] += rng.normal(0.0, 0.01, (B, n, 6))
    ang = np.deg2rad(rng.uniform(0.0, 25.0, (B, n)))
    ax = rng.normal(size=(B, n, 3))
    ax /= np.linalg.norm(ax, axis=-1, keepdims=True)
    dq = np.concatenate([np.cos(ang / 2)[..., None], np.sin(ang / 2)[..., None] * ax], axis=-1)
    X[..., 7:11] = _quat_mul(X[..., 7:11], dq)
    X[..., 11:14] = rng.normal(0.0, 0.05, (B, n, 3))
    U += rng.normal(0.0, 0.003, (B, n, 3))
    nu = np.linalg.norm(U, axis=-1, keepdims=True)
    U *= np.clip(nu, tmin[:, None, None], tmax[:, None, None]) / nu
    sigma = rng.uniform(sigma_range[0], sigma_range[1], B)
    return np.ascontiguousarray(X), np.ascontiguousarray(U), sigma, params


def sample_trajectory(prob: DescentProblem):
    """Configuration C2: the initial-guess nodes of the sample problem (K+1 nodes), sigma = tf_guess,
    dt = 1/(K+1) (rocketland.jl:318, initial_solve.jl:134)."""
    from .first_round import linear_points
    pts = linear_points(prob)
    X = np.stack([p.state for p in pts])[None]
    U = np.stack([p.control for p in pts])[None]
    return X, U, np.array([prob.tf_guess]), 1.0 / (prob.K + 1)
