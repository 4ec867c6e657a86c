"""Initial guess — mirror of module `FirstRound` (reference initial_solve.jl:113-135).  Host code."""
from __future__ import annotations

import numpy as np

from .defns import DescentProblem, LinPoint


def rotation_between(u, v) -> np.ndarray:
    """Rotations.jl `rotation_between(u, v)` as used at initial_solve.jl:121: the unit quaternion
    (w, x, y, z) of the shortest rotation taking u onto v, `normalize([|u||v| + u.v ; u x v])`."""
    u = np.asarray(u, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    normprod = np.sqrt(np.dot(u, u) * np.dot(v, v))
    if normprod < np.finfo(np.float64).eps:
        raise ValueError("Input vectors must be nonzero.")
    w = normprod + np.dot(u, v)
    if abs(w) < 100 * np.finfo(np.float64).eps:
        # antiparallel: any axis perpendicular to u
        k = int(np.argmin(np.abs(u)))
        e = np.zeros(3)
        e[k] = 1.0
        axis = np.cross(u, e)
    else:
        axis = np.cross(u, v)
    q = np.array([w, axis[0], axis[1], axis[2]])
    return q / np.linalg.norm(q)


def linear_points(problem: DescentProblem):
    """initial_solve.jl:113-129 — K+1 nodes on a straight line in mass, position and velocity,
    attitude = rotation of body +x onto -v, control = hover thrust [m g, 0, 0]."""
    K = problem.K
    pts = []
    for k in range(K + 1):
        mk = (K - k) / K * problem.mwet + (k / K) * problem.mdry
        rIk = (K - k) / K * problem.rIi + (k / K) * problem.rIf
        vIk = (K - k) / K * problem.vIi + (k / K) * problem.vIf
        qBIk = rotation_between([1.0, 0.0, 0.0], -vIk)
        state = np.concatenate([[mk], rIk, vIk, qBIk, [0.0, 0.0, 0.0]])
        pts.append(LinPoint(state, np.array([mk * problem.g, 0.0, 0.0])))
    return pts


def linear_initial(problem: DescentProblem, cache):
    """initial_solve.jl:131-135."""
    from .dynamics import linearize_dynamics
    pts = linear_points(problem)
    return pts, linearize_dynamics(pts, problem.tf_guess, 1 / (problem.K + 1), cache)
