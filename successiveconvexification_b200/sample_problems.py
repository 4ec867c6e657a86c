"""Problem library — mirror of module `SampleProblems` (reference sample_problems.jl:1-33).

`normalize_problem` reproduces the reference's non-dimensionalisation literally, including its
quirks: `vIf` is overwritten by the scaled `vIi` (sample_problems.jl:15),
`rFB` is scaled by 1/Ut (16) and `nuTol` falls back to its default (not forwarded).
"""
from __future__ import annotations

import numpy as np

from .aerodynamics import load_aerodata, rescale_aerodata
from .defns import DescentProblem


def normalize_problem(dp: DescentProblem) -> DescentProblem:
    """sample_problems.jl:5-23."""
    Ul = float(np.max(dp.rIi))
    Ut = float(dp.tf_guess)
    Um = float(dp.mwet)
    return DescentProblem(
        g=dp.g / (Ul / Ut ** 2), mdry=dp.mdry / Um, mwet=dp.mwet / Um,
        Tmin=dp.Tmin / (Um * Ul / Ut ** 2), Tmax=dp.Tmax / (Um * Ul / Ut ** 2),
        omMax=dp.omMax / Ut, jB=dp.jB * (1 / (Um * Ul ** 2)),
        rTB=dp.rTB * (1 / Ul), rIi=dp.rIi * (1 / Ul),
        rIf=dp.rIf * (1 / Ul), vIi=dp.vIi * (1 / (Ul / Ut)),
        vIf=dp.vIi * (1 / (Ul / Ut)), qBIf=dp.qBIf, qBIi=dp.qBIi,
        wBi=dp.wBi, wBf=dp.wBf, rFB=dp.rFB * (1 / Ut),
        deltaMax=dp.deltaMax, thetaMax=dp.thetaMax, gammaGs=dp.gammaGs,
        alpha=dp.alpha / (Ut ** 2 / Ul), K=dp.K, imax=dp.imax, wNu=dp.wNu, wID=dp.wID,
        wDS=dp.wDS, wCst=dp.wCst, wTviol=dp.wTviol, delTol=dp.delTol,
        tf_guess=dp.tf_guess / Ut, ri=dp.ri, rh0=dp.rh0, rh1=dp.rh1,
        rh2=dp.rh2, alph=dp.alph, bet=dp.bet, dpMax=dp.dpMax / (Um / (Ul * Ut ** 2)),
        rho=dp.rho / (Um / Ul ** 3), sos=dp.sos / (Ul / Ut),
        aero=rescale_aerodata(dp.aero, Ul, Ut, Um))


def _base_kwargs():
    # sample_problems.jl:26-27 / 30-31
    return dict(g=9.82, mwet=66018.0, mdry=65947.0, Tmin=0.1 * 4.686588e6, Tmax=4.686588e6,
                jB=np.diag([72487.03125, 2.0734175e6, 2.0734175e6]),
                alpha=0.000345, rTB=[-4.26114, 0, 0], rFB=[2.0, 0, 0], rIi=[1000.0, 1000.0, 100.0],
                rIf=[0.0, 0.0, 0.0], vIi=[-100.0, -200.0, 0], sos=352.0, wNu=1e4)


def base_prob() -> DescentProblem:
    """sample_problems.jl:26-27 (exo-atmospheric)."""
    return DescentProblem(**_base_kwargs())


def base_prob_scaled() -> DescentProblem:
    """sample_problems.jl:28."""
    return normalize_problem(base_prob())


def base_prob_aero(liftdrag) -> DescentProblem:
    """sample_problems.jl:30-31.  `liftdrag` is the path of the reference's aero/lift_drag.csv (or an
    .npz produced by tests/golden/make_aero_fixture.py), or an already loaded AtmosphericData."""
    aero = liftdrag if not isinstance(liftdrag, (str, bytes)) and not hasattr(liftdrag, "__fspath__") \
        else load_aerodata(liftdrag)
    return DescentProblem(aero=aero, **_base_kwargs())


def base_prob_aero_scaled(liftdrag) -> DescentProblem:
    """sample_problems.jl:32."""
    return normalize_problem(base_prob_aero(liftdrag))


def dispersed_setup(cache, dim_problem: DescentProblem, rIi, vIi, mwet=None, install: bool = True):
    """Monte-Carlo dispersions of a DIMENSIONAL problem, set up on the device in one launch: per
    trajectory `normalize_problem` (sample_problems.jl:5-23, with Ul = max(rIi_b), Um = mwet_b), `ProbInfo` of the
    normalised problem (master.jl:73-83) and its `linear_points` initial guess (initial_solve.jl:113-129).
    rIi, vIi (B, 3) and mwet (B,) are dimensional.  With `install` the records become the cache's per-trajectory
    parameters.  -> X (B, K+1, 14), U (B, K+1, 3), sigma (B,), scales (B, 3) = [Ul, Ut, Um], params (structured, B)."""
    from .defns import CDimProblem
    from .dynamics import _ctx
    from .workloads import PROBINFO_DTYPE
    ctx = _ctx(cache)
    rIi = np.ascontiguousarray(rIi, dtype=np.float64)
    vIi = np.ascontiguousarray(vIi, dtype=np.float64)
    B, K = rIi.shape[0], dim_problem.K
    if rIi.shape != (B, 3) or vIi.shape != (B, 3):
        raise ValueError("expected rIi, vIi of shape (B, 3)")
    mw = None if mwet is None else np.ascontiguousarray(mwet, dtype=np.float64)
    X, U = np.empty((B, K + 1, 14)), np.empty((B, K + 1, 3))
    sigma, scales = np.empty(B), np.empty((B, 3))
    params = np.zeros(B, dtype=PROBINFO_DTYPE)
    base = CDimProblem.from_problem(dim_problem)
    ctx.dispersed_setup_ptr(base, rIi.ctypes.data, vIi.ctypes.data, mw.ctypes.data if mw is not None else 0, B,
                            X.ctypes.data, U.ctypes.data, sigma.ctypes.data, scales.ctypes.data, params.ctypes.data, install)
    return X, U, sigma, scales, params
