"""Boundary records of the hot path — mirror of module `RocketlandDefns` (reference master.jl:1-136).

Only the types the linearise-and-discretise path touches are mirrored: `DescentProblem`
(master.jl:17-71), `ProbInfo` (73-83), `LinPoint` (85-88), `LinRes` (90-93), `AtmosphericData` /
`ExoatmosphericData` (6-16) and `IntegratorCache` (113-120).  Field names, defaults and meaning
follow the reference; arrays are numpy float64 and matrices are stored the way Julia stores them
(column-major, `order="F"`), so a `LinRes.derivative` can be handed to host code column by column
exactly as `eachcol(derivative)` does at rocketland.jl:125-126.
"""
from __future__ import annotations

import ctypes
import dataclasses
from dataclasses import dataclass, field
from typing import Any, Optional

import numpy as np

STATE_DIM = 14          # rocketland.jl:16
CONTROL_DIM = 3         # rocketland.jl:17
INP_DIM = STATE_DIM + 2 * CONTROL_DIM + 1      # 21, dynamics.jl:136-139
ACC_WIDTH = STATE_DIM + 2 * CONTROL_DIM + 3    # 23, rocketland.jl:22
ACC_HEIGHT = STATE_DIM                          # rocketland.jl:23

AERO_EXO = 0
AERO_TABLE = 1


def _f64(x, shape=None):
    a = np.array(x, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class AerodynamicInfo:
    """abstract type AerodynamicInfo (master.jl:6)."""


class ExoatmosphericData(AerodynamicInfo):
    """master.jl:8 — no atmosphere; the hot path uses zero aerodynamic force."""

    def __repr__(self):
        return "ExoatmosphericData()"


@dataclass
class AeroTable:
    """One `extrapolate(scale(interpolate(A, BSpline(Cubic(Line(OnGrid())))), aoa, mach), Flat())`
    object (aerodynamics.jl:19-21), kept as the RAW samples plus the two StepRangeLen axes.  The
    cubic-B-spline prefilter runs on the device when the table is uploaded."""
    samples: np.ndarray            # (n_cos, n_mach), Fortran order, as reshape() at aerodynamics.jl:19
    cos0: float
    dcos: float
    mach0: float
    dmach: float

    @property
    def shape(self):
        return self.samples.shape


@dataclass
class AtmosphericData(AerodynamicInfo):
    """master.jl:10-16."""
    drag_itrp: AeroTable
    lift_itrp: AeroTable
    trq_itrp: AeroTable
    force_scalar: float = 1.0
    length_scalar: float = 1.0


@dataclass
class DescentProblem:
    """master.jl:17-71 (keyword constructor defaults at 65-70)."""
    g: float = 1.0
    mdry: float = 1.0
    mwet: float = 2.0
    Tmin: float = 0.3
    Tmax: float = 5.0
    deltaMax: float = 20.0
    thetaMax: float = 90.0
    gammaGs: float = 20.0
    omMax: float = 60.0
    dpMax: float = 50000.0
    jB: np.ndarray = field(default_factory=lambda: np.diag([1e-2, 1e-2, 1e-2]))
    alpha: float = 0.01
    rho: float = 1.225
    sos: float = 5.0
    rTB: np.ndarray = field(default_factory=lambda: _f64([-1e-2, 0, 0]))
    rFB: np.ndarray = field(default_factory=lambda: _f64([1e-2, 0, 0]))
    rIi: np.ndarray = field(default_factory=lambda: _f64([4.0, 4.0, 0.0]))
    rIf: np.ndarray = field(default_factory=lambda: _f64([0.0, 0.0, 0.0]))
    vIi: np.ndarray = field(default_factory=lambda: _f64([0, -2, 2]))
    vIf: np.ndarray = field(default_factory=lambda: _f64([-0.1, 0.0, 0.0]))
    qBIi: np.ndarray = field(default_factory=lambda: _f64([1.0, 0, 0, 0]))
    qBIf: np.ndarray = field(default_factory=lambda: _f64([1.0, 0, 0, 0]))
    wBi: np.ndarray = field(default_factory=lambda: _f64([0.0, 0.0, 0.0]))
    wBf: np.ndarray = field(default_factory=lambda: _f64([0.0, 0, 0]))
    aero: AerodynamicInfo = field(default_factory=ExoatmosphericData)
    K: int = 50
    imax: int = 15
    wNu: float = 1e5
    wID: float = 1e-3
    wDS: float = 1e-1
    wCst: float = 10.0
    wTviol: float = 100.0
    nuTol: float = 1e-10
    delTol: float = 1e-3
    tf_guess: float = 1.0
    ri: float = 1.0
    rh0: float = 0.0
    rh1: float = 0.25
    rh2: float = 0.90
    alph: float = 2.0
    bet: float = 3.2

    def __post_init__(self):
        self.jB = _f64(self.jB, (3, 3))
        for name in ("rTB", "rFB", "rIi", "rIf", "vIi", "vIf", "wBi", "wBf"):
            setattr(self, name, _f64(getattr(self, name), (3,)))
        for name in ("qBIi", "qBIf"):
            setattr(self, name, _f64(getattr(self, name), (4,)))

    def replace(self, **kw) -> "DescentProblem":
        return dataclasses.replace(self, **kw)


class CProbInfo(ctypes.Structure):
    """`scvx_probinfo` of include/scvx_b200.h (3x3 matrices column-major)."""
    _fields_ = [
        ("a", ctypes.c_double), ("g0", ctypes.c_double), ("sos", ctypes.c_double),
        ("jB", ctypes.c_double * 9), ("jBi", ctypes.c_double * 9),
        ("rTB", ctypes.c_double * 3), ("rFB", ctypes.c_double * 3),
        ("force_scalar", ctypes.c_double), ("length_scalar", ctypes.c_double),
        ("Tmin", ctypes.c_double),
        ("aero_kind", ctypes.c_int32), ("_pad", ctypes.c_int32),
    ]


class CDimProblem(ctypes.Structure):
    """`scvx_dim_problem` of include/scvx_b200.h: the dimensional DescentProblem fields the path consumes."""
    _fields_ = [
        ("g", ctypes.c_double), ("mdry", ctypes.c_double), ("mwet", ctypes.c_double), ("Tmin", ctypes.c_double),
        ("Tmax", ctypes.c_double), ("alpha", ctypes.c_double), ("sos", ctypes.c_double), ("tf_guess", ctypes.c_double),
        ("jB", ctypes.c_double * 9), ("rTB", ctypes.c_double * 3), ("rFB", ctypes.c_double * 3), ("rIf", ctypes.c_double * 3),
        ("aero_kind", ctypes.c_int32), ("K", ctypes.c_int32),
    ]

    @classmethod
    def from_problem(cls, dp: "DescentProblem") -> "CDimProblem":
        c = cls()
        for name in ("g", "mdry", "mwet", "Tmin", "Tmax", "alpha", "sos", "tf_guess"):
            setattr(c, name, float(getattr(dp, name)))
        c.jB = (ctypes.c_double * 9)(*np.asarray(dp.jB, dtype=np.float64).reshape(3, 3).T.ravel())   # column-major
        c.rTB = (ctypes.c_double * 3)(*np.asarray(dp.rTB, dtype=np.float64))
        c.rFB = (ctypes.c_double * 3)(*np.asarray(dp.rFB, dtype=np.float64))
        c.rIf = (ctypes.c_double * 3)(*np.asarray(dp.rIf, dtype=np.float64))
        c.aero_kind = AERO_TABLE if isinstance(dp.aero, AtmosphericData) else AERO_EXO
        c.K = int(dp.K)
        return c


@dataclass
class ProbInfo:
    """master.jl:73-83.  `ProbInfo(problem)` copies alpha→a, g→g0, sos, jB, inv(jB), rTB, rFB, aero
    (constructor at master.jl:82).  `Tmin` rides along for the thrust-lower-bound rows
    (rocketland.jl:199-200, 261-263)."""
    a: float
    g0: float
    sos: float
    jB: np.ndarray
    jBi: np.ndarray
    rTB: np.ndarray
    rFB: np.ndarray
    aero: AerodynamicInfo
    Tmin: float = 0.0

    def __init__(self, from_: Optional[DescentProblem] = None, **kw):
        if from_ is not None:
            self.a = float(from_.alpha)
            self.g0 = float(from_.g)
            self.sos = float(from_.sos)
            self.jB = _f64(from_.jB, (3, 3))
            self.jBi = np.linalg.inv(self.jB)
            self.rTB = _f64(from_.rTB, (3,))
            self.rFB = _f64(from_.rFB, (3,))
            self.aero = from_.aero
            self.Tmin = float(from_.Tmin)
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def aero_kind(self) -> int:
        return AERO_TABLE if isinstance(self.aero, AtmosphericData) else AERO_EXO

    def to_c(self) -> CProbInfo:
        c = CProbInfo()
        c.a, c.g0, c.sos = self.a, self.g0, self.sos
        c.jB[:] = list(np.asarray(self.jB, dtype=np.float64).reshape(-1, order="F"))
        c.jBi[:] = list(np.asarray(self.jBi, dtype=np.float64).reshape(-1, order="F"))
        c.rTB[:] = list(self.rTB)
        c.rFB[:] = list(self.rFB)
        if isinstance(self.aero, AtmosphericData):
            c.force_scalar, c.length_scalar = self.aero.force_scalar, self.aero.length_scalar
        else:
            c.force_scalar, c.length_scalar = 0.0, 0.0
        c.Tmin = self.Tmin
        c.aero_kind = self.aero_kind
        return c


@dataclass
class LinPoint:
    """master.jl:85-88."""
    state: np.ndarray      # 14
    control: np.ndarray    # 3

    def __post_init__(self):
        self.state = _f64(self.state, (STATE_DIM,))
        self.control = _f64(self.control, (CONTROL_DIM,))


@dataclass
class LinRes:
    """master.jl:90-93.  `derivative` is 14x21 = [A | B- | B+ | Sigma] (column-major)."""
    endpoint: np.ndarray
    derivative: np.ndarray

    # named views of the first-order-hold matrices (old_dynamics.jl:84-98, 135-148)
    @property
    def A(self):
        return self.derivative[:, 0:14]

    @property
    def Bm(self):
        return self.derivative[:, 14:17]

    @property
    def Bp(self):
        return self.derivative[:, 17:20]

    @property
    def Sigma(self):
        return self.derivative[:, 20]


@dataclass
class IntegratorCache:
    """master.jl:113-120.  Five of the six reference fields are typed `Any`; the device context
    lives in `sim_prob`, the rest stay unused by the device path."""
    sim_prob: Any = None
    sense_prob: Any = None
    sim_int: Any = None
    sense_int: Any = None
    params: Any = None
    info: Optional[ProbInfo] = None
