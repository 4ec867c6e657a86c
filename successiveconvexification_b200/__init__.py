"""successiveconvexification_b200 — B200-native linearise-and-discretise path of
BenChung/SuccessiveConvexification behind the reference's own `Dynamics` entry points.

Host-side records and problem library (pure Python, mirrors of master.jl / sample_problems.jl /
initial_solve.jl) import without the CUDA library; everything under `dynamics` needs
libscvx_b200.so and fails loudly without it (no CPU fallback).
"""
from .defns import (AtmosphericData, DescentProblem, ExoatmosphericData, IntegratorCache, LinPoint, LinRes,
                    ProbInfo)
from . import aerodynamics, first_round, sample_problems   # noqa: F401
from . import dynamics                                      # noqa: F401

__all__ = ["AtmosphericData", "DescentProblem", "ExoatmosphericData", "IntegratorCache", "LinPoint", "LinRes",
           "ProbInfo", "aerodynamics", "dynamics", "first_round", "sample_problems"]
__version__ = "0.1.0"
