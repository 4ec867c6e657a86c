"""ctypes front-end of the CPU oracle (oracle/scvx_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module; the product package never does (tests/test_abi.py greps for it).
PARITY UNPINNED by the reference's own tests — see the header of scvx_oracle.cpp.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build as _build

_dp = ctypes.POINTER(ctypes.c_double)
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = _build.OUT
        if not os.path.exists(path):
            _build.build()
        L = ctypes.CDLL(path)
        L.oracle_spline_eval.restype = ctypes.c_double
        L.oracle_linearize_batch.restype = ctypes.c_int
        L.oracle_predict_batch.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleTables:
    """Prefiltered drag/lift (and, for the fins variant, torque) coefficient tables ((n1+2) x (n2+2), column-major)."""

    def __init__(self, drag_samples, lift_samples, cos0, dcos, mach0, dmach, trq_samples=None):
        drag = np.asfortranarray(drag_samples, dtype=np.float64)
        lift = np.asfortranarray(lift_samples, dtype=np.float64)
        n1, n2 = drag.shape
        self.geom = np.array([n1, n2, cos0, dcos, mach0, dmach], dtype=np.float64)
        self.drag = prefilter(drag)
        self.lift = prefilter(lift)
        self.trq = prefilter(np.asfortranarray(trq_samples, dtype=np.float64)) if trq_samples is not None else None

    @classmethod
    def from_aero(cls, aero):
        d, l, t = aero.drag_itrp, aero.lift_itrp, getattr(aero, "trq_itrp", None)
        return cls(d.samples, l.samples, d.cos0, d.dcos, d.mach0, d.dmach, None if t is None else t.samples)


def prefilter(samples):
    s = np.asfortranarray(samples, dtype=np.float64)
    n1, n2 = s.shape
    coef = np.zeros((n1 + 2, n2 + 2), dtype=np.float64, order="F")
    lib().oracle_prefilter(_p(s), n1, n2, _p(coef))
    return coef


def spline_eval(coef, geom, x, y):
    g = np.zeros(2)
    v = lib().oracle_spline_eval(_p(coef), _p(_c(geom)), ctypes.c_double(x), ctypes.c_double(y), _p(g))
    return v, g


def _tb(tables):
    if tables is None:
        return None, None, None
    return _p(tables.drag), _p(tables.lift), _p(tables.geom)


def _params(infos):
    """infos: ProbInfo-like object(s) with .to_c() -> ctypes array + count."""
    from successiveconvexification_b200.defns import CProbInfo
    assert lib().oracle_sizeof_probinfo() == ctypes.sizeof(CProbInfo)
    if isinstance(infos, np.ndarray):       # structured PROBINFO_DTYPE array (workloads.probinfo_array)
        arr = np.ascontiguousarray(infos)
        assert arr.dtype.itemsize == ctypes.sizeof(CProbInfo)
        keep = (CProbInfo * arr.shape[0]).from_buffer_copy(arr.tobytes())
        return keep, arr.shape[0]
    if not isinstance(infos, (list, tuple)):
        infos = [infos]
    arr = (CProbInfo * len(infos))(*[i.to_c() for i in infos])
    assert lib().oracle_sizeof_probinfo() == ctypes.sizeof(CProbInfo)
    return arr, len(infos)


def rhs(info, tables, x, u, sigma):
    arr, _ = _params(info)
    out = np.zeros(14)
    d, l, g = _tb(tables)
    lib().oracle_rhs(arr, d, l, g, _p(_c(x)), _p(_c(u)), ctypes.c_double(sigma), _p(out))
    return out


def aero_force(info, tables, bv, vel):
    arr, _ = _params(info)
    out = np.zeros(3)
    d, l, g = _tb(tables)
    lib().oracle_aero_force(arr, d, l, g, _p(_c(bv)), _p(_c(vel)), _p(out))
    return out


def rk4(info, tables, inp, dt, npts=10, mode=0):
    arr, _ = _params(info)
    out = np.zeros(14)
    d, l, g = _tb(tables)
    lib().oracle_rk4(arr, d, l, g, _p(_c(inp)), ctypes.c_double(dt), npts, mode, _p(out))
    return out


def linearize_interval(info, tables, inp, dt, npts=10, mode=0):
    """-> 14x23 block (Fortran order): col 0 endpoint, cols 1..21 D, col 22 z."""
    arr, _ = _params(info)
    blk = np.zeros((14, 23), order="F")
    d, l, g = _tb(tables)
    lib().oracle_linearize_interval(arr, d, l, g, _p(_c(inp)), ctypes.c_double(dt), npts, mode, _p(blk))
    return blk


def linearize_batch(infos, tables, X, U, sigma, dt, npts=10, mode=0, want_lin_err=True, want_tlb=True,
                    nthreads=0):
    """X (B, n_nodes, 14), U (B, n_nodes, 3), sigma (B,) C-contiguous (= Julia 14 x n_nodes x B).
    -> blocks (B, n_nodes-1, 23, 14), lin_err (B, n_nodes-1, 14), tlb (B, n_nodes, 4), threads_used."""
    X, U, sigma = _c(X), _c(U), _c(sigma)
    B, n_nodes, _ = X.shape
    arr, n = _params(infos)
    assert n in (1, B)
    blocks = np.zeros((B, n_nodes - 1, 23, 14))
    lin_err = np.zeros((B, n_nodes - 1, 14)) if want_lin_err else None
    tlb = np.zeros((B, n_nodes, 4)) if want_tlb else None
    d, l, g = _tb(tables)
    used = lib().oracle_linearize_batch(arr, n, d, l, g, _p(X), _p(U), _p(sigma), ctypes.c_double(dt), npts, mode,
                                        n_nodes, B, _p(blocks), _p(lin_err), _p(tlb), nthreads)
    return blocks, lin_err, tlb, used


def linearize_batch_ex(infos, tables, X, U, sigma, dt, npts=10, mode=0, nthreads=0, precision=0):
    """As `linearize_batch`, in IEEE double (precision=0) or IEEE binary128 rounded to double on output
    (precision=1), plus the branch signature of every interval.
    -> blocks (B, n_nodes-1, 23, 14), sig (B, n_nodes-1) uint64."""
    X, U, sigma = _c(X), _c(U), _c(sigma)
    B, n_nodes, _ = X.shape
    arr, n = _params(infos)
    assert n in (1, B)
    blocks = np.zeros((B, n_nodes - 1, 23, 14))
    sig = np.zeros((B, n_nodes - 1), dtype=np.uint64)
    d, l, g = _tb(tables)
    L = lib()
    L.oracle_linearize_batch_ex.restype = ctypes.c_int
    L.oracle_linearize_batch_ex(arr, n, d, l, g, _p(X), _p(U), _p(sigma), ctypes.c_double(dt), npts, mode, n_nodes, B,
                                _p(blocks), None, None, nthreads, precision,
                                sig.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)))
    return blocks, sig


def linearize_batch_fins(infos, tables, X, U5, sigma, dt, npts=10, mode=0, want_lin_err=True, nthreads=0):
    """SURVEY.md §8f-4 variant (fin forces + aero torque, control_dim = 5): X (B, n_nodes, 14), U5 (B, n_nodes, 5).
    -> blocks (B, n_nodes-1, 27, 14) = [endpoint | D (25 columns) | z], lin_err (B, n_nodes-1, 14)."""
    X, U5, sigma = _c(X), _c(U5), _c(sigma)
    B, n_nodes, _ = X.shape
    assert U5.shape == (B, n_nodes, 5) and tables.trq is not None
    arr, n = _params(infos)
    blocks = np.zeros((B, n_nodes - 1, 27, 14))
    lin_err = np.zeros((B, n_nodes - 1, 14)) if want_lin_err else None
    L = lib()
    L.oracle_linearize_batch_fins.restype = ctypes.c_int
    rc = L.oracle_linearize_batch_fins(arr, n, _p(tables.drag), _p(tables.lift), _p(tables.trq), _p(tables.geom), _p(X), _p(U5),
                                       _p(sigma), ctypes.c_double(dt), npts, mode, n_nodes, B, _p(blocks), _p(lin_err), nthreads)
    assert rc > 0
    return blocks, lin_err


def predict_batch(infos, tables, X, U, sigma, dt, npts=10, mode=0, nthreads=0):
    X, U, sigma = _c(X), _c(U), _c(sigma)
    B, n_nodes, _ = X.shape
    arr, n = _params(infos)
    out = np.zeros((B, n_nodes - 1, 14))
    d, l, g = _tb(tables)
    used = lib().oracle_predict_batch(arr, n, d, l, g, _p(X), _p(U), _p(sigma), ctypes.c_double(dt), npts, mode,
                                      n_nodes, B, _p(out), nthreads)
    return out, used


def max_threads():
    return lib().oracle_max_threads()
