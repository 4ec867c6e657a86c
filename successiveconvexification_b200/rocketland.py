"""Host-side mirror of the part of the reference's `Rocketland` module that this path feeds: the trajectory-dependent
SOCP rows.

The reference rebuilds the K dynamics equality blocks (`rocketland.jl:117-133`) and the K+1 linearised thrust-lower-bound
rows (`rocketland.jl:194-201`) once in `build_model` and then refreshes them with K*21 + 3(K+1) `MOI.modify` calls per
iteration (`rocketland.jl:251-265`).  Their sparsity pattern is iteration invariant, so the device writes the value array
of a fixed compressed-sparse-column pattern instead: a direct conic-solver interface (ECOS-style `A, b` / `G, h`) is
refreshed with one copy.  Everything else of `build_model` (objective, cones, boundary conditions) is trajectory
independent and stays with the host.

Variables (local column order = the reference's creation order `dxv, duv, dsig, nuv`, `rocketland.jl:73-76`):
`dxv[j,n] -> 14n + j`, `duv[j,n] -> 14(K+1) + 3n + j`, `dsig -> 17(K+1)`, `nuv[j,n] -> 17(K+1) + 1 + 14n + j`.
Rows: `14n + i` (dynamics row i of interval n, Zeros cone), then `14K + n` (thrust lower bound of node n, Nonpositives).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .defns import IntegratorCache

state_dim, control_dim = 14, 3          # rocketland.jl:16-24
acc_width, acc_height = 23, 14          # rocketland.jl:22-23


def socp_dims(n_nodes: int):
    """-> (n_rows, n_cols, nnz) of the trajectory-dependent rows for K = n_nodes - 1 intervals."""
    lib = _lib.load()
    nr, nc, nz = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    _lib.check(lib.scvx_socp_dims(int(n_nodes), ctypes.byref(nr), ctypes.byref(nc), ctypes.byref(nz)))
    return nr.value, nc.value, nz.value


def socp_pattern(n_nodes: int):
    """CSC pattern (0-based) -> (n_rows, n_cols, colptr int32 (n_cols+1), rowind int32 (nnz))."""
    lib = _lib.load()
    nr, nc, nz = socp_dims(n_nodes)
    colptr, rowind = np.empty(nc + 1, np.int32), np.empty(nz, np.int32)
    _lib.check(lib.scvx_socp_pattern(int(n_nodes), colptr.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                     rowind.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
    return nr, nc, colptr, rowind


def variable_columns(K: int):
    """Local column indices of the reference's variable arrays: dict of dxv (14, K+1), duv (3, K+1), dsig, nuv (14, K+1)."""
    n = K + 1
    return {
        "dxv": np.arange(14 * n).reshape(n, 14).T,
        "duv": 14 * n + np.arange(3 * n).reshape(n, 3).T,
        "dsig": 17 * n,
        "nuv": 17 * n + 1 + np.arange(14 * n).reshape(n, 14).T,
    }


def socp_values_batch(cache: IntegratorCache, blocks, lin_err, tlb):
    """Value arrays of the fixed pattern for B trajectories from `linearize_batch`'s outputs at the same inputs:
    blocks (B, K, 23, 14), lin_err (B, K, 14), tlb (B, K+1, 4)  ->  vals (B, nnz), rhs (B, n_rows)."""
    from .dynamics import _ctx
    ctx = _ctx(cache)
    blocks = np.ascontiguousarray(blocks, dtype=np.float64)
    lin_err = np.ascontiguousarray(lin_err, dtype=np.float64)
    tlb = np.ascontiguousarray(tlb, dtype=np.float64)
    B, n_nodes = tlb.shape[0], tlb.shape[1]
    if blocks.shape != (B, n_nodes - 1, acc_width, acc_height) or lin_err.shape != (B, n_nodes - 1, 14) or tlb.shape[2] != 4:
        raise ValueError("expected blocks (B, K, 23, 14), lin_err (B, K, 14), tlb (B, K+1, 4)")
    nr, _, nz = socp_dims(n_nodes)
    vals, rhs = np.empty((B, nz)), np.empty((B, nr))
    ctx.socp_values_ptr(blocks.ctypes.data, lin_err.ctypes.data, tlb.ctypes.data, n_nodes, B, vals.ctypes.data, rhs.ctypes.data)
    return vals, rhs
