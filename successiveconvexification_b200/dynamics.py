"""Device-backed mirror of module `Dynamics` (reference dynamics.jl) — the drop-in boundary.

Same names, argument meaning and error behaviour as the reference entry points:
    linearize_dynamics(states, tf_guess, base_dt, cache)   dynamics.jl:321-334
    predict_state(x, uk, up, sigma, dt, pinfo, cache)       dynamics.jl:315-317
    simulate_zygote / sensitivity_zygote(inp, dt, cache)    dynamics.jl:308-313
    simulate / sensitivity(inp, dt, cache)                  dynamics.jl:288-305 (deviation form)
    make_state(a, b, sig)                                   dynamics.jl:318-320
    IntegratorCache(prob, info, lin_mod)                    dynamics.jl:258-260
plus additive batched entry points (`linearize_batch`, `predict_batch`) for Monte-Carlo batches.
All numerics run in the CUDA library behind include/scvx_b200.h; nothing here computes dynamics.

Variant note: the device implements the deterministic rk4 + exact forward Jacobian
(`sensitivity_zygote`, "V2").  `simulate`/`sensitivity` return the same quantities in the live code's
deviation form so that `linearize_dynamics`' post-processing (dynamics.jl:327-330) is unchanged.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .defns import (AERO_TABLE, ACC_WIDTH, AtmosphericData, CProbInfo, DescentProblem, INP_DIM, IntegratorCache,
                    LinPoint, LinRes, ProbInfo, STATE_DIM)

MODE_LITERAL = 0      # reproduces dynamics.jl:126-128 (stage increments not scaled by the sub-step)
MODE_TEXTBOOK = 1     # classical RK4
KERNEL_AUTO, KERNEL_DUALWARP, KERNEL_STAGED = 0, 1, 2
TABLE_DRAG, TABLE_LIFT, TABLE_TORQUE = 0, 1, 2

state_idx = slice(0, 14)          # dynamics.jl:136  stateC = 1:14
uk_idx = slice(14, 17)            # dynamics.jl:137  ukC
up_idx = slice(17, 20)            # dynamics.jl:138  upC
sigma_idx = 20                    # dynamics.jl:139  sigmaC


class DeviceContext:
    """Owns one `scvx_ctx` (opaque C handle).  Stored in `IntegratorCache.sim_prob`."""

    def __init__(self, device_ids: Optional[Sequence[int]] = None):
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        if device_ids is None:
            rc = self._lib.scvx_create(ctypes.byref(self._h), None, 0)
        else:
            ids = (ctypes.c_int * len(device_ids))(*device_ids)
            rc = self._lib.scvx_create(ctypes.byref(self._h), ids, len(device_ids))
        _lib.check(rc)
        self.n_params = 0
        # `mode`: stage rule of the rk4-based entry points simulate_zygote / sensitivity_zygote (and the default of the
        # batched calls): LITERAL = the reference's arithmetic (dynamics.jl:126-128), the parity contract.
        # `live_mode`: rule used where the device stands in for the reference's LIVE entry points linearize_dynamics /
        # predict_state / simulate / sensitivity, which integrate the continuous dynamics with an adaptive BS3 solve
        # (dynamics.jl:288-305); the consistent fixed-step integrator for those is classical RK4, so TEXTBOOK.
        self.mode = MODE_LITERAL
        self.live_mode = MODE_TEXTBOOK
        self.npts = 10

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.scvx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setup
    def set_params(self, infos):
        if isinstance(infos, ProbInfo):
            infos = [infos]
        arr = (CProbInfo * len(infos))(*[i.to_c() if isinstance(i, ProbInfo) else i for i in infos])
        _lib.check(self._lib.scvx_set_params(self._h, arr, len(infos)))
        self.n_params = len(infos)

    def set_params_raw(self, arr, n):
        _lib.check(self._lib.scvx_set_params(self._h, arr, n))
        self.n_params = n

    def set_aero(self, aero: AtmosphericData):
        for which, t in ((TABLE_DRAG, aero.drag_itrp), (TABLE_LIFT, aero.lift_itrp), (TABLE_TORQUE, aero.trq_itrp)):
            s = np.asfortranarray(t.samples, dtype=np.float64)
            _lib.check(self._lib.scvx_set_aero_table(self._h, which, s.ctypes.data, s.shape[0], s.shape[1],
                                                     t.cos0, t.dcos, t.mach0, t.dmach, 0))

    def aero_coefficients(self, which: int, n_cos: int, n_mach: int) -> np.ndarray:
        out = np.zeros((n_cos + 2, n_mach + 2), order="F")
        _lib.check(self._lib.scvx_get_aero_coefficients(self._h, which, out.ctypes.data))
        return out

    def set_kernel(self, which: int):
        _lib.check(self._lib.scvx_set_kernel(self._h, which))

    def set_stream(self, cuda_stream: int):
        """Stream of device-pointer calls.  `cuda_stream` is a stream handle as torch reports it
        (`torch.cuda.current_stream().cuda_stream`): torch's default stream has handle 0, which the C ABI reserves for
        "back to the library's own stream", so 0 is passed on as cudaStreamLegacy (handle 1) — the same stream."""
        CUDA_STREAM_LEGACY = 1
        _lib.check(self._lib.scvx_set_stream(self._h, ctypes.c_void_p(cuda_stream or CUDA_STREAM_LEGACY)))

    def use_library_stream(self):
        """Device-pointer calls go back to the library's own (non-blocking) stream: `scvx_set_stream(ctx, NULL)`."""
        _lib.check(self._lib.scvx_set_stream(self._h, None))

    def synchronize(self):
        _lib.check(self._lib.scvx_synchronize(self._h))

    def launch_count(self) -> int:
        return int(self._lib.scvx_launch_count(self._h))

    def last_kernel_ms(self) -> float:
        ms = ctypes.c_double()
        _lib.check(self._lib.scvx_last_kernel_ms(self._h, ctypes.byref(ms)))
        return ms.value

    def linearize_compact_ptr(self, X, U, sigma, base_dt, npts, mode, n_nodes, B, out_compact, out_tlb=0, layout=0):
        _lib.check(self._lib.scvx_linearize_batch_compact(self._h, X, U, sigma, base_dt, npts, mode, n_nodes, B,
                                                          int(layout), out_compact, out_tlb or None))

    # ---- raw pointer calls (host or device addresses)
    def linearize_ptr(self, X, U, sigma, base_dt, npts, mode, n_nodes, B, out_blocks, out_lin_err=0, out_tlb=0):
        _lib.check(self._lib.scvx_linearize_batch(self._h, X, U, sigma, base_dt, npts, mode, n_nodes, B,
                                                  out_blocks, out_lin_err or None, out_tlb or None))

    def defect_cost_ptr(self, X, lin_err, n_nodes, B, wNu, out_defect, out_cost=0):
        _lib.check(self._lib.scvx_defect_cost_batch(self._h, X, lin_err, n_nodes, B, float(wNu), out_defect,
                                                    out_cost or None))

    def linear_points_ptr(self, rIi, vIi, mwet, mwet_shared, mdry, rIf, vIf, g, K, B, X, U):
        rf = (ctypes.c_double * 3)(*[float(v) for v in rIf])
        vf = (ctypes.c_double * 3)(*[float(v) for v in vIf])
        _lib.check(self._lib.scvx_linear_points_batch(self._h, rIi, vIi, mwet or None, float(mwet_shared), float(mdry), rf, vf,
                                                      float(g), int(K), int(B), X, U))

    def dispersed_setup_ptr(self, base, rIi, vIi, mwet, B, X, U, sigma, scales=0, out_params=0, install=True):
        """`base`: a CDimProblem; the other arguments raw host or device addresses."""
        _lib.check(self._lib.scvx_dispersed_setup_batch(self._h, ctypes.addressof(base), rIi, vIi, mwet or None, int(B), X, U,
                                                        sigma, scales or None, out_params or None, 1 if install else 0))

    def socp_values_ptr(self, blocks, lin_err, tlb, n_nodes, B, out_vals, out_rhs=0):
        _lib.check(self._lib.scvx_socp_values_batch(self._h, blocks, lin_err or None, tlb, n_nodes, B, out_vals,
                                                    out_rhs or None))

    def linearize_fins_ptr(self, X, U5, sigma, base_dt, npts, mode, n_nodes, B, out_blocks, out_lin_err=0):
        _lib.check(self._lib.scvx_linearize_batch_fins(self._h, X, U5, sigma, base_dt, npts, mode, n_nodes, B, out_blocks,
                                                       out_lin_err or None))

    def set_fin_tables(self, lift, drag, mach0, dmach, defl0, ddefl):
        """Stage the fin-force tables (`aerodynamics.load_fin_table`): (n_mach, n_defl) column-major lift and drag."""
        for which, t in ((0, lift), (1, drag)):
            a = np.asfortranarray(t, dtype=np.float64)
            _lib.check(self._lib.scvx_set_fin_table(self._h, which, a.ctypes.data, a.shape[0], a.shape[1], float(mach0),
                                                    float(dmach), float(defl0), float(ddefl), 0))

    def fin_force(self, mach, deflection):
        """Fin lift / drag increments at (mach, deflection) pairs from the staged tables (cubic B-spline, Flat ends)."""
        m = np.ascontiguousarray(mach, dtype=np.float64).reshape(-1)
        d = np.ascontiguousarray(deflection, dtype=np.float64).reshape(-1)
        if m.shape != d.shape:
            raise ValueError("mach and deflection must have the same length")
        lift, drag = np.empty_like(m), np.empty_like(m)
        _lib.check(self._lib.scvx_fin_force_batch(self._h, m.ctypes.data, d.ctypes.data, m.size, lift.ctypes.data,
                                                  drag.ctypes.data))
        return lift, drag

    def predict_ptr(self, X, U, sigma, base_dt, npts, mode, n_nodes, B, out):
        _lib.check(self._lib.scvx_predict_batch(self._h, X, U, sigma, base_dt, npts, mode, n_nodes, B, out))


def make_dynamics_module(info: ProbInfo):
    """dynamics.jl:141-215 generates the V1 `Linearizer` module by symbolic codegen.  The device path
    needs no generated code; a descriptor is returned so the documented bring-up sequence
    (rocketland.jl:26-31) keeps its shape."""
    return {"name": "Linearizer", "info": info, "backend": "scvx_b200"}


def make_cache(prob: DescentProblem, info: Optional[ProbInfo] = None, lin_mod=None,
               device_ids: Optional[Sequence[int]] = None) -> IntegratorCache:
    """`IntegratorCache(prob, info, lin_mod)` (dynamics.jl:258-260): builds the device context,
    uploads the parameters and (for AtmosphericData) the spline tables."""
    info = info if info is not None else ProbInfo(prob)
    ctx = DeviceContext(device_ids)
    ctx.set_params(info)
    if isinstance(info.aero, AtmosphericData):
        ctx.set_aero(info.aero)
    return IntegratorCache(sim_prob=ctx, sense_prob=None, sim_int=None, sense_int=None,
                           params=[1.0, info], info=info)


def _ctx(cache: IntegratorCache) -> DeviceContext:
    ctx = cache.sim_prob
    if not isinstance(ctx, DeviceContext):
        raise TypeError("cache does not hold a device context; build it with make_cache(prob, info)")
    return ctx


def make_state(a: LinPoint, b: LinPoint, sig: float) -> np.ndarray:
    """dynamics.jl:318-320."""
    return np.concatenate([a.state, a.control, b.control, [sig]])


# -------------------------------------------------------------------------------------------
# batched entry points (numpy, host memory).  Shapes are C-order views of the Julia arrays:
#   X (B, n_nodes, 14) == Julia (14, n_nodes, B);  blocks (B, n_int, 23, 14) == Julia (14, 23, n_int, B)
# -------------------------------------------------------------------------------------------
def linearize_batch(cache: IntegratorCache, X, U, sigma, base_dt: float, npts: int = 10, mode: int = MODE_LITERAL,
                    lin_err: bool = True, tlb: bool = True):
    ctx = _ctx(cache)
    X = np.ascontiguousarray(X, dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64).reshape(-1)
    if X.ndim != 3 or X.shape[2] != 14 or U.shape != (X.shape[0], X.shape[1], 3) or sigma.shape[0] != X.shape[0]:
        raise ValueError("expected X (B, n_nodes, 14), U (B, n_nodes, 3), sigma (B,)")
    B, n_nodes, _ = X.shape
    ni = max(n_nodes - 1, 0)
    blocks = np.empty((B, ni, ACC_WIDTH, STATE_DIM))
    err = np.empty((B, ni, STATE_DIM)) if lin_err else None
    tl = np.empty((B, n_nodes, 4)) if tlb else None
    ctx.linearize_ptr(X.ctypes.data, U.ctypes.data, sigma.ctypes.data, float(base_dt), int(npts), int(mode),
                      n_nodes, B, blocks.ctypes.data, err.ctypes.data if lin_err else 0, tl.ctypes.data if tlb else 0)
    return blocks, err, tl


COMPACT_DOUBLES, COMPACT_DATA = 230, 229      # include/scvx_b200.h
COMPACT_FULL, COMPACT_NO_Z = 0, 1


def compact_record_doubles(layout: int = COMPACT_FULL) -> int:
    n = _lib.load().scvx_compact_record_doubles(int(layout))
    if n < 0:
        _lib.check(n)
    return n


def compact_layout() -> np.ndarray:
    """Dense offsets (column * 14 + row) of compact slots 0..228 (`scvx_compact_layout`)."""
    idx = np.empty(COMPACT_DATA, np.int32)
    _lib.check(_lib.load().scvx_compact_layout(idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
    return idx


def linearize_batch_compact(cache: IntegratorCache, X, U, sigma, base_dt: float, npts: int = 10, mode: int = MODE_LITERAL,
                            tlb: bool = True, layout: int = COMPACT_FULL):
    """`linearize_batch` with compact result records: only the 229 data entries of every 14x23 block (+ a per-interval
    non-finite flag) cross PCIe; layout COMPACT_NO_Z also drops the z column (215 + 1).
    -> compact (B, n_int, 230 | 216), tlb (B, n_nodes, 4) | None.  `expand_compact` restores the dense blocks and lin_err
    on the host."""
    ctx = _ctx(cache)
    X = np.ascontiguousarray(X, dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64).reshape(-1)
    if X.ndim != 3 or X.shape[2] != 14 or U.shape != (X.shape[0], X.shape[1], 3) or sigma.shape[0] != X.shape[0]:
        raise ValueError("expected X (B, n_nodes, 14), U (B, n_nodes, 3), sigma (B,)")
    B, n_nodes, _ = X.shape
    ni = max(n_nodes - 1, 0)
    comp = np.empty((B, ni, compact_record_doubles(layout)))
    tl = np.empty((B, n_nodes, 4)) if tlb else None
    ctx.linearize_compact_ptr(X.ctypes.data, U.ctypes.data, sigma.ctypes.data, float(base_dt), int(npts), int(mode),
                              n_nodes, B, comp.ctypes.data, tl.ctypes.data if tlb else 0, layout)
    return comp, tl


def expand_compact(compact, X, U=None, sigma=None, lin_err: bool = True, n_threads: int = 0):
    """Host-side expander (`scvx_expand_compact`): compact (B, n_int, 230 | 216) + X (B, n_nodes, 14) [+ U, sigma for the
    NO_Z layout, whose z column is re-formed on the host] ->
    blocks (B, n_int, 23, 14), lin_err (B, n_int, 14) | None, number of intervals flagged non-finite."""
    compact = np.ascontiguousarray(compact, dtype=np.float64)
    X = np.ascontiguousarray(X, dtype=np.float64)
    B, ni, rec = compact.shape
    layout = {COMPACT_DOUBLES: COMPACT_FULL, 216: COMPACT_NO_Z}[rec]
    U = None if U is None else np.ascontiguousarray(U, dtype=np.float64)
    sigma = None if sigma is None else np.ascontiguousarray(sigma, dtype=np.float64).reshape(-1)
    blocks = np.empty((B, ni, ACC_WIDTH, STATE_DIM))
    err = np.empty((B, ni, STATE_DIM)) if lin_err else None
    n = _lib.load().scvx_expand_compact(compact.ctypes.data, layout, X.ctypes.data,
                                        U.ctypes.data if U is not None else None,
                                        sigma.ctypes.data if sigma is not None else None, ni + 1, B, blocks.ctypes.data,
                                        err.ctypes.data if lin_err else None, int(n_threads))
    if n < 0:
        _lib.check(int(n))
    return blocks, err, int(n)


FINS_CONTROL_DIM, FINS_INP_DIM, FINS_ACC_WIDTH = 5, 25, 27      # control_dim 3 -> 5: inp 14 + 2*5 + 1, acc_width 14 + 2*5 + 3


def linearize_batch_fins(cache: IntegratorCache, X, U5, sigma, base_dt: float, npts: int = 10, mode: int = MODE_LITERAL,
                         lin_err: bool = True):
    """SURVEY.md §8f-4 variant: the fin-force (`ff = u[4]*fd1 + u[5]*fd2`) and aero-torque terms the reference carries as
    comments (dynamics.jl:60-63, 66, 69; torque of aerodynamics.jl:45, 49-56), control_dim = 5.  NO REFERENCE CONSUMER.
    X (B, n_nodes, 14), U5 (B, n_nodes, 5) -> blocks (B, n_int, 27, 14) = [endpoint | D (25 columns) | z], lin_err."""
    ctx = _ctx(cache)
    X = np.ascontiguousarray(X, dtype=np.float64)
    U5 = np.ascontiguousarray(U5, dtype=np.float64)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64).reshape(-1)
    if X.ndim != 3 or X.shape[2] != 14 or U5.shape != (X.shape[0], X.shape[1], 5) or sigma.shape[0] != X.shape[0]:
        raise ValueError("expected X (B, n_nodes, 14), U5 (B, n_nodes, 5), sigma (B,)")
    B, n_nodes, _ = X.shape
    ni = max(n_nodes - 1, 0)
    blocks = np.empty((B, ni, FINS_ACC_WIDTH, STATE_DIM))
    err = np.empty((B, ni, STATE_DIM)) if lin_err else None
    ctx.linearize_fins_ptr(X.ctypes.data, U5.ctypes.data, sigma.ctypes.data, float(base_dt), int(npts), int(mode), n_nodes,
                           B, blocks.ctypes.data, err.ctypes.data if lin_err else 0)
    return blocks, err


def predict_batch(cache: IntegratorCache, X, U, sigma, base_dt: float, npts: int = 10, mode: int = MODE_LITERAL):
    ctx = _ctx(cache)
    X = np.ascontiguousarray(X, dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64).reshape(-1)
    B, n_nodes, _ = X.shape
    out = np.empty((B, n_nodes - 1, STATE_DIM))
    ctx.predict_ptr(X.ctypes.data, U.ctypes.data, sigma.ctypes.data, float(base_dt), int(npts), int(mode),
                    n_nodes, B, out.ctypes.data)
    return out


def defect_cost(cache: IntegratorCache, X, lin_err, wNu: float):
    """Fused cost / defect evaluation of the SCvx ratio test (rocketland.jl:289-290): per trajectory
    `defect = norm(x_{k+1} - predict_state(x_k, ...) for k)` and `J = -x[1, K+1] + wNu * defect`, from the `lin_err`
    array of `linearize_batch` at the same inputs.  -> (defect (B,), cost (B,))."""
    ctx = _ctx(cache)
    X = np.ascontiguousarray(X, dtype=np.float64)
    lin_err = np.ascontiguousarray(lin_err, dtype=np.float64)
    B, n_nodes, _ = X.shape
    if lin_err.shape != (B, n_nodes - 1, 14):
        raise ValueError("expected lin_err (B, n_nodes-1, 14)")
    defect, cost = np.empty(B), np.empty(B)
    ctx.defect_cost_ptr(X.ctypes.data, lin_err.ctypes.data, n_nodes, B, wNu, defect.ctypes.data, cost.ctypes.data)
    return defect, cost


def linear_points_batch(cache: IntegratorCache, problem: DescentProblem, rIi, vIi, mwet=None):
    """Batched `linear_points` (initial_solve.jl:113-129) on the device for B dispersed initial conditions:
    rIi, vIi (B, 3), optional per-trajectory mwet (B,).  -> X (B, K+1, 14), U (B, K+1, 3)."""
    ctx = _ctx(cache)
    rIi = np.ascontiguousarray(rIi, dtype=np.float64)
    vIi = np.ascontiguousarray(vIi, dtype=np.float64)
    B, K = rIi.shape[0], problem.K
    if rIi.shape != (B, 3) or vIi.shape != (B, 3):
        raise ValueError("expected rIi, vIi of shape (B, 3)")
    mw = None if mwet is None else np.ascontiguousarray(mwet, dtype=np.float64)
    X, U = np.empty((B, K + 1, 14)), np.empty((B, K + 1, 3))
    ctx.linear_points_ptr(rIi.ctypes.data, vIi.ctypes.data, mw.ctypes.data if mw is not None else 0, problem.mwet,
                          problem.mdry, problem.rIf, problem.vIf, problem.g, K, B, X.ctypes.data, U.ctypes.data)
    return X, U


def blocks_to_linres(blocks_one_traj: np.ndarray):
    """(n_int, 23, 14) C-order -> [LinRes]; derivative is the 14x21 column-major matrix (a Fortran view)."""
    return [LinRes(endpoint=blk[0].copy(), derivative=np.asfortranarray(blk[1:22].T)) for blk in blocks_one_traj]


# -------------------------------------------------------------------------------------------
# the reference entry points
# -------------------------------------------------------------------------------------------
def linearize_dynamics(states: Sequence[LinPoint], tf_guess: float, base_dt: float, cache: IntegratorCache):
    """dynamics.jl:321-334: one LinRes per adjacent pair of nodes."""
    ctx = _ctx(cache)
    X = np.stack([p.state for p in states])[None]
    U = np.stack([p.control for p in states])[None]
    blocks, _, _ = linearize_batch(cache, X, U, [float(tf_guess)], base_dt, ctx.npts, ctx.live_mode, lin_err=False,
                                   tlb=False)
    return blocks_to_linres(blocks[0])


def _one_interval(inp):
    inp = np.asarray(inp, dtype=np.float64)
    if inp.shape != (INP_DIM,):
        raise ValueError("inp must have 21 entries [x; uk; up; sigma]")
    X = np.zeros((1, 2, 14)); U = np.zeros((1, 2, 3))
    X[0, 0] = inp[state_idx]; U[0, 0] = inp[uk_idx]; U[0, 1] = inp[up_idx]
    return inp, X, U


def simulate_zygote(inp, dt: float, cache: IntegratorCache, npts: int = 10, mode: Optional[int] = None) -> np.ndarray:
    """dynamics.jl:308-310: rk4(inp, dt, info) -> absolute end state (14).  LITERAL stage rule unless told otherwise."""
    inp, X, U = _one_interval(inp)
    return predict_batch(cache, X, U, [inp[sigma_idx]], dt, npts, _ctx(cache).mode if mode is None else mode)[0, 0]


def sensitivity_zygote(inp, dt: float, cache: IntegratorCache, mode: Optional[int] = None):
    """dynamics.jl:311-313: (y(14), J^T (21x14)) as Zygote.forward_jacobian returns them."""
    inp, X, U = _one_interval(inp)
    ctx = _ctx(cache)
    blocks, _, _ = linearize_batch(cache, X, U, [inp[sigma_idx]], dt, ctx.npts, ctx.mode if mode is None else mode,
                                   lin_err=False, tlb=False)
    blk = blocks[0, 0]
    return blk[0].copy(), blk[1:22].copy()          # rows of blk[1:22] are the columns of D, i.e. J^T


def simulate(inp, dt: float, cache: IntegratorCache) -> np.ndarray:
    """dynamics.jl:288-296: absolute end state (the reference returns `(u .+ inp)[1:14]`).  The reference integrates
    the continuous dynamics (adaptive BS3); the device uses the consistent fixed-step rule (`live_mode`, TEXTBOOK)."""
    ctx = _ctx(cache)
    return simulate_zygote(inp, dt, cache, ctx.npts, ctx.live_mode)


def sensitivity(inp, dt: float, cache: IntegratorCache):
    """dynamics.jl:298-305 in the live code's DEVIATION form: val(21) = x(dt) - x(0) (zero-padded),
    mat(21x21) with mat[0:14, :] = D - [I 0]; `linearize_dynamics` adds x and I back (327-330).  `live_mode` rule."""
    inp = np.asarray(inp, dtype=np.float64)
    y, JT = sensitivity_zygote(inp, dt, cache, _ctx(cache).live_mode)
    val = np.zeros(INP_DIM); val[:14] = y - inp[:14]
    mat = np.zeros((INP_DIM, INP_DIM), order="F"); mat[:14, :] = JT.T
    mat[np.arange(14), np.arange(14)] -= 1.0
    return val, mat


def predict_state(initial_state, uk, up, sigma, dt, pinfo, cache: IntegratorCache) -> np.ndarray:
    """dynamics.jl:315-317."""
    return simulate(np.concatenate([initial_state, uk, up, [sigma]]), dt, cache)
