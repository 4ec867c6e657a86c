import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
AERO_NPZ = os.path.join(GOLDEN, "aero_lift_drag.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def prob_aero():
    from successiveconvexification_b200 import sample_problems as sp
    return sp.base_prob_aero_scaled(AERO_NPZ)


@pytest.fixture(scope="session")
def prob_exo(prob_aero):
    from successiveconvexification_b200.defns import ExoatmosphericData
    return prob_aero.replace(aero=ExoatmosphericData())


@pytest.fixture(scope="session")
def oracle_tables(prob_aero):
    from oracle import oracle
    return oracle.OracleTables.from_aero(prob_aero.aero)


PARITY_FLOOR = 1e-4
PARTS = {"endpoint": slice(0, 1), "A": slice(1, 15), "Bm": slice(15, 18), "Bp": slice(18, 21),
         "Sigma": slice(21, 22), "z": slice(22, 23)}


def parity_metric_per_interval(got, ref, floor_frac=PARITY_FLOOR):
    """(n_intervals, 6) array of  max |delta| / max(|ref|, floor_frac * max|part|)  per interval and part
    (order of PARTS).  NaN/Inf in either operand give inf."""
    got = np.asarray(got).reshape(-1, 23, 14)
    ref = np.asarray(ref).reshape(-1, 23, 14)
    out = np.zeros((got.shape[0], len(PARTS)))
    for k, (name, sl) in enumerate(PARTS.items()):
        g, r = got[:, sl, :], ref[:, sl, :]
        scale = np.abs(r).max(axis=(1, 2), keepdims=True)
        if name == "z":      # z = endpoint - D*inp is formed by cancellation of terms of the size of D and endpoint
            scale = np.abs(ref[:, 0:22, :]).max(axis=(1, 2), keepdims=True)
        den = np.maximum(np.abs(r), floor_frac * scale)
        den[den == 0.0] = 1.0
        with np.errstate(invalid="ignore", over="ignore"):
            m = (np.abs(g - r) / den).max(axis=(1, 2)) if g.size else np.zeros(0)
        out[:, k] = np.where(np.isfinite(m), m, np.inf)
    return out


def parity_report(got, ref, floor_frac=PARITY_FLOOR):
    """Parity protocol item (i) on (..., 23, 14) blocks.  Per part (endpoint, A, B-, B+, Sigma, z) and per interval:
        max |delta| / max(|ref|, floor_frac * max|part|)
    i.e. 1e-10 RELATIVE per matrix entry for every entry within 4 decades of the part's largest, and an
    ABSOLUTE 1e-14 * max|part| for smaller ones (cancellation zeros).  SURVEY.md §8d proposes a floor of
    1e-12*max|block|; measured here, two independent correct FP64 CPU implementations (dual-number C++ vs
    complex-step numpy) already disagree by 6e-16*max|part| on such entries, i.e. 6e-4 by that metric, so the
    floor is placed where FP64 can resolve it (DESIGN.md "Parity metric").  Returns {part: metric}."""
    m = parity_metric_per_interval(got, ref, floor_frac)
    return {name: (float(m[:, k].max()) if m.shape[0] else 0.0) for k, name in enumerate(PARTS)}


def structural_constants():
    """(mask (23,14) bool, values (23,14)): the entries of a block that are structural constants of the dynamics
    (the complement of the compact layout, csrc/scvx_compact.h): position columns e_r, D[m,m] = 1,
    zero patterns of the mass row and of the q / w rows."""
    mask = np.ones((23, 14), dtype=bool)
    vals = np.zeros((23, 14))
    rows = {0: (0, 14), 1: (1, 7)}
    for c in range(23):
        if c == 0 or c >= 15:
            lo, hi = 0, 14
        elif c == 1:
            lo, hi = 1, 7
        elif c <= 4:
            lo, hi = 0, 0
        elif c <= 7:
            lo, hi = 1, 7
        elif c <= 11:
            lo, hi = 1, 11
        else:
            lo, hi = 1, 14
        mask[c, lo:hi] = False
    vals[1, 0] = 1.0
    vals[2, 1] = vals[3, 2] = vals[4, 3] = 1.0
    return mask, vals


def parity_protocol(got, ref):
    """All three items of the parity protocol on (..., 23, 14) blocks:
      (i)   `metric`: parity_report (the pass/fail quantity, <= 1e-10);
      (ii)  `strict_rel_max`: the strict per-entry relative maximum over entries with |ref| > 0, for information — it is
            dominated by cancellation zeros (e.g. the w_x row with jB22 == jB33: FP64 ~1e-17 where the exact value is 0
            or ~1e-53), which is why it is not the pass/fail quantity;
      (iii) `structural_max`: max |got - constant| / max|block| over the structural constants (must be <= 1e-16;
            the kernels produce them exactly, so this is 0)."""
    g = np.asarray(got).reshape(-1, 23, 14)
    r = np.asarray(ref).reshape(-1, 23, 14)
    rep = {"metric": parity_report(g, r)}
    nz = np.abs(r) > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(nz, np.abs(g - r) / np.abs(r), 0.0)
    rep["strict_rel_max"] = float(rel.max()) if rel.size else 0.0
    if rel.size:
        i, c, row = np.unravel_index(int(rel.argmax()), rel.shape)
        rep["strict_rel_where"] = {"interval": int(i), "block_column": int(c), "row": int(row),
                                   "ref": float(r[i, c, row]), "got": float(g[i, c, row])}
    mask, vals = structural_constants()
    scale = np.abs(r).max(axis=(1, 2), keepdims=True)
    scale[scale == 0.0] = 1.0
    rep["structural_max"] = float((np.abs(g - vals[None]) / scale)[:, mask].max()) if g.size else 0.0
    return rep


PARITY_TOL = 1e-10     # BASELINE.json north_star: 1e-10 relative per matrix entry


def assert_parity(got, ref, tol=PARITY_TOL):
    rep = parity_report(got, ref)
    bad = {k: v for k, v in rep.items() if not (v <= tol)}
    assert not bad, f"parity failed: {rep}"
    return rep


def assert_structural_constants(got):
    """Parity protocol item (iii): the structural constants of every block are exact."""
    g = np.asarray(got).reshape(-1, 23, 14)
    mask, vals = structural_constants()
    assert np.array_equal(g[:, mask], np.broadcast_to(vals[mask], (g.shape[0], int(mask.sum())))), \
        "a structural constant of a block is not exact"


# ---------------------------------------------------------------------------------------------------------------
# Conditioning-aware parity (LITERAL stage rule at large sigma).  The reference's rk4 does not scale its stage
# increments by the sub-step (dynamics.jl:126-128), so for sigma >> 1 the discrete map amplifies rounding errors: no
# FP64 implementation — the reference's own included — holds 1e-10 there.  How far FP64 CAN resolve each interval is
# measured, not assumed:
#   * `refq`: the oracle evaluates the same operation sequence in IEEE binary128 — the exact value, for this purpose;
#   * kappa*eps of an interval = the RESOLUTION OF THE REFERENCE ARITHMETIC there: the largest distance from `refq` of
#     N_PERTURB + 1 FP64 evaluations of the oracle, one at the inputs and N_PERTURB at inputs moved by at most one ulp
#     per entry (each of them is as valid an FP64 answer as the reference's own; a single evaluation under-estimates
#     the spread when its rounding errors happen to cancel).
# The device result must
#   * hold 1e-10 against the binary128 value wherever FP64 can (kappa*eps <= WELL_CONDITIONED), and
#   * elsewhere be no further from it than K_COND times the resolution of the reference arithmetic.
# Evaluations that took different branches than the binary128 one (|dp| >= 0.95, clamps, spline cells) differ by a
# discontinuity of the map, not by rounding: such perturbed evaluations are left out of the spread, and intervals whose
# UNPERTURBED FP64 evaluation branches differently are counted but not compared.
# ---------------------------------------------------------------------------------------------------------------
WELL_CONDITIONED = 1e-11
K_COND = 8.0
N_PERTURB = 8


def reference_resolution(P, tables, X, U, sigma, dt, npts=10, mode=0, n_perturb=N_PERTURB, seed=0, nthreads=0):
    """-> ref64 (FP64 oracle at the inputs), refq (binary128), same (FP64 and binary128 took the same branches),
    kap (n_intervals,) = kappa*eps, the resolution of the reference arithmetic per interval (see above)."""
    from oracle import oracle
    refq, sigq = oracle.linearize_batch_ex(P, tables, X, U, sigma, dt, npts, mode, nthreads=nthreads, precision=1)
    rng = np.random.default_rng(seed)
    ulp = 2.0 ** -52
    kap = np.zeros(sigq.size)
    ref64 = same = None
    for j in range(n_perturb + 1):
        if j == 0:
            Xj, Uj, sj = X, U, sigma
        else:
            Xj = X * (1.0 + ulp * rng.integers(-1, 2, X.shape))
            Uj = U * (1.0 + ulp * rng.integers(-1, 2, U.shape))
            sj = sigma * (1.0 + ulp * rng.integers(-1, 2, sigma.shape))
        b, sg = oracle.linearize_batch_ex(P, tables, Xj, Uj, sj, dt, npts, mode, nthreads=nthreads, precision=0)
        ok = (sg == sigq).reshape(-1)
        m = parity_metric_per_interval(b, refq).max(axis=1)
        kap = np.maximum(kap, np.where(ok & np.isfinite(m), m, 0.0))
        if j == 0:
            ref64, same = b, ok
    return ref64, refq, same, kap


def conditioned_parity(got, refq, same, kap):
    got = np.asarray(got).reshape(-1, 23, 14)
    err = parity_metric_per_interval(got, refq).max(axis=1)
    finite = np.isfinite(np.asarray(refq).reshape(-1, 23 * 14)).all(axis=1)
    use = same & finite
    well = use & (kap <= WELL_CONDITIONED)
    ill = use & ~well
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = np.where(ill, err / np.maximum(kap, 1e-300), 0.0)
    return {"intervals": int(got.shape[0]), "branch_mismatch": int((~same).sum()),
            "non_finite_reference": int((same & ~finite).sum()),
            "well_conditioned": int(well.sum()), "ill_conditioned": int(ill.sum()),
            "max_metric_well_conditioned": float(err[well].max()) if well.any() else 0.0,
            "max_kappa_eps": float(kap[use].max()) if use.any() else 0.0,
            "max_err_over_kappa_eps_ill_conditioned": float(ratio.max()) if ill.any() else 0.0,
            "max_metric_ill_conditioned": float(err[ill].max()) if ill.any() else 0.0}


def assert_conditioned_parity(got, refq, same, kap, tol=PARITY_TOL, k_cond=K_COND):
    rep = conditioned_parity(got, refq, same, kap)
    assert rep["max_metric_well_conditioned"] <= tol, f"well-conditioned intervals miss {tol}: {rep}"
    assert rep["max_err_over_kappa_eps_ill_conditioned"] <= k_cond, f"ill-conditioned intervals exceed {k_cond} x kappa*eps: {rep}"
    assert rep["branch_mismatch"] <= 0.02 * rep["intervals"] + 1, f"too many branch mismatches: {rep}"
    return rep
