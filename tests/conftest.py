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


def parity_report(got, ref, floor_frac=PARITY_FLOOR):
    """Parity protocol on (..., 23, 14) blocks.  Per part (endpoint, A, B-, B+, Sigma, z) and per interval:
        max |delta| / max(|ref|, floor_frac * max|part|)
    i.e. 1e-10 RELATIVE per matrix entry for every entry within 4 decades of the part's largest, and an
    ABSOLUTE 1e-14 * max|part| for smaller ones (cancellation zeros).  SURVEY.md §8d proposes a floor of
    1e-12*max|block|; measured here, two independent correct FP64 CPU implementations (dual-number C++ vs
    complex-step numpy) already disagree by 6e-16*max|part| on such entries, i.e. 6e-4 by that metric, so the
    floor is placed where FP64 can resolve it (DESIGN.md "Parity metric").  Returns {part: metric}."""
    got = np.asarray(got).reshape(-1, 23, 14)
    ref = np.asarray(ref).reshape(-1, 23, 14)
    parts = {"endpoint": slice(0, 1), "A": slice(1, 15), "Bm": slice(15, 18), "Bp": slice(18, 21),
             "Sigma": slice(21, 22), "z": slice(22, 23)}
    out = {}
    for name, sl in parts.items():
        g, r = got[:, sl, :], ref[:, sl, :]
        scale = np.abs(r).max(axis=(1, 2), keepdims=True)
        if name == "z":      # z = endpoint - D*inp is formed by cancellation of terms of the size of D and endpoint
            scale = np.abs(ref[:, 0:22, :]).max(axis=(1, 2), keepdims=True)
        den = np.maximum(np.abs(r), floor_frac * scale)
        den[den == 0.0] = 1.0
        out[name] = float((np.abs(g - r) / den).max()) if g.size else 0.0
    return out


PARITY_TOL = 1e-10     # BASELINE.json north_star: 1e-10 relative per matrix entry


def assert_parity(got, ref, tol=PARITY_TOL):
    rep = parity_report(got, ref)
    bad = {k: v for k, v in rep.items() if not (v <= tol)}
    assert not bad, f"parity failed: {rep}"
    return rep
