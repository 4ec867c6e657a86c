"""Host-side pieces of the compact result format (include/scvx_b200.h): layout and expander.  No device work."""
import ctypes

import numpy as np

from conftest import structural_constants
from successiveconvexification_b200 import dynamics, workloads
from successiveconvexification_b200.defns import ProbInfo


def test_layout_counts():
    idx = dynamics.compact_layout()
    mask, vals = structural_constants()
    assert idx.shape == (229,) and len(set(idx.tolist())) == 229
    assert int((~mask).sum()) == 229 and int(mask.sum()) == 93
    assert np.array_equal(np.sort(idx), np.flatnonzero(~mask.reshape(-1)))
    # the constants: three unit position columns and D[m, m] = 1
    assert vals.sum() == 4.0 and vals[1, 0] == 1.0 and vals[2, 1] == vals[3, 2] == vals[4, 3] == 1.0


def test_expander_restores_oracle_blocks(prob_aero, oracle_tables):
    """compact(pack of the oracle's blocks) -> expand == the oracle's blocks: the structural constants the expander fills
    in are exactly the ones the reference algorithm produces (dual-number forward mode yields exact 0 / 1 there)."""
    from oracle import oracle
    X, U, sigma, P = workloads.monte_carlo_batch(prob_aero, 9, 11, 5, sigma_range=(0.8, 1.5))
    blocks, err, _, _ = oracle.linearize_batch(P, oracle_tables, X, U, sigma, 0.1)
    mask, vals = structural_constants()
    assert np.array_equal(blocks[..., mask], np.broadcast_to(vals[mask], blocks.shape[:2] + (93,)))
    idx = dynamics.compact_layout()
    comp = np.zeros(blocks.shape[:2] + (230,))
    comp[..., :229] = blocks.reshape(blocks.shape[:2] + (322,))[..., idx]
    comp[1, 2, 229] = 1.0                                  # a flagged interval is counted
    eb, ee, flagged = dynamics.expand_compact(comp, X, n_threads=3)
    assert flagged == 1
    assert np.array_equal(eb, blocks) and np.array_equal(ee, err)
    eb1, _, _ = dynamics.expand_compact(comp, X, lin_err=False, n_threads=1)
    assert np.array_equal(eb1, blocks)
