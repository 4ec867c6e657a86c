"""The case compute-sanitizer is run on (profiles/r2_*check.txt hold the summaries):

    compute-sanitizer --tool racecheck|synccheck|memcheck python -m pytest tests/test_sanitizer_case.py -m gpu -q

Small enough for a sanitizer run, large enough to put several 32-interval passes through one CTA of tangent_kernel
(mbarrier ring two steps deep, single-buffered TMA staging reused after a __syncwarp, setmaxnreg re-partition,
cross-pass phase-parity bookkeeping) for npts = 1, 3 and 10 in both stage rules, plus the compact pack kernel and the
predict / defect-cost / SOCP-value kernels.  Every result is checked against the CPU oracle, so a clean sanitizer run
is also a correct run.  Part of the normal -m gpu suite as well."""
import numpy as np
import pytest

from conftest import assert_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("npts,B", [(1, 190), (3, 190), (10, 100)])
def test_staged_path_small(prob_aero, oracle_tables, npts, B):
    from oracle import oracle
    from successiveconvexification_b200 import dynamics as dyn, rocketland, workloads
    cache = dyn.make_cache(prob_aero)
    cache.sim_prob.set_kernel(dyn.KERNEL_STAGED)
    K = 50                                              # B * 50 intervals = 297 / 157 groups on <= 148 CTAs: 2-3 passes per CTA
    X, U, s, P = workloads.monte_carlo_batch(prob_aero, K, B, 17, sigma_range=(0.8, 1.5))
    for mode in (0, 1):
        blocks, err, tlb = dyn.linearize_batch(cache, X, U, s, 1 / (K + 1), npts, mode)
        ref, rerr, rtlb, _ = oracle.linearize_batch(P, oracle_tables, X, U, s, 1 / (K + 1), npts, mode)
        assert_parity(blocks, ref)
        assert np.abs(err - rerr).max() <= 1e-12 * max(1.0, np.abs(rerr).max())
        assert np.abs(tlb - rtlb).max() <= 1e-15
        comp, _ = dyn.linearize_batch_compact(cache, X, U, s, 1 / (K + 1), npts, mode)
        eb, ee, flagged = dyn.expand_compact(comp, X)
        assert flagged == 0 and np.array_equal(eb, blocks) and np.array_equal(ee, err)
    end = dyn.predict_batch(cache, X, U, s, 1 / (K + 1), npts, 1)
    assert np.abs(end - ref[:, :, 0, :]).max() <= 1e-12 * np.abs(ref[:, :, 0, :]).max()
    defect, cost = dyn.defect_cost(cache, X, err, prob_aero.wNu)
    assert defect == pytest.approx(np.sqrt((err ** 2).sum(axis=(1, 2))), rel=1e-12)
    vals, const = rocketland.socp_values_batch(cache, blocks, err, tlb)
    assert np.isfinite(vals).all() and np.array_equal(const[:, :14 * K], err.reshape(B, -1))
    cache.sim_prob.close()
