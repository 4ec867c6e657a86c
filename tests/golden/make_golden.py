"""Generate the golden vectors of the hot path with the CPU oracle.

    python tests/golden/make_golden.py

PARITY UNPINNED by the reference: it ships no tests or known-answer vectors for this path and Julia
is not installed here, so these vectors are produced by oracle/scvx_oracle.cpp,
which tests/test_oracle.py pins by independent means (complex-step numpy restatement, scipy natural
spline, finite differences, the survey's anchors).  Inputs are regenerated from seeds by the tests;
only outputs are stored.

  golden_c2.npz : configuration C2 (sample problem base_prob_aero_scaled, K=50 -> 51 nodes, sigma=1,
                  dt=1/51) — blocks (50, 23, 14) for {aero, exo} x {LITERAL, TEXTBOOK}
  golden_mc.npz : small Monte-Carlo sweep batch (seed 4242, B=6, K=5, per-trajectory parameters,
                  aero, LITERAL, sigma ~ U(0.8, 1.5)) — blocks, lin_err, thrust-lower-bound rows
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle                                                      # noqa: E402
from successiveconvexification_b200 import sample_problems as sp, workloads    # noqa: E402
from successiveconvexification_b200.defns import ExoatmosphericData, ProbInfo  # noqa: E402


def main():
    aero_npz = os.path.join(HERE, "aero_lift_drag.npz")
    prob = sp.base_prob_aero_scaled(aero_npz)
    tb = oracle.OracleTables.from_aero(prob.aero)
    X, U, sigma, dt = workloads.sample_trajectory(prob)
    out = {}
    for aname, info, tables in (("aero", ProbInfo(prob), tb),
                                ("exo", ProbInfo(prob.replace(aero=ExoatmosphericData())), None)):
        for mname, mode in (("literal", 0), ("textbook", 1)):
            blocks, _, _, _ = oracle.linearize_batch(info, tables, X, U, sigma, dt, 10, mode, False, False)
            out[f"{aname}_{mname}"] = blocks[0]
    np.savez_compressed(os.path.join(HERE, "golden_c2.npz"), **out)

    Xm, Um, sm, Pm = workloads.monte_carlo_batch(prob, 5, 6, 4242, sweep=True, sigma_range=(0.8, 1.5))
    blocks, err, tlb, _ = oracle.linearize_batch(Pm, tb, Xm, Um, sm, 1.0 / 6.0, 10, 0, True, True)
    np.savez_compressed(os.path.join(HERE, "golden_mc.npz"), blocks=blocks, lin_err=err, tlb=tlb)
    for f in ("golden_c2.npz", "golden_mc.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
