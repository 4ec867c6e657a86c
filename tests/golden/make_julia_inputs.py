"""Write the inputs tools/gen_golden.jl feeds to the reference's own Julia code (plain binary + text manifest, so Julia
needs no extra package to read them).

    python tests/golden/make_julia_inputs.py

  julia_inputs.f64 : Float64 little-endian stream; per case X (14 x n_nodes x B), U (3 x n_nodes x B), sigma (B) in
                     Julia (column-major) order = the C order of the (B, n_nodes, 14) arrays used everywhere here
  julia_inputs.txt : one line per case  `name B n_nodes dt offset_X offset_U offset_sigma`  (offsets in doubles)
All cases use the aero-table sample problem `SampleProblems.base_prob_aero_scaled` (shared parameters: the reference's
ProbInfo has no constructor other than from a DescentProblem) and the reference's LITERAL rk4.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from successiveconvexification_b200 import sample_problems as sp, workloads      # noqa: E402


def cases():
    prob = sp.base_prob_aero_scaled(os.path.join(HERE, "aero_lift_drag.npz"))
    X, U, sigma, dt = workloads.sample_trajectory(prob)                                     # C2: 51 nodes, sigma = 1
    yield "c2", X, U, sigma, dt
    X, U, sigma, _ = workloads.monte_carlo_batch(prob, 8, 6, 777, sigma_range=(0.8, 1.5))   # both |dp| branches
    yield "mc_sigma_near_1", X, U, sigma, 1.0 / 9.0
    X, U, sigma, _ = workloads.monte_carlo_batch(prob, 6, 4, 778, sigma_range=(1.0, 15.0))  # the bench's sigma range
    yield "mc_sigma_1_15", X, U, sigma, 1.0 / 7.0


def build():
    chunks, lines, off = [], [], 0
    for name, X, U, sigma, dt in cases():
        B, n, _ = X.shape
        oX, oU, oS = off, off + X.size, off + X.size + U.size
        off = oS + sigma.size
        chunks += [X.reshape(-1), U.reshape(-1), sigma.reshape(-1)]
        lines.append(f"{name} {B} {n} {dt!r} {oX} {oU} {oS}")
    return np.concatenate(chunks).astype("<f8"), "\n".join(lines) + "\n"


if __name__ == "__main__":
    data, manifest = build()
    data.tofile(os.path.join(HERE, "julia_inputs.f64"))
    open(os.path.join(HERE, "julia_inputs.txt"), "w").write(
        "# inputs of tools/gen_golden.jl: name B n_nodes dt offset_X offset_U offset_sigma (offsets in doubles)\n" + manifest)
    print(len(data), "doubles")
