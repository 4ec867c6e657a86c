"""Generate tests/golden/aero_lift_drag.npz from the reference's aero table.

The reference's `aero/lift_drag.csv` (header `aoa,mach,drag,lift,torque`; 181 cos(aoa) samples,
fastest-varying, x 61 Mach samples; consumed at aerodynamics.jl:12-21) is the only data fixture
the hot path has.  /root/reference does not exist on the GPU box, so the three sample columns are
stored here as float64 arrays in CSV row order (bit-exact: the CSV decimal strings are parsed
once, by numpy, exactly as `load_aerodata` would parse them).

    python tests/golden/make_aero_fixture.py [/root/reference/aero/lift_drag.csv]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/aero/lift_drag.csv"
    with open(src) as fh:
        header = fh.readline().strip().split(",")
    raw = np.loadtxt(src, delimiter=",", skiprows=1, dtype=np.float64)
    assert raw.shape == (181 * 61, 5), raw.shape
    cols = {n: raw[:, header.index(n)].copy() for n in ("aoa", "mach", "drag", "lift", "torque")}
    # grid sanity: aoa column is cos(aoa) = -1:1/90:1 (fastest), mach = 0:0.025:1.5
    assert np.allclose(cols["aoa"][:181], -1 + np.arange(181) / 90.0, atol=1e-12)
    assert np.allclose(cols["mach"][::181], np.arange(61) * 0.025, atol=1e-12)
    out = os.path.join(HERE, "aero_lift_drag.npz")
    np.savez_compressed(out, drag=cols["drag"], lift=cols["lift"], torque=cols["torque"])
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
