"""Build the CPU oracle shared library (TEST INFRASTRUCTURE, not product code).

    python oracle/build.py        ->  oracle/libscvx_oracle.so

`-ffp-contract=off` keeps the arithmetic the plain IEEE-754 double operations the Julia
reference executes (no fused multiply-add), `-fopenmp` lets bench.py's cpu_baseline /
--impl reference legs use every host core.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "scvx_oracle.cpp")
OUT = os.path.join(HERE, "libscvx_oracle.so")


def build(force: bool = False) -> str:
    if (not force) and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["g++", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-std=c++17",
           "-Wall", "-o", OUT, SRC, "-lquadmath"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
