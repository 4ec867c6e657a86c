"""Build libscvx_b200.so (the C-ABI shared library of include/scvx_b200.h) for sm_100a, in-tree.

    python successiveconvexification_b200/csrc/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
working tree.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libscvx_b200.so")
SOURCES = ["scvx_api.cu", "scvx_kernels_basic.cu", "scvx_kernels_staged.cu", "scvx_kernels_socp.cu",
           "scvx_kernels_compact.cu"]
# measurement helpers of bench.py / profiles/ (DFMA peak, DMMA probe): a separate library, not part of the product ABI
TOOLS_OUT = os.path.join(PKG, "libscvx_benchtools.so")
TOOLS_SOURCES = ["bench_tools.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--fmad=true", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _newest_source():
    deps = glob.glob(os.path.join(HERE, "*.cu")) + glob.glob(os.path.join(HERE, "*.cuh")) + \
        glob.glob(os.path.join(HERE, "*.h")) + glob.glob(os.path.join(PKG, "..", "include", "*.h"))
    return max(os.path.getmtime(p) for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    _build(SOURCES, OUT, force, verbose)
    _build(TOOLS_SOURCES, TOOLS_OUT, force, verbose)
    return OUT


def _build(sources, out, force, verbose):
    OUT = out
    srcs = [os.path.join(HERE, s) for s in sources]
    if (not force) and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest_source():
        return OUT
    objs = []
    for s in srcs:
        o = os.path.splitext(s)[0] + ".o"
        extra = os.environ.get("SCVX_NVCC_EXTRA", "").split()
        cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        subprocess.check_call(cmd)
        objs.append(o)
    subprocess.check_call([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
