"""Build kernel variants of libscvx_b200.so for A/B runs on the GPU box (compile-time switches of the STAGED kernels).

    python profiles/build_variants.py            ->  successiveconvexification_b200/variants/libscvx_b200_<name>.so
    SCVX_B200_LIB=successiveconvexification_b200/variants/libscvx_b200_<name>.so python profiles/quick_gpu.py

Only scvx_kernels_staged.cu is recompiled per variant; the other objects come from the default build."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from successiveconvexification_b200.csrc import build as b

VARIANTS = {
    "base": [],
    "tables_global": ["-DSCVX_A_SMEM_TABLES=0"],
    "evict_last": ["-DSCVX_A_EVICT_LAST=1"],
    "evict_last_global": ["-DSCVX_A_EVICT_LAST=1", "-DSCVX_A_SMEM_TABLES=0"],
    "mb3": ["-DSCVX_A_MINBLOCKS=3"],
    "mb3_park": ["-DSCVX_A_MINBLOCKS=3", "-DSCVX_A_PARK=1"],
    "mb4_park": ["-DSCVX_A_MINBLOCKS=4", "-DSCVX_A_PARK=1"],
    "lean_lift": ["-DSCVX_A_LEAN_LIFT=1"],
    "carve25": ["-DSCVX_A_CARVEOUT=25"],
    "carve15": ["-DSCVX_A_CARVEOUT=15"],
    "smem_drag": ["-DSCVX_A_SMEM_TABLES=1"],
    "smem_both": ["-DSCVX_A_SMEM_TABLES=2"],
    "smem_window": ["-DSCVX_A_SMEM_TABLES=3"],
    "hoist_lift": ["-DSCVX_A_HOIST_LIFT=1"],
    "first_body": ["-DSCVX_T_FIRST_BODY=1"],
    "no_first_body": ["-DSCVX_T_FIRST_BODY=0"],
    "four_bodies": ["-DSCVX_T_FIRST_BODY=2"],
    "step_unroll2": ["-DSCVX_T_STEP_UNROLL=2"],
    "split_reduce": ["-DSCVX_T_SPLIT_REDUCE=1"],
    "prefetch_epilogue": ["-DSCVX_A_PREFETCH_EPILOGUE=1"],
    "vt64": ["-DSCVX_A_VT=64", "-DSCVX_A_MINBLOCKS=4"],
    "vt256": ["-DSCVX_A_VT=256", "-DSCVX_A_MINBLOCKS=1"],
    "prefetch_pass": ["-DSCVX_T_PREFETCH_PASS=1"],
    "branchless_lift": ["-DSCVX_A_BRANCHLESS_LIFT=1"],
    "stage_unroll2": ["-DSCVX_A_STAGE_UNROLL=2"],
    "stage_unroll4": ["-DSCVX_A_STAGE_UNROLL=4"],
    "mbar_hint": ["-DSCVX_MBAR_HINT=10000000"],
    "mbar_hint_1us": ["-DSCVX_MBAR_HINT=1000"],
    "producer_sleep200": ["-DSCVX_PRODUCER_SLEEP_NS=200"],
    "producer_sleep1000": ["-DSCVX_PRODUCER_SLEEP_NS=1000"],
    "producer_sleep500": ["-DSCVX_PRODUCER_SLEEP_NS=500"],
    "producer_sleep3000": ["-DSCVX_PRODUCER_SLEEP_NS=3000"],
    "sleep500_rec100": ["-DSCVX_PRODUCER_SLEEP_NS=500", "-DSCVX_RECORD_SLEEP_NS=100"],
    "sleep500_rec400": ["-DSCVX_PRODUCER_SLEEP_NS=500", "-DSCVX_RECORD_SLEEP_NS=400"],
}


def main(names):
    b.build()
    out_dir = os.path.join(b.PKG, "variants")
    os.makedirs(out_dir, exist_ok=True)
    objs = [os.path.join(b.HERE, s[:-3] + ".o") for s in b.SOURCES if s != "scvx_kernels_staged.cu"]
    for name in names:
        flags = VARIANTS[name]
        o = os.path.join(out_dir, f"staged_{name}.o")
        subprocess.check_call([b._nvcc()] + b.NVCC_FLAGS + flags + ["-c", os.path.join(b.HERE, "scvx_kernels_staged.cu"), "-o", o])
        so = os.path.join(out_dir, f"libscvx_b200_{name}.so")
        subprocess.check_call([b._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so] + objs + [o])
        print(so)


if __name__ == "__main__":
    main(sys.argv[1:] or list(VARIANTS))
