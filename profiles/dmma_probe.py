"""The DMMA experiment (BASELINE.json north_star: "tensor cores (DMMA) only if ncu shows benefit"): can the FP64 tensor
path add throughput to the FP64 FMA pipe on B200, or do the two share one datapath?  Prints one JSON line.
    python profiles/dmma_probe.py"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from successiveconvexification_b200 import _lib
tools = _lib.load_benchtools()
out = (ctypes.c_double * 8)()
rc = tools.scvx_bench_dmma_probe(0, out)
assert rc == 0, rc
tf = ctypes.c_double(); tools.scvx_bench_fp64_peak(0, ctypes.byref(tf))
print(json.dumps({"dfma_only_tf": out[0], "dmma_m8n8k4_only_tf": out[1], "dmma_m16n8k8_only_tf": out[2],
                  "alternate_warps": {"dfma_tf": out[3], "dmma_tf": out[4], "sum_tf": out[3] + out[4]},
                  "same_warp_interleaved": {"dfma_tf": out[5], "dmma_tf": out[6], "sum_tf": out[5] + out[6]},
                  "fp64_peak_microbenchmark_tf": tf.value}))
