#!/bin/bash
# which of the two programmatic launches costs: SCVX_PDL = 0 none, 1 value kernels only, 2 tangent kernels only, 3 both
set -u
O=gpurun_out; mkdir -p $O
L=$O/r2_ab_pdl4.log; : > $L
for v in 0 2 3 0 2 3; do
  echo "== SCVX_PDL=$v mode=LITERAL" >> $L
  SCVX_PDL=$v timeout 300 python profiles/quick_gpu.py 32768 0 >> $L 2>&1 || echo "FAILED rc=$?" >> $L
done
cat $L
