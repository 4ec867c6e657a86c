#!/bin/bash
# intervals per SM and chunk (768 = 3 waves of the value kernel / 24 passes of the tangent kernel)
set -u
O=gpurun_out; mkdir -p $O
L=$O/r2_ab_chunk.log; : > $L
for rep in 1 2; do
  for v in 768 1536 2304 3072 512; do
    echo "== SCVX_CHUNK_PER_SM=$v" >> $L
    SCVX_CHUNK_PER_SM=$v timeout 120 python profiles/quick_gpu.py 32768 0 >> $L 2>&1 || echo "FAILED rc=$?" >> $L
  done
done
cat $L
