#!/bin/bash
# value kernel: four-stage loop unrolled by 2 / by 4 (one basic block spanning stages)
set -u
O=gpurun_out; mkdir -p $O
L=$O/r2_ab_stage_unroll.log; : > $L
for v in base stage_unroll2 stage_unroll4; do
  for mode in 0 1; do
    echo "== variant $v mode $mode" >> $L
    SCVX_B200_LIB=successiveconvexification_b200/variants/libscvx_b200_$v.so timeout 60 python profiles/quick_gpu.py 32768 $mode >> $L 2>&1 || echo "FAILED rc=$?" >> $L
  done
done
cat $L
