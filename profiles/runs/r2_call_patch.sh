#!/bin/bash
# value kernel reading the spline coefficients from the patch-expanded table (one 128-byte line per cell and table)
set -u
O=gpurun_out; mkdir -p $O
L=$O/r2_ab_branchless_lift.log; : > $L
for rep in 1 2; do
  for v in base branchless_lift; do
    for mode in 0 1; do
      echo "== variant $v mode $mode" >> $L
      SCVX_B200_LIB=successiveconvexification_b200/variants/libscvx_b200_$v.so timeout 120 python profiles/quick_gpu.py 32768 $mode >> $L 2>&1 || echo "FAILED rc=$?" >> $L
    done
  done
done
cat $L
