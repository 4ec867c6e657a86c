#!/bin/bash
set -u
O=gpurun_out
rm -f $O/r2_ab4.log
for m in 0 1; do for v in base tables_global base tables_global; do
  echo "== variant $v mode $m" >> $O/r2_ab4.log
  SCVX_B200_LIB=successiveconvexification_b200/variants/libscvx_b200_$v.so timeout 300 python profiles/quick_gpu.py 32768 $m >> $O/r2_ab4.log 2>&1
done; done
cat $O/r2_ab4.log
