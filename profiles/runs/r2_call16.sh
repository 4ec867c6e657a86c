#!/bin/bash
set -u
O=gpurun_out
rm -f $O/r2_ab3.log
for v in base evict_last evict_last_global base evict_last; do
  echo "== variant $v" >> $O/r2_ab3.log
  SCVX_B200_LIB=successiveconvexification_b200/variants/libscvx_b200_$v.so timeout 300 python profiles/quick_gpu.py 32768 0 >> $O/r2_ab3.log 2>&1
done
cat $O/r2_ab3.log
