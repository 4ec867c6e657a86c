#!/bin/bash
set -u
O=gpurun_out
rm -f $O/r2_occ_variants2.log
for v in base mb3 mb3_park mb4_park; do
  echo "== variant $v" >> $O/r2_occ_variants2.log
  SCVX_B200_LIB=successiveconvexification_b200/variants/libscvx_b200_$v.so timeout 300 python profiles/quick_gpu.py 32768 0 >> $O/r2_occ_variants2.log 2>&1
done
cat $O/r2_occ_variants2.log
