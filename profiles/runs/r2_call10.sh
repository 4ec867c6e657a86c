#!/bin/bash
set -u
O=gpurun_out
rm -f $O/r2_carveout_variants.log
for v in base carve25 carve15; do
  echo "== variant $v" >> $O/r2_carveout_variants.log
  SCVX_B200_LIB=successiveconvexification_b200/variants/libscvx_b200_$v.so timeout 300 python profiles/quick_gpu.py 32768 0 >> $O/r2_carveout_variants.log 2>&1
done
cat $O/r2_carveout_variants.log
