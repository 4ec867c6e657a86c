#!/bin/bash
# mbarrier waits of the tangent kernel: suspend-time hint on try_wait / back-off of the producers' wait for a free ring slot
set -u
O=gpurun_out; mkdir -p $O
L=$O/r2_ab_mbar_wait2.log; : > $L
for rep in 1 2; do
  for v in producer_sleep200 producer_sleep500 producer_sleep3000 sleep500_rec100 sleep500_rec400; do
    echo "== variant $v" >> $L
    SCVX_B200_LIB=successiveconvexification_b200/variants/libscvx_b200_$v.so timeout 120 python profiles/quick_gpu.py 32768 0 >> $L 2>&1 || echo "FAILED rc=$?" >> $L
  done
done
cat $L
