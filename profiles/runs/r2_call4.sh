#!/bin/bash
# Round 2, fourth GPU call: suite + bench after the NO_Z layout, shared-memory-staged aero tables A/B (timing + ncu).
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -s > $O/r2_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2_pytest_gpu.log
tail -4 $O/r2_pytest_gpu.log
timeout 900 python bench.py > $O/r2_bench_1gpu.json 2> $O/r2_bench_1gpu.err; echo "bench exit $?"; tail -3 $O/r2_bench_1gpu.err
rm -f $O/r2_smem_variants.log
for v in base smem_drag smem_both; do
  echo "== variant $v" >> $O/r2_smem_variants.log
  SCVX_B200_LIB=successiveconvexification_b200/variants/libscvx_b200_$v.so timeout 300 python profiles/quick_gpu.py 32768 0 >> $O/r2_smem_variants.log 2>&1
done
cat $O/r2_smem_variants.log
C="python bench.py --steps 2 --warmup 3 --traj-per-gpu 8192 --no-e2e --no-cpu --no-assembly --no-parity --no-extra"
for v in smem_drag smem_both; do
  export SCVX_B200_LIB=successiveconvexification_b200/variants/libscvx_b200_$v.so
  $C > $O/plain_$v.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:stage_value_kernel -s 4 -c 1 -o $O/r2_value_$v $C > $O/ncu_value_$v.log 2>&1
done
ls -la $O/r2_value_smem*.ncu-rep
