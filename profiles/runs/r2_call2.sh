#!/bin/bash
# Round 2, second GPU call: parity suite after the stream fix, bench line, value-kernel occupancy variants.
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -s > $O/r2_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2_pytest_gpu.log
tail -4 $O/r2_pytest_gpu.log
timeout 900 python bench.py > $O/r2_bench_1gpu.json 2> $O/r2_bench_1gpu.err; echo "bench exit $?"; tail -3 $O/r2_bench_1gpu.err
for v in base mb3 mb3_park mb4_park; do
  echo "== variant $v" >> $O/r2_variants.log
  SCVX_B200_LIB=successiveconvexification_b200/variants/libscvx_b200_$v.so timeout 300 python profiles/quick_gpu.py 32768 0 >> $O/r2_variants.log 2>&1
done
cat $O/r2_variants.log
