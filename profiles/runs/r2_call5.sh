#!/bin/bash
# Round 2, fifth GPU call: suite (fins variant, fin tables), bench with the shared-parameter kernels.
set -u
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > $O/r2_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2_pytest_gpu.log
tail -4 $O/r2_pytest_gpu.log
python profiles/quick_gpu.py 32768 0 > $O/r2_quick_sp.log 2>&1; cat $O/r2_quick_sp.log
python profiles/quick_gpu.py 32768 1 >> $O/r2_quick_sp.log 2>&1; tail -4 $O/r2_quick_sp.log
timeout 900 python bench.py > $O/r2_bench_1gpu.json 2> $O/r2_bench_1gpu.err; echo "bench exit $?"; tail -3 $O/r2_bench_1gpu.err
