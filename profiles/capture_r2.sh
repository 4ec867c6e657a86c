#!/bin/bash
# Round-2 profile capture (run under gpurun on one B200).  Follows /opt/skills/guides/B200_PROFILING.md: every ncu run is
# preceded by the same command exiting 0 without ncu; numbers printed under ncu are never bench values.
set -u
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r2_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2_pytest_gpu.log
tail -3 $O/r2_pytest_gpu.log
python __graft_entry__.py --smoke > $O/r2_smoke.log 2>&1; tail -1 $O/r2_smoke.log
python bench.py > $O/r2_bench.json 2> $O/r2_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference.json 2>> $O/r2_bench.err
python bench.py --mode 1 --no-cpu --no-extra > $O/r2_bench_textbook.json 2>> $O/r2_bench.err
python bench.py --sigma-range 0.8,1.5 --no-cpu --no-extra > $O/r2_bench_sigma_near_1.json 2>> $O/r2_bench.err
C="python bench.py --steps 2 --warmup 3 --traj-per-gpu 8192 --no-e2e --no-cpu --no-assembly --no-parity --no-extra"
$C > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 23 -c 24 --csv --log-file $O/r2_launches.csv $C > $O/ncu_launches.log 2>&1
for k in tangent_kernel stage_value_kernel; do
  $C > $O/plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o $O/r2_$k $C > $O/ncu_$k.log 2>&1
done
C2="python bench.py --steps 1 --warmup 3 --traj-per-gpu 8192 --no-cpu --no-assembly --no-parity --no-extra"
$C2 > $O/plain2.log 2>&1 &&
ncu --set full --clock-control none -k regex:compact_pack_kernel -s 2 -c 1 -o $O/r2_compact_pack_kernel $C2 > $O/ncu_pack.log 2>&1
ls -la $O/r2_*.ncu-rep
