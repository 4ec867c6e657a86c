#!/bin/bash
# Round-1 profile capture (run under gpurun on one B200).  Follows /opt/skills/guides/B200_PROFILING.md:
# every ncu run is preceded by the same command exiting 0 without ncu; numbers printed under ncu are never bench values.
set -u
C="python bench.py --steps 2 --warmup 3 --traj-per-gpu 8192 --no-e2e --no-cpu --no-assembly"
O=gpurun_out
$C > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 23 -c 24 --csv --log-file $O/r1_launches.csv $C > $O/ncu_launches.log 2>&1
for k in tangent_kernel stage_value_kernel; do
  $C > $O/plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o $O/r1_$k $C > $O/ncu_$k.log 2>&1
done
python bench.py > $O/r1_bench.json 2> $O/r1_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r1_bench_reference.json 2>> $O/r1_bench.err
