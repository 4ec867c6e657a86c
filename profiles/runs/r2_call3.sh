#!/bin/bash
# Round 2, third GPU call: ncu baseline of the round (launch list + full sets) and the value-kernel occupancy A/B.
set -u
O=gpurun_out
C="python bench.py --steps 2 --warmup 3 --traj-per-gpu 8192 --no-e2e --no-cpu --no-assembly --no-parity --no-extra"
$C > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 23 -c 24 --csv --log-file $O/r2_launches.csv $C > $O/ncu_launches.log 2>&1
for k in tangent_kernel stage_value_kernel; do
  $C > $O/plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o $O/r2_$k $C > $O/ncu_$k.log 2>&1
done
export SCVX_B200_LIB=successiveconvexification_b200/variants/libscvx_b200_mb3_park.so
$C > $O/plain_mb3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stage_value_kernel -s 4 -c 1 -o $O/r2_value_mb3_park $C > $O/ncu_value_mb3.log 2>&1
ls -la $O/*.ncu-rep | tail -5
