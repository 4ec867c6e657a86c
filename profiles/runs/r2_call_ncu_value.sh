#!/bin/bash
# ncu capture of the value kernel of the final build (branch-free lift terms)
set -u
O=gpurun_out; mkdir -p $O
C="python bench.py --steps 2 --warmup 3 --traj-per-gpu 8192 --no-e2e --no-cpu --no-assembly --no-parity --no-extra"
$C > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stage_value_kernel -s 4 -c 1 -f -o $O/r2_stage_value_kernel $C > $O/ncu_stage_value_kernel.log 2>&1
ls -la $O/r2_stage_value_kernel.ncu-rep
