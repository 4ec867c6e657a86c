#!/bin/bash
# mid-round ncu capture of the tangent kernel (source view) after the consumer stage-body changes
set -u
O=gpurun_out; mkdir -p $O
C="python bench.py --steps 2 --warmup 3 --traj-per-gpu 8192 --no-e2e --no-cpu --no-assembly --no-parity --no-extra"
$C > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tangent_kernel -s 4 -c 1 -f -o $O/r2b_tangent_kernel $C > $O/ncu_tangent_b.log 2>&1
ls -la $O/r2b_tangent_kernel.ncu-rep
