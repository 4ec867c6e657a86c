#!/bin/bash
for f in variants/*.so; do
  cp $f successiveconvexification_b200/libscvx_b200.so
  echo "== $f"; C="python bench.py --steps 1 --warmup 3 --traj-per-gpu 4096 --no-e2e --no-cpu --kernel 3"; $C > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 6 --csv --log-file gpurun_out/lv.csv $C > /dev/null 2>&1; grep -v "^==" gpurun_out/lv.csv | awk -F'","' '{print $5, $NF}' | tail -3
done
