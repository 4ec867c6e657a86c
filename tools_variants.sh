#!/bin/bash
# run tools_quick_gpu.py for every prebuilt variant library under variants/
for f in variants/*.so; do
  cp $f successiveconvexification_b200/libscvx_b200.so
  echo "== $f"; timeout 120 python tools_quick_gpu.py 2>&1 | grep "kernel 3" | tail -1
done
