#!/bin/bash
# Round 2, sixth GPU call: tangent-kernel slot roles (control / state columns) — suite + timing.
set -u
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2_pytest_gpu.log
tail -4 $O/r2_pytest_gpu.log
python profiles/quick_gpu.py 32768 0 > $O/r2_quick_slots.log 2>&1; cat $O/r2_quick_slots.log
python profiles/quick_gpu.py 32768 1 >> $O/r2_quick_slots.log 2>&1; tail -4 $O/r2_quick_slots.log
