#!/bin/bash
# per-trajectory sweep path (shared record + per-trajectory a / Tmin): GPU suite, then the bench line with `extra` (C4)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2_sweep_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_sweep_pytest.log
tail -5 gpurun_out/r2_sweep_pytest.log
python bench.py > gpurun_out/r2_sweep_bench.json 2> gpurun_out/r2_sweep_bench.err; echo "bench rc=$?"
cat gpurun_out/r2_sweep_bench.json
