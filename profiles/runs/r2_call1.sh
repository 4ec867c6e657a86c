#!/bin/bash
# Round 2, first GPU call: parity suite, bench line, DMMA probe, racecheck of the STAGED path.
set -u
O=gpurun_out
mkdir -p $O
SCVX_DUMP=1 timeout 1500 python -m pytest tests -m gpu -q -s > $O/r2_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2_pytest_gpu.log
tail -5 $O/r2_pytest_gpu.log
timeout 900 python bench.py > $O/r2_bench_1gpu.json 2> $O/r2_bench_1gpu.err; echo "bench exit $?"
python profiles/dmma_probe.py > $O/r2_dmma_probe.json 2> $O/r2_dmma_probe.err; cat $O/r2_dmma_probe.json
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_sanitizer_case.py -m gpu -q -x > $O/r2_racecheck.log 2>&1; echo "racecheck exit $?"
tail -5 $O/r2_racecheck.log
