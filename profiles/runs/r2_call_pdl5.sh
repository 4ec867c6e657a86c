#!/bin/bash
# tangent kernel started on the SMs the value kernel's last wave leaves: SCVX_PDL = 0 plain launches, 1 programmatic
# launch + grid-wide wait, 2 programmatic launch + per-block record flags.  Then the GPU suite with the default (2).
set -u
O=gpurun_out; mkdir -p $O
L=$O/r2_ab_pdl5.log; : > $L
for v in 0 1 2 0 1 2; do
  echo "== SCVX_PDL=$v mode=LITERAL" >> $L
  SCVX_PDL=$v timeout 120 python profiles/quick_gpu.py 32768 0 >> $L 2>&1 || echo "FAILED rc=$?" >> $L
done
for v in 0 2; do
  echo "== SCVX_PDL=$v mode=TEXTBOOK" >> $L
  SCVX_PDL=$v timeout 120 python profiles/quick_gpu.py 32768 1 >> $L 2>&1 || echo "FAILED rc=$?" >> $L
done
cat $L
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2_pdl_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2_pdl_pytest.log
tail -5 $O/r2_pdl_pytest.log
