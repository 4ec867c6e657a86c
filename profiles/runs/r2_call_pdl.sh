#!/bin/bash
# programmatic dependent launch of the value / tangent kernels over two record buffers: A/B (SCVX_PDL=0 serialises
# the launches in full) on one box, then the GPU suite with the overlap on
set -u
O=gpurun_out; mkdir -p $O
L=$O/r2_ab_pdl.log; : > $L
for rep in 1 2; do
  for v in 0 1; do
    echo "== SCVX_PDL=$v mode=LITERAL" >> $L
    SCVX_PDL=$v timeout 300 python profiles/quick_gpu.py 32768 0 >> $L 2>&1 || echo "FAILED rc=$?" >> $L
  done
done
for v in 0 1; do
  echo "== SCVX_PDL=$v mode=TEXTBOOK" >> $L
  SCVX_PDL=$v timeout 300 python profiles/quick_gpu.py 32768 1 >> $L 2>&1 || echo "FAILED rc=$?" >> $L
done
cat $L
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2_pdl_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2_pdl_pytest.log
tail -5 $O/r2_pdl_pytest.log
