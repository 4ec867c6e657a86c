#!/bin/bash
# Round 2, 8-GPU call: weak-scaling bench line (device-resident + compact end to end), NCCL result collection check,
# multi-device context test, aggregate D2H probe.
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus 8 --steps 10 --warmup 3 > $O/r2_bench_8gpu.json 2> $O/r2_bench_8gpu.err; echo "bench8 exit $?"; tail -2 $O/r2_bench_8gpu.err
$TR profiles/nccl_gather_check.py > $O/r2_nccl_gather_check.txt 2>&1; tail -2 $O/r2_nccl_gather_check.txt
$TR profiles/d2h_aggregate.py > $O/r2_d2h_aggregate.txt 2>&1; tail -3 $O/r2_d2h_aggregate.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_device" > $O/r2_pytest_multi_device.log 2>&1; tail -2 $O/r2_pytest_multi_device.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2_bench_2gpu.json 2> $O/r2_bench_2gpu.err; echo "bench2 exit $?"
