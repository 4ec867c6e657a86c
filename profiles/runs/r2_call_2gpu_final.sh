#!/bin/bash
# final build on two GPUs: torchrun bench line (trajectory shards, NCCL status reduce + gather_to_root slab) and the
# multi-device context test
set -u
O=gpurun_out; mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2_bench_2gpu.json 2> $O/r2_bench_2gpu.err; echo "bench2 exit $?"
cat $O/r2_bench_2gpu.json | cut -c1-400
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_device" > $O/r2_pytest_multi_device.log 2>&1; tail -2 $O/r2_pytest_multi_device.log
