#!/bin/bash
# round 2, end: 8-GPU bench (replicas + the sharded extras)
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/z13_bench_8gpu.log 2>&1
echo "rc=$?" >> gpurun_out/z13_bench_8gpu.log
tail -c 700 gpurun_out/z13_bench_8gpu.log
