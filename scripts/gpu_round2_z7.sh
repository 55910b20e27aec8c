#!/bin/bash
# round 2, late: 4-GPU bench (replicas + the sharded extras with the full-Q product in the slabs)
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/z7_bench_4gpu.log 2>&1
echo "rc=$?" >> gpurun_out/z7_bench_4gpu.log
tail -c 900 gpurun_out/z7_bench_4gpu.log
