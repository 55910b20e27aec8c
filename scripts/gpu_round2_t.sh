#!/bin/bash
# multi-GPU bench line (replicas + sharded extras) at N ranks: bash scripts/gpu_round2_t.sh N
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
N=${1:-4}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/t_bench_${N}gpu.log 2>&1
tail -c 1800 gpurun_out/t_bench_${N}gpu.log
