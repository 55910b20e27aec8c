#!/bin/bash
# bench: headline with extras, large-K variant, S2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/l_bench_s1.log 2>&1
( time python bench.py --steps 2 --warmup 3 --workload s1-largeK ) > gpurun_out/l_bench_largeK.log 2>&1
( time python bench.py --steps 3 --warmup 3 --workload s2 ) > gpurun_out/l_bench_s2.log 2>&1
tail -c 1200 gpurun_out/l_bench_s1.log; tail -c 600 gpurun_out/l_bench_largeK.log; tail -c 600 gpurun_out/l_bench_s2.log
