#!/bin/bash
# end of the round: headline bench (roofline of the k-space kernel alone), large-K capture + bench, S2 bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( time python bench.py --steps 10 --warmup 3 --no-extras ) > gpurun_out/w_bench_s1.log 2>&1
tail -c 900 gpurun_out/w_bench_s1.log
timeout 900 ncu --set full --clock-control none -k regex:"windowKspaceKernel|windowFrontKernel" -s 4 -c 4 \
    -o gpurun_out/r02n_largeK python scripts/profile_moves.py 500 s1-largeK > gpurun_out/w_ncu.log 2>&1
tail -2 gpurun_out/w_ncu.log
( time python bench.py --steps 2 --warmup 3 --workload s1-largeK ) > gpurun_out/w_bench_largeK.log 2>&1
tail -c 600 gpurun_out/w_bench_largeK.log
( time python bench.py --steps 3 --warmup 3 --workload s2 ) > gpurun_out/w_bench_s2.log 2>&1
tail -c 600 gpurun_out/w_bench_s2.log
