#!/bin/bash
# what the driver runs at the end of a round: GPU tests, smoke, the bench line (N = 1) and the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; tail -3 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
( time python bench.py ) > gpurun_out/final_bench.log 2>&1; tail -c 700 gpurun_out/final_bench.log
( time python bench.py --impl reference ) > gpurun_out/final_bench_reference.log 2>&1; tail -c 900 gpurun_out/final_bench_reference.log
