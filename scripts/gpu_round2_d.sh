#!/bin/bash
# front kernel + pipelined window k-space kernel: parity suite, bench, ncu capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/d_pytest.log
python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/d_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"windowKspaceKernel|windowFrontKernel" -s 12 -c 4 \
    -o gpurun_out/r02d_kspace python scripts/profile_moves.py 2000 > gpurun_out/d_ncu.log 2>&1
tail -c 600 gpurun_out/d_pytest.log; tail -3 gpurun_out/d_ncu.log
