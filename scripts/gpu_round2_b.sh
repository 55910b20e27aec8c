#!/bin/bash
# new window k-space kernel: parity suite, then bench A/B against the old kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_pytest.log
python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/b_bench_new.log 2>&1
FAUNUS_B200_OLD_KSPACE=1 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/b_bench_old.log 2>&1
tail -c 1500 gpurun_out/b_pytest.log
