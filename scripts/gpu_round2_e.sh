#!/bin/bash
# quick loop: S1 parity tests only, bench, launch list with per-kernel durations
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "s1_matches_oracle or windowed_equals or device_walk" ) > gpurun_out/e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/e_pytest.log
python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/e_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 200 --csv --log-file gpurun_out/e_launches.csv \
    python scripts/profile_moves.py 2000 > gpurun_out/e_ncu.log 2>&1
tail -c 400 gpurun_out/e_pytest.log
