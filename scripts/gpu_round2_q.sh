#!/bin/bash
# k-space schedule experiment: parity of the windowed paths, bench without extras, warm launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
TAG=${1:-q}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "window or runs or s1 or walk or system_energy_and_moves" > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
( time python bench.py --steps 10 --warmup 3 --no-extras ) > gpurun_out/${TAG}_bench.log 2>&1
tail -c 2500 gpurun_out/${TAG}_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 100 -c 300 --csv --log-file gpurun_out/${TAG}_launches_warm.csv \
    python scripts/profile_moves.py 2000 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
