#!/bin/bash
# launch list with warm caches (no flush between kernels): closer to what a kernel costs inside a window
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 100 -c 300 --csv --log-file gpurun_out/g_launches_warm.csv \
    python scripts/profile_moves.py 2000 > gpurun_out/g_ncu.log 2>&1
tail -2 gpurun_out/g_ncu.log
