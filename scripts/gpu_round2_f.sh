#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${2:-6} -c ${3:-2} \
    -o gpurun_out/$4 python scripts/profile_moves.py 2000 > gpurun_out/f_ncu.log 2>&1
tail -3 gpurun_out/f_ncu.log
