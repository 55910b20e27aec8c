#!/bin/bash
# ncu --set full of the window k-space kernel (new), source page included
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:windowKspaceKernel -s 6 -c 2 \
    -o gpurun_out/r02b_kspace python scripts/profile_moves.py 2000 > gpurun_out/c_ncu.log 2>&1
tail -3 gpurun_out/c_ncu.log
ls -la gpurun_out/*.ncu-rep
