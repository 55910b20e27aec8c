#!/bin/bash
# profiles of the round: ncu --set full of the window kernels, cold and warm launch lists
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"windowKspaceKernel|windowFrontKernel|batchPairScreenKernel|windowTailKernel" -s 16 -c 8 \
    -o gpurun_out/r02h_window python scripts/profile_moves.py 2000 > gpurun_out/m_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/r02_launches_cold.csv \
    python scripts/profile_moves.py 2000 >> gpurun_out/m_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 100 -c 300 --csv --log-file gpurun_out/r02_launches_warm.csv \
    python scripts/profile_moves.py 2000 >> gpurun_out/m_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"widomScreenKernel" -c 2 -o gpurun_out/r02i_widom python scripts/profile_widom.py >> gpurun_out/m_ncu.log 2>&1
tail -3 gpurun_out/m_ncu.log
