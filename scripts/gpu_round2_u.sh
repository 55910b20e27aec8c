#!/bin/bash
# final profiles of the round: smoke, ncu --set full of the window kernels, cold / warm launch lists, sanitizer passes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/u_smoke.log 2>&1; tail -2 gpurun_out/u_smoke.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"windowKspaceKernel|windowFrontKernel|batchPairScreenKernel|windowTailKernel" -s 16 -c 8 \
    -o gpurun_out/r02p_window python scripts/profile_moves.py 2000 > gpurun_out/u_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/r02p_launches_cold.csv \
    python scripts/profile_moves.py 2000 >> gpurun_out/u_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 100 -c 300 --csv --log-file gpurun_out/r02p_launches_warm.csv \
    python scripts/profile_moves.py 2000 >> gpurun_out/u_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"nonbondedForceKernel|ewaldForceKernel" -c 4 -o gpurun_out/r02l_forces python scripts/profile_forces.py >> gpurun_out/u_ncu.log 2>&1
tail -3 gpurun_out/u_ncu.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/u_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/u_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/u_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/u_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/u_synccheck.log 2>&1
echo "synccheck rc=$?" >> gpurun_out/u_synccheck.log
tail -4 gpurun_out/u_memcheck.log; tail -4 gpurun_out/u_racecheck.log; tail -4 gpurun_out/u_synccheck.log
