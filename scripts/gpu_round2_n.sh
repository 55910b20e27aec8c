#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/n_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/n_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n_synccheck.log 2>&1
echo "synccheck rc=$?" >> gpurun_out/n_synccheck.log
tail -4 gpurun_out/n_memcheck.log; tail -4 gpurun_out/n_racecheck.log; tail -4 gpurun_out/n_synccheck.log
