#!/bin/bash
# first GPU call of round 2: whole -m gpu suite, sanitizer passes over smoke(), a baseline bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nproc > gpurun_out/a_nproc.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/a_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/a_racecheck.log
( time python bench.py --steps 10 --warmup 3 --no-extras ) > gpurun_out/a_bench.log 2>&1
tail -c 600 gpurun_out/a_pytest.log; tail -5 gpurun_out/a_memcheck.log; tail -5 gpurun_out/a_racecheck.log; tail -c 1500 gpurun_out/a_bench.log
