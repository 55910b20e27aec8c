#!/bin/bash
# sanitizer over run-heavy tests (graph replay, device walk, conditional proposals, forces)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
    -k "device_walk or runs_through or forces_match_oracle or test_bulk_example_trace or shards_with_particle" > gpurun_out/x_memcheck_tests.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/x_memcheck_tests.log
tail -6 gpurun_out/x_memcheck_tests.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
    -k "device_walk or forces_of_the_bulk" > gpurun_out/x_racecheck_tests.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/x_racecheck_tests.log
tail -6 gpurun_out/x_racecheck_tests.log
