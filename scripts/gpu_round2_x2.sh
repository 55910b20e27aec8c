#!/bin/bash
# memcheck over the tests of the code added late in the round (forces, sharded energy splits, graphs with a changing box, restore)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
    -k "forces or shards or volume_moves or restore or matter_change or packed_state or particle_and_group" > gpurun_out/x2_memcheck_tests.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/x2_memcheck_tests.log
tail -6 gpurun_out/x2_memcheck_tests.log
