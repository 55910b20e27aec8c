#!/bin/bash
# round 2, the very end: whole GPU suite and smoke with the per-nx column groups
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/z26_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/z26_pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z26_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/z26_smoke.log
tail -n 3 gpurun_out/z26_pytest.log gpurun_out/z26_smoke.log
