#!/bin/bash
# round 2, end: IPBC through the product; whole suite
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/z17_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/z17_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z17_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/z17_smoke.log
tail -n 4 gpurun_out/z17_pytest.log gpurun_out/z17_smoke.log
