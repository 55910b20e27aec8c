#!/bin/bash
# round 2, late: FP32-screened full pair sum
mkdir -p gpurun_out
timeout 600 python scripts/profile_fullpair.py s1 > gpurun_out/z8_fullpair_s1.log 2>&1
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/z8_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/z8_pytest.log
tail -n 4 gpurun_out/z8_fullpair_s1.log gpurun_out/z8_pytest.log
