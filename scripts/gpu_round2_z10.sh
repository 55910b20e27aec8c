#!/bin/bash
# round 2, end: whole GPU suite, smoke, bench line, reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/z10_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/z10_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z10_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/z10_smoke.log
timeout 900 python bench.py > gpurun_out/z10_bench_s1.log 2>&1
echo "bench rc=$?" >> gpurun_out/z10_bench_s1.log
tail -n 4 gpurun_out/z10_pytest.log gpurun_out/z10_smoke.log
tail -c 600 gpurun_out/z10_bench_s1.log
