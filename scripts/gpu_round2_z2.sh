#!/bin/bash
# round 2, late: slabs of the matrix-product rebuild, ncu capture of the kernel, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "full_q_matrix or shards or ewald_doctest or tempering or volume" > gpurun_out/z2_pytest_sel.log 2>&1
echo "pytest rc=$?" >> gpurun_out/z2_pytest_sel.log
timeout 600 python scripts/profile_fullq.py s1 4 > gpurun_out/z2_fullq_s1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ewaldFullGemm -c 1 -f -o gpurun_out/r02q_fullq python scripts/profile_fullq.py s1 1 > gpurun_out/z2_ncu.log 2>&1
timeout 900 python bench.py > gpurun_out/z2_bench_s1.log 2>&1
tail -n 5 gpurun_out/z2_pytest_sel.log gpurun_out/z2_fullq_s1.log
tail -c 1600 gpurun_out/z2_bench_s1.log
