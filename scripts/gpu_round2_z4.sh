#!/bin/bash
# round 2, late: three-multiplication complex product
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "full_q_matrix or shards or ewald_doctest" > gpurun_out/z4_pytest_sel.log 2>&1
echo "pytest rc=$?" >> gpurun_out/z4_pytest_sel.log
timeout 600 python scripts/profile_fullq.py s1 4 > gpurun_out/z4_fullq_s1.log 2>&1
tail -n 5 gpurun_out/z4_pytest_sel.log gpurun_out/z4_fullq_s1.log
