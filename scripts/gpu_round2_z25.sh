#!/bin/bash
# round 2, the end: per-nx column groups in the full-Q product
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "full_q or shards or ewald_doctest" > gpurun_out/z25_pytest_sel.log 2>&1
echo "pytest rc=$?" >> gpurun_out/z25_pytest_sel.log
timeout 300 python scripts/profile_fullq.py s1 4 > gpurun_out/z25_fullq_s1.log 2>&1
tail -n 3 gpurun_out/z25_pytest_sel.log; tail -n 4 gpurun_out/z25_fullq_s1.log
