#!/bin/bash
# round 2, late: unit phases computed once, windows from the first nz; blocks per SM of the particle ranges
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "full_q_matrix or shards or ewald_doctest" > gpurun_out/z3_pytest_sel.log 2>&1
echo "pytest rc=$?" >> gpurun_out/z3_pytest_sel.log
for b in 6 12 24; do
  FAUNUS_B200_FULLQ_BLOCKS=$b timeout 600 python scripts/profile_fullq.py s1 4 > gpurun_out/z3_fullq_s1_b$b.log 2>&1
done
tail -n 5 gpurun_out/z3_pytest_sel.log gpurun_out/z3_fullq_s1_b*.log
