#!/bin/bash
# round 2, late: the matrix-product rebuild of Q(k)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/z_smi.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full_q_matrix or ewald_doctest" > gpurun_out/z_pytest_fullq.log 2>&1
echo "fullq pytest rc=$?" >> gpurun_out/z_pytest_fullq.log
timeout 600 python scripts/profile_fullq.py s1 5 > gpurun_out/z_fullq_s1.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/z_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/z_pytest.log
tail -5 gpurun_out/z_pytest_fullq.log gpurun_out/z_fullq_s1.log gpurun_out/z_pytest.log
