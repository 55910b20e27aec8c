#!/bin/bash
# force kernels: parity tests + the whole GPU suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "forces" > gpurun_out/p_forces.log 2>&1
tail -15 gpurun_out/p_forces.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/p_pytest.log 2>&1
tail -5 gpurun_out/p_pytest.log
