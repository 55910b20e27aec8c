#!/bin/bash
# programmatic dependent launch variant: parity of the run paths under the variant, then bench side by side
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
export FAUNUS_B200_LIB=$PWD/faunus_b200/_build/variants/v_pdl/libfaunus_b200.so
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "window or runs or s1 or walk or system_energy_and_moves or bulk_example" > gpurun_out/pdl_pytest.log 2>&1
tail -3 gpurun_out/pdl_pytest.log
unset FAUNUS_B200_LIB
bash scripts/gpu_round2_r.sh v_pdl 2>&1 | tail -6
