#!/bin/bash
# after the force / sharded-energy changes: the whole GPU suite, force kernel profile, bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/v_pytest.log 2>&1
tail -4 gpurun_out/v_pytest.log
timeout 600 ncu --set full --clock-control none -k regex:"nonbondedForceKernel|ewaldForceKernel|forceSumKernel" -c 6 -o gpurun_out/r02l_forces python scripts/profile_forces.py > gpurun_out/v_ncu.log 2>&1
tail -2 gpurun_out/v_ncu.log
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/v_bench_s1.log 2>&1
tail -c 1500 gpurun_out/v_bench_s1.log
