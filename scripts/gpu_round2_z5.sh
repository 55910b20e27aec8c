#!/bin/bash
# round 2, late: ncu capture + launch list of the full-Q product, sanitizer over its tests, bench line
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ewaldFullG -c 2 -f -o gpurun_out/r02r_fullq python scripts/profile_fullq.py s1 1 > gpurun_out/z5_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ewald -c 40 --csv --log-file gpurun_out/r02r_fullq_launches.csv python scripts/profile_fullq.py s1 3 > gpurun_out/z5_ncu_list.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full_q_matrix or shards" > gpurun_out/z5_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/z5_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full_q_matrix" > gpurun_out/z5_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/z5_racecheck.log
timeout 900 python bench.py > gpurun_out/z5_bench_s1.log 2>&1
tail -n 6 gpurun_out/z5_memcheck.log gpurun_out/z5_racecheck.log
tail -c 700 gpurun_out/z5_bench_s1.log
