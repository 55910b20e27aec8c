#!/bin/bash
# round 2, late: the full-Q product at the two other workloads of SURVEY 8(d)
mkdir -p gpurun_out
timeout 900 python scripts/profile_fullq.py s1-largeK 3 > gpurun_out/z6_fullq_largeK.log 2>&1
timeout 900 python scripts/profile_fullq.py s2 3 > gpurun_out/z6_fullq_s2.log 2>&1
tail -n 5 gpurun_out/z6_fullq_largeK.log gpurun_out/z6_fullq_s2.log
