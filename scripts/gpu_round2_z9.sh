#!/bin/bash
mkdir -p gpurun_out
for gy in 1 2 4 8; do
  FAUNUS_B200_FULL_GY=$gy timeout 600 python scripts/profile_fullpair.py s1 > gpurun_out/z9_fullpair_gy$gy.log 2>&1
  echo "gy=$gy"; tail -n 3 gpurun_out/z9_fullpair_gy$gy.log
done
