#!/bin/bash
# round 2, end: ncu capture of the screened full pair sum (and of the all-FP64 kernel it replaces)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fullScreenKernel|fullStreamKernel" -c 2 -f -o gpurun_out/r02s_fullpair python scripts/profile_fullpair.py s1 > gpurun_out/z15_ncu.log 2>&1
tail -n 3 gpurun_out/z15_ncu.log
