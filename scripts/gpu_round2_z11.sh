#!/bin/bash
mkdir -p gpurun_out
for w in n=64,ncutoff=5 n=768,ncutoff=7 n=2304,ncutoff=11 n=20000,ncutoff=12; do
  timeout 300 python scripts/profile_fullq.py $w 6 > gpurun_out/z11_fullq_$w.log 2>&1
  echo $w; tail -n 4 gpurun_out/z11_fullq_$w.log
done
