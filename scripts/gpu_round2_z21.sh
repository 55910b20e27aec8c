#!/bin/bash
mkdir -p gpurun_out
for r in 512 1024; do
FAUNUS_B200_GAP_STATS=1 FAUNUS_B200_RUN=$r timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/z21_bench_gap_$r.log 2>&1
grep "chained runs" gpurun_out/z21_bench_gap_$r.log
python - <<PY
import json
for l in open("gpurun_out/z21_bench_gap_$r.log"):
    if l.startswith("{"):
        d = json.loads(l); print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 2))
PY
done
