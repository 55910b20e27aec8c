#!/bin/bash
mkdir -p gpurun_out
for m in 20000 100000; do
  for i in 1 2; do
    FAUNUS_B200_MOVES_PER_STEP=$m timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/z14_bench_m${m}_$i.log 2>&1
    python - <<PY
import json
for l in open("gpurun_out/z14_bench_m${m}_$i.log"):
    if l.startswith("{"):
        d = json.loads(l); print("moves/step", $m, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 2), d["host_split_us_per_move"])
PY
  done
done
