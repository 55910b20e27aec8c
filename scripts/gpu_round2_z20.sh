#!/bin/bash
mkdir -p gpurun_out
for r in 512 1024; do
  for i in 1 2; do
    FAUNUS_B200_RUN=$r timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/z20_bench_run${r}_$i.log 2>&1
    python - <<PY
import json
for l in open("gpurun_out/z20_bench_run${r}_$i.log"):
    if l.startswith("{"):
        d = json.loads(l); print("run", $r, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["runs"]["moves_per_run"], d["runs"]["windows_per_run"])
PY
  done
done
