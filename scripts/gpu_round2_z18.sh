#!/bin/bash
# round 2, end: run input / decisions on copy streams beside the kernels
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/z18_bench_$i.log 2>&1
  python - <<PY
import json
for l in open("gpurun_out/z18_bench_$i.log"):
    if l.startswith("{"):
        d = json.loads(l); print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 2), d["host_split_us_per_move"])
PY
done
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/z18_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/z18_pytest.log
tail -n 4 gpurun_out/z18_pytest.log
