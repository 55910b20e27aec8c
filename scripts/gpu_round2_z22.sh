#!/bin/bash
# round 2, the end: whole suite, smoke, the bench line as the driver runs it, the reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/z22_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/z22_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z22_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/z22_smoke.log
timeout 900 python bench.py > gpurun_out/z22_bench_s1.log 2>&1
echo "bench rc=$?" >> gpurun_out/z22_bench_s1.log
python - <<PY
import json
for l in open("gpurun_out/z22_bench_s1.log"):
    if l.startswith("{"):
        d = json.loads(l); print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", d["roofline"]["frac"], d["sharded_summary"], d["clocks"])
PY
tail -n 3 gpurun_out/z22_pytest.log gpurun_out/z22_smoke.log
