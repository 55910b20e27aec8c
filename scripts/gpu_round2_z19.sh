#!/bin/bash
# round 2, end: no overflow copy node without the cell list; whole suite, smoke, full bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/z19_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/z19_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z19_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/z19_smoke.log
timeout 900 python bench.py > gpurun_out/z19_bench_s1.log 2>&1
echo "bench rc=$?" >> gpurun_out/z19_bench_s1.log
python - <<PY
import json
for l in open("gpurun_out/z19_bench_s1.log"):
    if l.startswith("{"):
        d = json.loads(l); print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 2), d["roofline"]["frac"], d["sharded_summary"], d["parity"])
PY
tail -n 3 gpurun_out/z19_pytest.log gpurun_out/z19_smoke.log
