#!/bin/bash
# round 2, end: the two other workloads of SURVEY 8(d) with the final kernels
mkdir -p gpurun_out
timeout 900 python bench.py --workload s1-largeK --no-extras --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/z16_bench_largeK.log 2>&1
echo "rc=$?" >> gpurun_out/z16_bench_largeK.log
timeout 900 python bench.py --workload s2 --no-extras --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/z16_bench_s2.log 2>&1
echo "rc=$?" >> gpurun_out/z16_bench_s2.log
for f in largeK s2; do python - <<PY
import json
for l in open("gpurun_out/z16_bench_$f.log"):
    if l.startswith("{"):
        d = json.loads(l); r = d.get("roofline") or {}
        print("$f", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "kspace frac", r.get("frac"), "us", r.get("us_per_launch"))
PY
done
tail -n 2 gpurun_out/z16_bench_largeK.log | tail -c 300
