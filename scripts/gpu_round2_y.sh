#!/bin/bash
# fb_configure_runs flags side by side: default, pair sums ahead (1), no graphs (2)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for flags in 0 1 2; do
  for rep in 1 2; do
    FAUNUS_B200_RUN_FLAGS=$flags python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/y_bench_${flags}_${rep}.log 2>&1
    python - $flags gpurun_out/y_bench_${flags}_${rep}.log <<'PY'
import json, sys
line = [l for l in open(sys.argv[2]) if l.startswith("{")]
if not line:
    print(sys.argv[1], "no result", open(sys.argv[2]).read()[-400:])
else:
    j = json.loads(line[-1])
    print("flags", sys.argv[1], "value %.0f e2e %.0f" % (j["value"], j["e2e"]["value"]))
PY
  done
done
