#!/bin/bash
# kernel variants side by side (faunus_b200.build --variant …): bench without extras, value / e2e / per-kernel split
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "window or runs or s1 or walk or system_energy_and_moves" > gpurun_out/r_pytest.log 2>&1
tail -3 gpurun_out/r_pytest.log
for v in default "$@"; do
    if [ "$v" = default ]; then unset FAUNUS_B200_LIB; else export FAUNUS_B200_LIB=$PWD/faunus_b200/_build/variants/$v/libfaunus_b200.so; fi
    for rep in 1 2 3; do
        python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r_bench_${v}_${rep}.log 2>&1
        python - "$v" gpurun_out/r_bench_${v}_${rep}.log <<'PY'
import json, sys
line = [l for l in open(sys.argv[2]) if l.startswith("{")]
if not line:
    print(sys.argv[1], "no result", open(sys.argv[2]).read()[-400:])
else:
    j = json.loads(line[-1])
    print(sys.argv[1], "value %.0f e2e %.0f" % (j["value"], j["e2e"]["value"]), j.get("device_time_split_us_per_move"),
          "ks+front us/launch %.2f" % j["roofline"]["us_per_launch"], "frac %.3f" % j["roofline"]["frac"])
PY
    done
done
