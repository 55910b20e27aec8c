#!/bin/bash
# CUDA-graph replay of the run launches: parity (all GPU tests), bench with and without graphs, light-fence tail variant
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s_pytest.log 2>&1
tail -4 gpurun_out/s_pytest.log
report() {
python - "$1" "$2" <<'PY'
import json, sys
line = [l for l in open(sys.argv[2]) if l.startswith("{")]
if not line:
    print(sys.argv[1], "no result", open(sys.argv[2]).read()[-600:])
else:
    j = json.loads(line[-1])
    print(sys.argv[1], "value %.0f e2e %.0f" % (j["value"], j["e2e"]["value"]), j.get("device_time_split_us_per_move"),
          "first-to-last us/move %.4f" % j["host_split_us_per_move"]["device_first_to_last_kernel"],
          "ks+front us/launch %.2f" % j["roofline"]["us_per_launch"], "frac %.3f" % j["roofline"]["frac"])
PY
}
for rep in 1 2; do
    python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/s_bench_graph_$rep.log 2>&1
    report graphs gpurun_out/s_bench_graph_$rep.log
    FAUNUS_B200_NO_GRAPHS=1 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/s_bench_plain_$rep.log 2>&1
    report plain gpurun_out/s_bench_plain_$rep.log
    FAUNUS_B200_LIB=$PWD/faunus_b200/_build/variants/v_fence/libfaunus_b200.so FAUNUS_B200_NO_GRAPHS=1 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/s_bench_fence_$rep.log 2>&1
    report fence-plain gpurun_out/s_bench_fence_$rep.log
done
