"""Run statistics and host/device split of the device-decided runs on the S1 workload (and examples/bulk)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import faunus_b200.native as native
out = {}
for name, cfg, sweeps in (("s1_2000", bench.workload(moves_per_step=2000), 10), ("s1_20000", bench.workload(moves_per_step=20000), 2),
                          ("bulk", json.load(open(os.path.join(ROOT, "tests/golden/bulk_input.json"))), 10)):
    for run in (512, 0):
        sim = native.B200Simulation(cfg, run=run)
        sim.sweep(1)
        w0, r0 = sim.window_time_ms(), sim.run_stats()
        t = time.perf_counter(); sim.sweep(sweeps); dt = time.perf_counter() - t
        w1, r1 = sim.window_time_ms(), sim.run_stats()
        d = {k: w1[k] - w0[k] for k in w1}; r = {k: r1[k] - r0[k] for k in r1}
        out[f"{name}_run{run}"] = {"moves_per_s": d["moves"] / dt, "device_us_per_move": 1e3 * d["total_ms"] / max(d["moves"], 1),
                                   "round_trips": d["round_trips"], "windows": d["windows"], "moves": d["moves"], **{"run_" + k: v for k, v in r.items()},
                                   "host_evaluate_us_per_move": 1e3 * d["host_evaluate_ms"] / max(d["moves"], 1),
                                   "host_sweep_us_per_move": 1e3 * d["host_sweep_ms"] / max(d["moves"], 1)}
print(json.dumps(out, indent=1))
