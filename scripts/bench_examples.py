"""Throughput of the bundled-example configurations (parity cases, not the bench line): oracle (CPU port, 1 thread)
vs the B200 path, same inputs and seeds."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _oraclelib import oracle_sim
from faunus_b200.native import B200Simulation

def rate(sim, sweeps):
    sim.sweep(1)
    sim.trace_enable()
    t = time.perf_counter(); sim.sweep(sweeps); dt = time.perf_counter() - t
    return len(sim.trace()["du"]) / dt

out = {}
for name, sweeps in (("bulk", 10), ("water", 10), ("minimal", 200)):
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", f"{name}_input.json")))
    o = oracle_sim(cfg); g = B200Simulation(cfg); g1 = B200Simulation(cfg, window=0)
    out[name] = {"oracle_moves_per_s": rate(o, sweeps), "b200_moves_per_s": rate(g, sweeps), "window": g.window,
                 "b200_one_move_per_launch": rate(g1, sweeps), "particles": g.num_particles}
print(json.dumps(out))
