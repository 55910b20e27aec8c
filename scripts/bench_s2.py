"""S2 "pm-1e6" (SURVEY §8d): RPM electrolyte N = 1 000 000, 1.0 M, L = 940 Å, Ewald alpha 0.12 / Rc 28 Å /
ncutoff 65 (K ≈ 5.8e5), windowed single-ion moves with the device cell list. Prints one JSON line.
No oracle at this size: the check is the reference's own drift invariant (src/montecarlo.cpp:85-99)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import faunus_b200.native as native
from faunus_b200.config import primitive_model

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
moves = 2000
ncut = 30.0 * (n / 1e5) ** (1.0 / 3.0)
t0 = time.perf_counter()
cfg = primitive_model(n=n, molarity=1.0, seed=5489, moves_per_sweep=moves,
                      coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 28.0, "alpha": 0.12, "ncutoff": ncut,
                               "ewaldscheme": "PBC"})
t1 = time.perf_counter()
sim = native.B200Simulation(cfg, window=64)
t2 = time.perf_counter()
info = sim.info()
K = [t["ewald"]["wavefunctions"] for t in info["energy"] if "ewald" in t][0]
sim.sweep(1)
w0 = sim.window_time_ms()
ta = time.perf_counter()
sim.sweep(3)
tb = time.perf_counter()
w1 = sim.window_time_ms()
sim.enable_timing(True)
s0 = sim.window_time_ms()
sim.sweep(2)
s1 = sim.window_time_ms()
sim.enable_timing(False)
t3 = time.perf_counter()
drift = sim.drift()
t4 = time.perf_counter()
print(json.dumps({
    "workload": f"pm-{n:.0e}: N={n}, K={K}, window 64, device cell list", "moves_per_s_e2e": 3 * moves / (tb - ta),
    "moves_per_s_device": 3 * moves / ((w1["total_ms"] - w0["total_ms"]) / 1e3),
    "us_per_move_split": {k: 1e3 * (s1[k] - s0[k]) / (2 * moves) for k in ("pair_ms", "ewald_ms", "other_ms")},
    "relative_drift": drift, "seconds": {"config": t1 - t0, "init": t2 - t1, "full_energy_drift_check": t4 - t3},
}))
