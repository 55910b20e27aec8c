"""Short profiling driver: one Widom sample event (32768 ion-pair insertions) + one full energy at N=1e5 (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import faunus_b200.native as native
from faunus_b200.config import primitive_model
cfg = primitive_model(n=100_000, molarity=1.0, seed=5489, moves_per_sweep=10, ghost_pairs=1,
                      coulomb={"type": "fanourgakis", "epsr": 78.7, "cutoff": 28.0})
sim = native.B200Simulation(cfg)
w = sim.widom_create({"molecule": "ghost", "ninsert": 32768})
sim.widom_sample(w, 2)
print("done", sim.system_energy()[0], sim.launch_count)
