"""Short profiling driver for the force kernels: examples/bulk-sized electrolyte with Ewald, forces of every term."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import faunus_b200.native as native
from faunus_b200.config import primitive_model
cfg = primitive_model(n=20000, coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 14.0, "alpha": 0.22, "ncutoff": 12})
sim = native.B200Simulation(cfg)
for _ in range(2):
    f = sim.forces(term=1), sim.forces(term=2)
print("done", abs(f[0]).max(), abs(f[1]).max())
