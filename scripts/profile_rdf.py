"""Short profiling driver: one atomrdf sample (Na-Cl) on the S1 configuration (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import faunus_b200.native as native
sim = native.B200Simulation(bench.workload(moves_per_step=10))
rid = sim.rdf_create({"name1": "Na", "name2": "Cl", "dr": 0.1, "file": "rdf.dat"})
for _ in range(2):
    sim.rdf_sample(rid)
print("pairs", int(sim.rdf_result(rid)[1].sum()))
