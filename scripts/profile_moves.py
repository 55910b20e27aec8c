"""Short profiling driver: N=1e5 workload, a few hundred moves (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import faunus_b200.native as native
moves = int(sys.argv[1]) if len(sys.argv) > 1 else 100
which = sys.argv[2] if len(sys.argv) > 2 else "s1"  # s1 | s1-largeK | s2
if which == "s2":
    sys.path.insert(0, os.path.join(ROOT, "tests"))
sim = native.B200Simulation(bench.workload(moves_per_step=moves, which=which))
sim.sweep(2)
print("done", sim.launch_count)
