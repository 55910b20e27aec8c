import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import bench
import faunus_b200.native as native
sim = native.B200Simulation(bench.workload(moves_per_step=10, which="s1"))
rdf = sim.rdf_create({"name1": "Na", "name2": "Cl", "dr": 0.1, "file": "rdf.dat"})
sim.enable_timing(True)
for rep in range(4):
    t0 = time.perf_counter(); sim.rdf_sample_shard(rdf, 0, 1); t1 = time.perf_counter()
    r = sim.rdf_result(rdf); t2 = time.perf_counter()
    print("sample %.2f ms  result %.2f ms  bins %d  device %s" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, len(r[1]), sim.device_time_ms()))
for rep in range(3):
    t0 = time.perf_counter(); sim.rdf_sample_shard(rdf, 0, 4); t1 = time.perf_counter()
    print("quarter sample %.2f ms" % ((t1 - t0) * 1e3))
