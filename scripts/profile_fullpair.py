"""Pair part of the full-system energy at S1 (N = 1e5, 5e9 pairs): the FP32-screened kernel (fullScreenKernel) against the
all-FP64 one (FAUNUS_B200_FULLPAIR=fp64), device time per evaluation (CUDA events) and the two energies."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import faunus_b200.native as native
import bench

which = sys.argv[1] if len(sys.argv) > 1 else "s1"
cfg = bench.workload(which=which)
n = len(cfg["particles"])
res = {}
for path in ("screen", "fp64"):
    os.environ["FAUNUS_B200_FULLPAIR"] = path
    sim = native.B200Simulation(cfg)
    sim.enable_timing(True)
    ms = []
    for _ in range(5):
        t0 = sim.device_time_ms()
        nb, rec = sim.system_energy_shard(0, 1)
        t1 = sim.device_time_ms()
        ms.append(t1["full_ms"] - t0["full_ms"])
    res[path] = (min(ms[1:]), nb)
    print(path, "ms per pair sum:", " ".join(f"{m:.3f}" for m in ms), "energy", repr(nb), flush=True)
    sim.close()
a, b = res["screen"], res["fp64"]
print(f"N = {n}: {n * (n - 1) / 2 / a[0] / 1e9:.3f}e12 pairs/s screened, {n * (n - 1) / 2 / b[0] / 1e9:.3f}e12 all-FP64, "
      f"speed-up {b[0] / a[0]:.2f}x, relative energy difference {abs(a[1] - b[1]) / abs(b[1]):.2e}")
