"""Full rebuild of Q(k) (fb_ewald_update_full) at S1 (N = 1e5, K = 57 950): the matrix-product kernel (fb_fullq.cuh) against
the one-block-per-cell kernel (FAUNUS_B200_FULLQ=cells), device time per rebuild (CUDA events on the context's stream) and
the largest difference between the two Q(k)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import faunus_b200.native as native
import bench

which = sys.argv[1] if len(sys.argv) > 1 else "s1"
repeats = int(sys.argv[2]) if len(sys.argv) > 2 else 5
if which.startswith("n="):  # e.g. n=2304,ncutoff=11,alpha=0.3: a small electrolyte (the crossover with the cell kernel)
    from faunus_b200.config import primitive_model
    kv = dict(item.split("=") for item in which.split(","))
    cfg = primitive_model(n=int(kv["n"]), molarity=1.0, seed=7, placement="lattice",
                          coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 9.0, "alpha": float(kv.get("alpha", 0.35)),
                                   "ncutoff": float(kv.get("ncutoff", 7))})
else:
    cfg = bench.workload(which=which)
lib = native.load()
out = {}
for path in ("gemm", "cells"):
    os.environ["FAUNUS_B200_FULLQ"] = path
    sim = native.B200Simulation(cfg)
    ctx = sim.ctx
    lib.fb_enable_timing(ctx, 1)
    ms = []
    for _ in range(repeats):
        assert lib.fb_ewald_update_full(ctx, 0) == 0
        ms.append(lib.fb_last_kernel_ms(ctx))
    kmax = 4_000_000
    q = np.zeros(2 * kmax)
    assert lib.fb_ewald_download(ctx, 0, q.ctypes.data_as(native.c_double_p), None, None) == 0
    out[path] = (ms, q)
    print(path, "ms per rebuild:", " ".join(f"{m:.3f}" for m in ms), flush=True)
    del sim
qa, qb = out["gemm"][1], out["cells"][1]
print("max |Q_gemm - Q_cells| =", np.abs(qa - qb).max(), "max |Q| =", np.abs(qb).max())
n = len(cfg["particles"]); K = int(np.flatnonzero(qb).max() // 2 + 1)
best = min(out["gemm"][0][1:])
print(f"N = {n}, K = {K}: {8.0 * n * K / best / 1e9:.2f} TFLOP/s algorithmic (8 flop per particle and k-vector), "
      f"speed-up {min(out['cells'][0][1:]) / best:.2f}x")
