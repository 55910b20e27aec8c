"""Per-call latency of fb_trial_energy (fused fast path) through the raw C ABI."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import faunus_b200.native as native
from faunus_b200.config import primitive_model

lib = native.load()
def probe(n, coulomb, label, reps=3000):
    cfg = primitive_model(n=n, coulomb=coulomb, moves_per_sweep=10)
    sim = native.B200Simulation(cfg)
    sim.sweep(1)  # establishes old groups etc.
    ctx = sim.ctx
    xyzq, ids = sim.particles()
    mv = native.FbTrialMove()
    mv.group_index = 0; mv.n_atoms = 1; mv.internal = 1
    un, uo, en, eo = (C.c_double() for _ in range(4))
    for with_ewald in ((0, 1) if coulomb["type"] == "ewald" else (0,)):
        mv.with_ewald = with_ewald
        rng = np.random.RandomState(1)
        idx = rng.randint(0, n, reps)
        t0 = time.perf_counter()
        for i in idx:
            mv.rel_index[0] = int(i); mv.atom_id[0] = int(ids[i])
            p = xyzq[i]
            mv.xyzq[0][0] = p[0] + 0.5; mv.xyzq[0][1] = p[1]; mv.xyzq[0][2] = p[2]; mv.xyzq[0][3] = p[3]
            rc = lib.fb_trial_energy(ctx, C.byref(mv), C.byref(un), C.byref(uo), C.byref(en), C.byref(eo))
            assert rc == 0, lib.fb_last_error(ctx)
            lib.fb_trial_commit(ctx, 0)
        dt = time.perf_counter() - t0
        print(f"{label:28s} N={n:7d} ewald={with_ewald} {1e6*dt/reps:7.2f} us/call")
    sim.close()

ew = {"type": "ewald", "epsr": 78.7, "cutoff": 28.0, "alpha": 0.12, "ncutoff": 30}
probe(100000, ew, "S1 K=56k")
probe(100000, {"type": "fanourgakis", "epsr": 78.7, "cutoff": 14.0}, "no k-space")
probe(2000, {"type": "fanourgakis", "epsr": 78.7, "cutoff": 14.0}, "tiny")
probe(20000, dict(ew, ncutoff=8, cutoff=14.0, alpha=0.22), "N=2e4 small K")
