import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import faunus_b200.native as native
from faunus_b200.config import primitive_model
from _oraclelib import oracle_sim
for n in (2000, 20000):
    cfg = primitive_model(n=n, molarity=1.0, seed=5489, moves_per_sweep=10, ghost_pairs=1,
                          coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 14.0, "alpha": 0.22, "ncutoff": 8, "ewaldscheme": "PBC"})
    g = native.B200Simulation(cfg)
    w = g.widom_create({"molecule": "ghost", "ninsert": 4096})
    g.widom_sample(w, 1)
    r = g.widom_result(w)
    du = r["last_du"]
    print(n, "count", r["count"], "sum_exp", r["sum_exp"], "finite", np.isfinite(du).mean(), "min", np.nanmin(du), "median", np.nanmedian(du), "neg frac", (du < 0).mean())
    if n == 2000:
        o = oracle_sim(cfg); wo = o.widom_create({"molecule": "ghost", "ninsert": 4096}); o.widom_sample(wo, 1)
        ro = o.widom_result(wo); print("oracle", ro["sum_exp"], np.nanmin(ro["last_du"]), np.abs(ro["last_du"] - du).max())
