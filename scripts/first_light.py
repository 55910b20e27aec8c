"""GPU first light: bulk example on the B200 path vs the oracle (energies, trace, timing)."""
import json, sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _oraclelib import oracle_sim
from faunus_b200.native import B200Simulation
from faunus_b200.config import primitive_model

cfg = json.load(open(os.path.join(ROOT, "tests/golden/bulk_input.json")))
o = oracle_sim(cfg); g = B200Simulation(cfg)
eo, to = o.system_energy(); eg, tg = g.system_energy()
print("system energy oracle", eo, to); print("system energy b200  ", eg, tg, "rel", abs(eg-eo)/abs(eo))
for s in (o, g): s.trace_enable()
t=time.time(); o.sweep(1); dto=time.time()-t
t=time.time(); g.sweep(1); dtg=time.time()-t
a, b = o.trace(), g.trace()
n=len(a["du"]); print("moves", n, "oracle s", dto, "b200 s", dtg, "b200 moves/s", n/dtg)
same = (a["accepted"]==b["accepted"]).all(); print("accept identical:", same, "acc", a["accepted"].mean())
fin = np.isfinite(a["du"]) & np.isfinite(b["du"])
print("max |du diff|", np.abs(a["du"][fin]-b["du"][fin]).max(), "max rel u_new", (np.abs(a["u_new"][fin]-b["u_new"][fin])/np.abs(a["u_new"][fin])).max())
print("drift", o.drift(), g.drift(), "launches", g.launch_count)
# synthetic N=20000 ewald
cfg2 = primitive_model(n=20000, coulomb={"type":"ewald","epsr":78.7,"cutoff":14.0,"alpha":0.22,"ncutoff":12})
t=time.time(); o2 = oracle_sim(cfg2); print("oracle create", time.time()-t)
t=time.time(); g2 = B200Simulation(cfg2); print("b200 create", time.time()-t)
e1,t1=o2.system_energy(); e2,t2=g2.system_energy(); print(t1, t2, "rel", np.abs(t1-t2)/np.abs(t1))
for s in (o2,g2): s.trace_enable()
t=time.time(); o2.sweep(200); dto=time.time()-t
t=time.time(); g2.sweep(200); dtg=time.time()-t
a,b=o2.trace(),g2.trace(); print("N=2e4 ewald: oracle moves/s", 200/dto, "b200 moves/s", 200/dtg, "accept same", (a["accepted"]==b["accepted"]).all(), "max du diff", np.abs(a["du"]-b["du"]).max())
