"""VirtualVolumeMove (src/analysis.cpp:825-843) restated in faunus_b200/csrc/host/analysis_virtual.hpp, driven by
the oracle's Hamiltonian: the energy change of the virtual scaling equals the difference of the full energies of two
independently built systems (the original and one whose cell and positions are scaled in Python)."""
import copy

import numpy as np
import pytest

from _oraclelib import oracle_sim


def test_virtual_volume_energy_change(bulk_input):
    dV = 25.0
    sim = oracle_sim(bulk_input)
    before = sim.system_energy()[0]
    vid = sim.virtualvolume_create({"dV": dV})
    sim.virtualvolume_sample(vid)
    res = sim.virtualvolume_result(vid)
    assert res["count"] == 1
    # the Space is scaled back (to rounding: the reference restores it the same way)
    assert sim.system_energy()[0] == pytest.approx(before, rel=1e-12)
    # independent: scale the input by hand (one atomic group: every position scales with the cell)
    scaled = copy.deepcopy(bulk_input)
    length = np.array(bulk_input["geometry"]["length"], dtype=float) * np.ones(3)
    s = ((length.prod() + dV) / length.prod()) ** (1.0 / 3.0)
    scaled["geometry"]["length"] = (length * s).tolist()
    for p in scaled["particles"]:
        p["pos"] = [x * s for x in p["pos"]]
    for g in scaled["groups"]:
        g["cm"] = [x * s for x in g["cm"]]
    after = oracle_sim(scaled).system_energy()[0]
    assert res["last_du"] == pytest.approx(after - before, rel=1e-9, abs=1e-9 * abs(before))
    assert res["sum_exp"] == pytest.approx(np.exp(-res["last_du"]), rel=1e-12)
    assert res["excess_pressure_kT_per_A3"] == pytest.approx(np.log(res["sum_exp"]) / dV, rel=1e-12)
    # a second sample accumulates; dV = 0 does nothing (src/analysis.cpp:827-829)
    sim.virtualvolume_sample(vid)
    assert sim.virtualvolume_result(vid)["count"] == 2
    idle = sim.virtualvolume_create({"dV": 0.0})
    sim.virtualvolume_sample(idle)
    assert sim.virtualvolume_result(idle)["count"] == 0
    with pytest.raises(RuntimeError):
        sim.virtualvolume_create({"dV": 1.0, "scaling": "isochoric"})
