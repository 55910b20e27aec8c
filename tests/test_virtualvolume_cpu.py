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


def test_virtual_translate(water_input):
    """VirtualTranslate (src/analysis.cpp:2794-2860): ΔU of moving the one molecule by dL along `dir` equals the
    difference of the full energies of two independently built systems; the Space is put back; the molecule is
    picked with one draw of the global generator"""
    from conftest import one_water_in_salt
    cfg = one_water_in_salt(water_input, coulomb={"type": "fanourgakis", "epsr": 1, "cutoff": 9})
    sim = oracle_sim(cfg)
    before = sim.system_energy()[0]
    x0, _ = sim.particles()
    vid = sim.virtualtranslate_create({"molecule": "water", "dL": 0.3, "dir": [1, 1, 0]})
    sim.virtualtranslate_sample(vid)
    res = sim.virtualtranslate_result(vid)
    assert res["count"] == 1
    x1, _ = sim.particles()
    assert np.allclose(x0, x1, rtol=0, atol=1e-12)
    moved = copy.deepcopy(cfg)
    d = 0.3 * np.array([1.0, 1.0, 0.0]) / np.sqrt(2.0)
    box = np.array(cfg["geometry"]["length"], dtype=float)
    for p in moved["particles"][:3]:
        q = np.array(p["pos"]) + d
        p["pos"] = (q - box * np.round(q / box)).tolist()
    c = np.array(moved["groups"][0]["cm"]) + d
    moved["groups"][0]["cm"] = (c - box * np.round(c / box)).tolist()
    after = oracle_sim(moved).system_energy()[0]
    assert res["last_du"] == pytest.approx(after - before, rel=1e-9, abs=1e-9 * abs(before))
    assert res["mean_force_kT_per_A"] == pytest.approx(np.log(res["sum_exp"]) / 0.3, rel=1e-12)
    with pytest.raises(RuntimeError):
        sim.virtualtranslate_create({"molecule": "salt", "dL": 0.3})   # atomic molecules are not allowed
