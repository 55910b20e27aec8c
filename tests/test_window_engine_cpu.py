"""Host logic of the windowed Metropolis engine on the CPU (faunus_b200/csrc/host/montecarlo.hpp: drawing ahead in
generator order, conditional proposals, evaluations queued behind each other, replay of the decisions): the
engine runs against a CPU stand-in for the device (oracle/window_shadow.hpp — a second simulation that carries every
shipped proposal out one at a time by the reference protocol and checks the shipped start positions against its
own state) and must reproduce the plain one-move-at-a-time run bit for bit."""
import ctypes as C
import json

import numpy as np
import pytest

from _oraclelib import oracle_sim
from conftest import small_electrolyte


def shadowed(cfg, moves):
    sim = oracle_sim(cfg)
    fn = sim.api.lib.fo_sim_set_shadow_window
    fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_char_p, C.c_int]
    assert fn(sim.handle, json.dumps(cfg).encode(), moves) == 0, sim.api.error()
    return sim


def shadow_stats(sim):
    fn = sim.api.lib.fo_sim_shadow_stats
    fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.POINTER(C.c_double)]
    out = (C.c_double * 3)()
    assert fn(sim.handle, out) == 0
    return {"conditional": int(out[0]), "queued": int(out[1]), "evaluations": int(out[2])}


CONFIGS = {
    # 40 ions, 300 moves per sweep: almost every proposal meets an undecided move on its atom
    "crowded_cutoff": lambda: small_electrolyte(n=40, moves_per_sweep=300,
                                                coulomb={"type": "fanourgakis", "epsr": 78.7, "cutoff": 10.0}),
    "ewald": lambda: small_electrolyte(n=150, moves_per_sweep=400,
                                       coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 9.0, "alpha": 0.35, "ncutoff": 5}),
    "hard_spheres": lambda: small_electrolyte(n=120, moves_per_sweep=300, energy_name="nonbonded_pm",
                                              coulomb={"epsr": 78.7}, sigma=3.0),
}


@pytest.mark.parametrize("moves", [1, 7, 64, 512])
@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_windowed_engine_reproduces_the_plain_run(name, moves):
    cfg = CONFIGS[name]()
    plain, windowed = oracle_sim(cfg), shadowed(cfg, moves)
    for s in (plain, windowed):
        s.trace_enable()
        s.sweep(3)
    a, b = plain.trace(), windowed.trace()
    assert len(a["du"]) == len(b["du"]) > 800
    assert np.array_equal(a["move_id"], b["move_id"])
    assert np.array_equal(a["accepted"], b["accepted"])
    for key in ("u_new", "u_old", "du"):   # the same arithmetic on the same numbers: equal, not close
        assert np.array_equal(a[key], b[key]), key
    xa, ia = plain.particles()
    xb, ib = windowed.particles()
    assert np.array_equal(xa, xb) and np.array_equal(ia, ib)
    xt, _ = windowed.particles(1)   # the trial Space is back in step with the accepted one
    assert np.array_equal(xb, xt)
    assert plain.sum_energy_changes == windowed.sum_energy_changes
    stats = shadow_stats(windowed)
    assert stats["evaluations"] > 0
    if moves >= 64:
        assert stats["conditional"] > 0, "no conditional proposal was shipped"
        if name == "ewald" and moves == 64:   # elsewhere whatever follows the ready proposals is usually blocked
            assert stats["queued"] > 0, "no evaluation was queued behind another one"
    info_a, info_b = plain.info(), windowed.info()
    assert [m for m in info_a["moves"]] == [m for m in info_b["moves"]]   # attempts, acceptance, msd per move


@pytest.mark.parametrize("moves", [5, 64])
def test_windowed_engine_with_rigid_molecules(water_input, moves):
    """windows of rigid-molecule moves and of single-atom moves take turns (a window is of one kind; the queue drains
    when the move sampler changes kind): water + NaCl, cutoff electrostatics, against the plain run"""
    from conftest import water_with_salt
    cfg = water_with_salt(water_input, coulomb={"type": "fanourgakis", "epsr": 1, "cutoff": 9})
    plain, windowed = oracle_sim(cfg), shadowed(cfg, moves)
    for s in (plain, windowed):
        s.trace_enable()
        s.sweep(2)
    a, b = plain.trace(), windowed.trace()
    assert len(a["du"]) > 500 and set(a["move_id"]) == {0, 1}
    assert np.array_equal(a["move_id"], b["move_id"])
    assert np.array_equal(a["accepted"], b["accepted"])
    for key in ("u_new", "u_old", "du"):
        assert np.array_equal(a[key], b[key]), key
    xa, _ = plain.particles()
    xb, _ = windowed.particles()
    assert np.array_equal(xa, xb)
    assert np.array_equal(plain.groups()[1], windowed.groups()[1])   # mass centres
    assert shadow_stats(windowed)["evaluations"] > 10
