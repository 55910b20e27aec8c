"""Matter changes (first half) in the ORACLE: GroupPairing::accumulateSpeciation
(src/energy.h:1390-1435) against an independent evaluation — the energy of a matter change in a state is the
difference of two full energies, with and without the listed active particles — and the translational-entropy bias
(src/montecarlo.cpp:282-299) against its closed form."""
import copy

import numpy as np

from _oraclelib import oracle_sim
from conftest import small_electrolyte

COULOMB = {"type": "fanourgakis", "epsr": 78.7, "cutoff": 10.0}


def _cfg():
    return small_electrolyte(n=80, ghost_pairs=3, moves_per_sweep=40, coulomb=COULOMB)


POS = [[3.0, -7.0, 11.0], [-9.0, 2.5, -4.0], [12.0, 12.0, -12.0]]


def _ghost_group(sim):
    rec, _ = sim.groups()
    return int(np.argmax(rec[:, 2] - rec[:, 1]))  # the group with inactive capacity


def test_insertion_energy_is_a_difference_of_full_energies():
    sim = oracle_sim(_cfg())
    g = _ghost_group(sim)
    e0 = sim.system_energy()[0]
    r = sim.matter_change([{"index": g, "size": 2, "atoms": [0, 1], "pos": POS[:2], "dNatomic": True}], mode=1)
    e1 = sim.system_energy()[0]
    # (the self-energy term sums ALL active particles for a matter change, src/externalpotential.cpp:94-126, so
    # neither energy is zero by itself; the pair part of the old state is: nothing listed is active there)
    assert r["accepted"] and abs((e1 - e0) - (r["u_new"] - r["u_old"])) < 1e-9 * abs(e0)
    # removal of one of the two: the old state holds the pairs, the new state none
    r2 = sim.matter_change([{"index": g, "size": 1, "atoms": [1], "dNatomic": True}], mode=1)
    e2 = sim.system_energy()[0]
    assert abs((e2 - e1) - (r2["u_new"] - r2["u_old"])) < 1e-9 * abs(e0)
    # a rejected insertion leaves everything as it was
    r3 = sim.matter_change([{"index": g, "size": 3, "atoms": [1, 2], "pos": POS[1:], "dNatomic": True}], mode=0)
    assert not r3["accepted"] and sim.system_energy()[0] == e2
    sim.sweep(2)
    assert abs(sim.drift()) < 1e-9


def test_translational_entropy_bias():
    sim = oracle_sim(_cfg())
    g = _ghost_group(sim)
    info = sim.state_json()
    L = info["geometry"]["length"]
    volume = float(np.prod(L)) if isinstance(L, list) else float(L) ** 3
    molar = 6.022137e23 / 1e27
    r = sim.matter_change([{"index": g, "size": 2, "atoms": [0, 1], "pos": POS[:2], "dNatomic": True}], mode=1)
    assert r["bias"] == np.log(1 / (volume * molar)) + np.log(2 / (volume * molar))
    r = sim.matter_change([{"index": g, "size": 1, "atoms": [1], "dNatomic": True}], mode=1)
    assert r["bias"] == -np.log(2 / (volume * molar))
