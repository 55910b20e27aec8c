"""Forces: the oracle's restatement of `Nonbonded::force` (src/energy.h:1584-1597), the pair forces
(src/potentials.h:32-40, 173-184, 600-606) and `Ewald::force` (src/energy.cpp:596-629) against the reference's
known-answer values and against minus the numerical gradient of the oracle's own energies; the product's S'(q)
tables against the oracle's, bit for bit. CPU only."""
import ctypes as C
import json

import numpy as np
import pytest

import faunus_b200.native as native
from _oraclelib import c_double_p, oracle_lib, oracle_sim, pair_energy, pair_force
from conftest import nacl_pair_input
from test_host_tables import SCHEMES
from test_oracle_golden import _input, _kT_per_kJmol

NEUTRAL = [{"A": {"sigma": 2.0, "eps": 0.9}}, {"B": {"sigma": 8.0, "eps": 0.1}}]
NO_COULOMB = {"type": "plain", "epsr": 80}


def test_lennard_jones_force_doctest():
    """src/potentials.cpp:769-779: custom A–B pair (ε = 2 kJ/mol, σ = 8 Å) at r = 9 Å along x"""
    cfg = _input(NEUTRAL, [{"nonbonded_coulomblj": {"coulomb": NO_COULOMB, "lennardjones": {
        "mixing": "LB", "custom": [{"A B": {"eps": 2.0, "sigma": 8}}]}}}])
    f = pair_force(cfg, "nonbonded_coulomblj", 0, 1, [[-9.0, 0.0, 0.0]])[0]  # b → a = a − b
    assert f[0] == pytest.approx(0.0142838474, rel=1e-8) and f[1] == 0.0 and f[2] == 0.0


def test_coulomb_plain_force_doctest():
    """src/potentials.cpp:1612-1626 (charges +1 / −1, 7 Å apart along z, ε_r = 80: 0.1429734149 towards b) and
    :647-664 (|F| = 0.1425956964 at r = (lB, 0.2, −0.1))"""
    ions = [{"A": {"q": 1.0, "sigma": 1.0, "eps": 0.0}}, {"B": {"q": -1.0, "sigma": 1.0, "eps": 0.0}}]
    cfg = _input(ions, [{"nonbonded_coulomblj": {"coulomb": NO_COULOMB, "lennardjones": {"mixing": "LB"}}}])
    f = pair_force(cfg, "nonbonded_coulomblj", 0, 1, [[0.0, 0.0, -7.0]])[0]
    assert f[2] == pytest.approx(0.1429734149, rel=1e-9) and f[0] == 0.0 and f[1] == 0.0
    lB = 7.0056973292
    f = pair_force(cfg, "nonbonded_coulomblj", 0, 1, [[lB, 0.2, -0.1]])[0]
    assert np.linalg.norm(f) == pytest.approx(0.1425956964, rel=1e-8)


def test_wca_force_closed_form():
    """src/potentials.h:173-184: zero beyond 2^(1/6) σ, else 6·4ε (2x² − x)/r² · r with x = (σ/r)⁶"""
    cfg = _input(NEUTRAL, [{"nonbonded_coulombwca": {"coulomb": NO_COULOMB, "wca": {"mixing": "LB"}}}])
    sigma, eps4 = 5.0, 4 * np.sqrt(0.9 * 0.1) * _kT_per_kJmol()
    for r in (4.0, 5.0, 5.6, 5.62, 7.0):
        f = pair_force(cfg, "nonbonded_coulombwca", 0, 1, [[0.0, r, 0.0]])[0]
        x = (sigma / r) ** 6
        want = 0.0 if r * r > sigma ** 2 * 1.2599210498948732 else eps4 * 6 * (2 * x * x - x) / r ** 2 * r
        assert f[1] == pytest.approx(want, rel=1e-12) and f[0] == 0.0 and f[2] == 0.0


def test_no_force_for_plain_coulomb_and_functors():
    """`Coulomb`, `HardSphere`, `FunctorPotential`, `SplinedPotential` inherit PairPotential::force, which throws
    (src/potentials.cpp:246-251)"""
    ions = [{"A": {"q": 1.0, "sigma": 2.0, "eps": 0.1}}, {"B": {"q": -1.0, "sigma": 2.0, "eps": 0.1}}]
    for name, body in (("nonbonded_pm", {"coulomb": {"epsr": 80}, "hardsphere": {"mixing": "arithmetic"}}),
                       ("nonbonded_pmwca", {"coulomb": {"epsr": 80}, "wca": {"mixing": "LB"}}),
                       ("nonbonded", {"default": [{"lennardjones": {"mixing": "LB"}}]})):
        with pytest.raises(RuntimeError, match="Force computation not implemented"):
            pair_force(_input(ions, [{name: body}]), name, 0, 1, [[3.0, 0.0, 0.0]])


@pytest.mark.parametrize("scheme", SCHEMES, ids=lambda s: s["type"] + ("+k" if "debyelength" in s else ""))
def test_coulomb_force_is_minus_the_energy_gradient(scheme):
    """every CoulombGalore scheme: lB zz/r³ [S(1 + κr) − q S'] e^{−κr} · r against the central difference of the
    oracle's pair energy (checks the analytic S'(q) of each scheme; spline tolerance utol = 0.005/lB)"""
    ions = [{"A": {"q": 1.0, "sigma": 1.0, "eps": 0.0}}, {"B": {"q": -2.0, "sigma": 1.0, "eps": 0.0}}]
    cfg = _input(ions, [{"nonbonded_coulomblj": {"coulomb": scheme, "lennardjones": {"mixing": "LB"}}}])
    rc = scheme.get("cutoff", 40.0)
    direction = np.array([1.0, -2.0, 2.0]) / 3.0
    h = 1e-4
    plain = _input(ions, [{"nonbonded_coulomblj": {"coulomb": {"type": "plain", "epsr": scheme["epsr"]},
                                                   "lennardjones": {"mixing": "LB"}}}])
    lB_zz = abs(pair_energy(plain, "nonbonded_coulomblj", 0, 1, [1.0])[0])
    for r in np.linspace(0.08, 0.97, 12) * rc:
        f = pair_force(cfg, "nonbonded_coulomblj", 0, 1, [direction * r])[0]
        u = pair_energy(cfg, "nonbonded_coulomblj", 0, 1, [r - h, r + h])
        slope = (u[1] - u[0]) / (2 * h)
        radial = float(f @ direction)
        assert np.allclose(f, radial * direction, rtol=0, atol=1e-12 * max(1.0, abs(radial)))
        # the slope of the SPLINED energy is what the difference quotient sees: good to a few utol / knot spacing in
        # S', i.e. to a small fraction of the natural scale lB |zz| / r² of the force
        utol = 0.005 / (lB_zz / 2.0)  # src/potentials.cpp:1634
        assert radial == pytest.approx(-slope, rel=1e-3, abs=max(2e-3, 10 * utol) * lB_zz / r ** 2)
    if "cutoff" in scheme:
        assert np.all(pair_force(cfg, "nonbonded_coulomblj", 0, 1, [direction * rc * 1.0001])[0] == 0.0)


@pytest.mark.parametrize("T", [298.15, 1100.0])
@pytest.mark.parametrize("scheme", SCHEMES, ids=lambda s: s["type"] + ("+k" if "debyelength" in s else ""))
def test_force_tables_identical(scheme, T):
    """the product's S'(q) table (host/potential_tables.hpp → fb_set_force_table) equals the oracle's"""
    lib = native.load()

    def table(fn):
        fn.restype = C.c_int
        fn.argtypes = [C.c_char_p, C.c_double, c_double_p, c_double_p, C.c_int]
        kn, co = np.zeros(2048), np.zeros(6 * 2048)
        n = fn(json.dumps(scheme).encode(), T, kn.ctypes.data_as(c_double_p), co.ctypes.data_as(c_double_p), 2048)
        assert n >= 2
        return n, kn[:n].copy(), co[:6 * (n - 1)].copy()

    a, b = table(oracle_lib().fo_coulomb_force_table), table(lib.fbh_coulomb_force_table)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def _salt_input(epss):
    """40 ions in a 20 Å box, splined Ewald real space + WCA + reciprocal space with a surface term"""
    rng = np.random.default_rng(7)
    n = 40
    pos = (rng.random((n, 3)) - 0.5) * 20.0
    atoms = [{"Na": {"q": 1.0, "sigma": 2.0, "eps": 0.5, "dp": 1.0}}, {"Cl": {"q": -1.0, "sigma": 2.0, "eps": 0.5, "dp": 1.0}}]
    return {
        "temperature": 298.15, "random": {"seed": "fixed"},
        "geometry": {"type": "cuboid", "length": [20.0, 20.0, 20.0]},
        "atomlist": atoms,
        "moleculelist": [{"salt": {"atoms": ["Na", "Cl"] * (n // 2), "atomic": True}}],
        "groups": [{"id": 0, "size": n, "cm": [0, 0, 0], "atomic": True, "compressible": False}],
        "particles": [{"id": i % 2, "pos": pos[i].tolist(), "q": 1.0 - 2.0 * (i % 2)} for i in range(n)],
        "energy": [{"nonbonded_coulombwca": {
            "coulomb": {"type": "ewald", "epsr": 80.0, "cutoff": 9.0, "alpha": 0.3, "ncutoff": 6, "epss": epss},
            "wca": {"mixing": "LB"}}}],
        "moves": [{"transrot": {"molecule": "salt", "dp": 1.0, "dprot": 0, "repeat": 1}}],
    }


def test_hamiltonian_force_follows_the_reference():
    """Hamiltonian::force = the terms in turn on one vector (src/energy.cpp:1162-1166); Ewald::force ASSIGNS
    (src/energy.cpp:610), so what the non-bonded term added before it is lost — as in the reference"""
    sim = oracle_sim(_salt_input(1.0))
    _, terms = sim.system_energy()
    assert len(terms) == 3  # self, nonbonded, ewald
    total = sim.forces()
    ewald = sim.forces(term=2)
    nonbonded = sim.forces(term=1)
    assert np.array_equal(total, ewald)
    assert np.abs(nonbonded).max() > 0 and np.all(sim.forces(term=0) == 0.0)
    # Newton's third law for the pair forces (forces[i] += f; forces[j] −= f)
    assert np.abs(nonbonded.sum(axis=0)).max() <= 1e-12 * np.abs(nonbonded).max()


@pytest.mark.parametrize("epss", [1.0, 80.0])
def test_ewald_and_pair_forces_are_minus_the_energy_gradient(epss):
    """F_i = −∂U/∂r_i by central differences of the oracle's own term energies, moving one ion at a time
    (trial_set: updateState → energy); the surface term needs ε_s ≥ 1 to be in the energy at all"""
    cfg = _salt_input(epss)
    sim = oracle_sim(cfg)
    xyzq, _ = sim.particles()
    f_pair, f_ewald = sim.forces(term=1), sim.forces(term=2)
    h = 1e-5
    for i in (0, 7, 23):
        for axis in range(3):
            u = []
            for sign in (-1.0, 1.0, 0.0):  # the last one puts the ion back
                pos = xyzq[i, :3].copy()
                pos[axis] += sign * h
                sim.trial_set(0, [i], [pos])
                sim.trial_commit(True)
                u.append(sim.system_energy()[1])
            slope = (u[1] - u[0]) / (2 * h)
            assert f_pair[i, axis] == pytest.approx(-slope[1], rel=3e-2, abs=5e-3)  # spline of S: utol = 0.005/lB
            assert f_ewald[i, axis] == pytest.approx(-slope[2], rel=1e-5, abs=1e-8)
