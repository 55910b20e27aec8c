"""The oracle (CPU restatement) against every known answer the reference holds for the ΔU path.
CPU only. file:line citations point into mlund/faunus."""
import ctypes as C
import json

import numpy as np
import pytest

from _oraclelib import c_double_p, oracle_lib, oracle_sim, pair_energy
from conftest import nacl_pair_input


def test_andrea_spline_doctest(reference_values):
    """src/tabulate.h:313-365"""
    ref = reference_values["andrea_doctest"]
    lib = oracle_lib()
    kn, co, nc = np.zeros(64), np.zeros(512), C.c_int()
    n = lib.fo_andrea_test(2e-6, 1e-4, 0, 10, kn.ctypes.data_as(c_double_p), 64, co.ctypes.data_as(c_double_p), 512,
                           C.byref(nc))
    assert n == ref["knots"] and nc.value == ref["coeffs"]
    assert kn[0] == pytest.approx(0.0) and kn[n - 1] == pytest.approx(10.0)
    assert kn[1] == pytest.approx(ref["r2_1"], rel=1e-5)
    assert kn[2] == pytest.approx(ref["r2_2"], rel=1e-5)
    assert co[0] == pytest.approx(2.0) and co[1] == pytest.approx(0.0, abs=1e-6) and co[2] == pytest.approx(0.5, rel=1e-5)
    assert co[nc.value - 1] == pytest.approx(ref["c_back"], rel=1e-5)
    f = lambda x: 0.5 * x * np.sin(x) + 2
    for x in (1e-9, 5.0, 10.0):
        assert lib.fo_andrea_test_eval(2e-6, 1e-4, 0, 10, x) == pytest.approx(f(x), rel=1e-5)


@pytest.mark.parametrize("scheme,key_k,key_rec", [("PBC", "K_pbc", "reciprocal_over_lB"),
                                                  ("PBCEigen", "K_pbc", "reciprocal_over_lB"),
                                                  ("IPBC", "K_ipbc", "reciprocal_ipbc_over_lB")])
def test_ewald_policies_doctest(reference_values, scheme, key_k, key_rec):
    """src/energy.cpp:74-99 (K) and :249-305 (self, surface, reciprocal)"""
    ref = reference_values["ewald_doctest"]
    cfg = {"epsr": 1.0, "alpha": 0.894427190999916, "epss": 1.0, "ncutoff": 11.0, "spherical_sum": True,
           "cutoff": 5.0, "ewaldscheme": scheme}
    xyzq = np.array([[0, 0, 0, 1.0], [1, 0, 0, -1.0]])
    out, lB = np.zeros(3), C.c_double()
    K = oracle_lib().fo_ewald_kat(json.dumps(cfg).encode(), 298.15, 10.0, xyzq.ctypes.data_as(c_double_p), 2,
                                  out.ctypes.data_as(c_double_p), C.byref(lB))
    assert K == ref[key_k]
    assert lB.value == pytest.approx(560.4557863339663, rel=1e-12)
    assert out[0] / lB.value == pytest.approx(ref["self_over_lB"], rel=1e-12)
    assert out[1] / lB.value == pytest.approx(ref["surface_over_lB"], rel=1e-12)
    assert out[2] / lB.value == pytest.approx(ref[key_rec], rel=1e-9)


def test_ewald_term_energy_change_doctest(reference_values):
    """src/energy.cpp:665-762: energy, full update after a displacement, partial updates (`all` and
    index list) with sync round trips"""
    ref = reference_values["ewald_doctest"]
    lB = 560.4557863339663
    reference_energy = (ref["surface_over_lB"] + ref["reciprocal_over_lB"]) * lB
    sim = oracle_sim(nacl_pair_input())
    _, terms = sim.system_energy()
    assert len(terms) == 3  # [self, nonbonded, ewald], src/energy.h:462-477 + energy.cpp:1134-1160
    assert terms[2] == pytest.approx(reference_energy, rel=1e-10)
    for use_all in (True, False):
        s = oracle_sim(nacl_pair_input())
        before = s.system_energy()[1][2]
        s.trial_set(0, [0], [[0.1, 0.1, 0.1]], all=use_all, internal=True)
        s.trial_commit(True)
        after = s.system_energy()[1][2]
        assert after == pytest.approx(ref["energy_after_move"], rel=1e-9)
        assert after - before == pytest.approx(ref["energy_change"], rel=1e-9)
        s.trial_set(0, [0], [[0.0, 0.0, 0.0]], all=use_all, internal=True)  # and back again
        s.trial_commit(True)
        assert s.system_energy()[1][2] == pytest.approx(reference_energy, rel=1e-9)


def test_bjerrum_lengths(reference_values, bulk_input, minimal_input):
    """src/potentials.cpp:660 and the lB printed in bulk.out.json / minimal.out.json"""
    lib = oracle_lib()

    def lB(coulomb, T):
        kn, co = np.zeros(256), np.zeros(6 * 256)
        vals = [C.c_double() for _ in range(4)]
        lib.fo_coulomb_table(json.dumps(coulomb).encode(), T, kn.ctypes.data_as(c_double_p),
                             co.ctypes.data_as(c_double_p), 256, *[C.byref(v) for v in vals])
        return vals[0].value, vals[1].value

    assert lB({"type": "plain", "epsr": 80}, 298.15)[0] == pytest.approx(
        reference_values["potentials_doctest"]["lB_epsr80_T298.15"], rel=1e-10)
    assert lB({"type": "fanourgakis", "epsr": 1, "cutoff": 14}, 1100)[0] == pytest.approx(
        reference_values["bulk"]["lB"], rel=1e-15)
    value, cutoff = lB({"type": "plain", "epsr": 80}, 300)
    assert value == pytest.approx(reference_values["minimal"]["lB"], rel=1e-15)
    assert cutoff == pytest.approx(reference_values["minimal"]["coulomb_cutoff"], rel=1e-15)


ATOMS_ABC = [{"A": {"sigma": 2.0, "eps": 0.9}}, {"B": {"sigma": 8.0, "eps": 0.1}}, {"C": {"sigma": 5.0, "eps": 1.1}}]


def _input(atomlist, energy, T=298.15):
    return {"temperature": T, "atomlist": atomlist,
            "moleculelist": [{"M": {"atoms": [list(a)[0] for a in atomlist], "atomic": True}}], "energy": energy}


def _kT_per_kJmol(T=298.15):
    return 1e3 / (T * 1.380658e-23 * 6.022137e23)


def test_lennard_jones_mixing_doctest():
    """src/potentials.cpp:711-780: LB / geometric / custom mixing at r = 0.9 nm vs the closed form"""
    d = 9.0
    lj = lambda sigma, eps: 4 * eps * _kT_per_kJmol() * ((sigma / d) ** 12 - (sigma / d) ** 6)
    lb = _input(ATOMS_ABC, [{"nonbonded": {"default": [{"lennardjones": {"mixing": "LB"}}]}}])
    assert pair_energy(lb, "nonbonded", 0, 0, [d])[0] == pytest.approx(lj(2.0, 0.9), rel=1e-12)
    assert pair_energy(lb, "nonbonded", 0, 1, [d])[0] == pytest.approx(lj(5.0, 0.3), rel=1e-12)
    geo = _input(ATOMS_ABC, [{"nonbonded": {"default": [{"lennardjones": {"mixing": "geometric"}}]}}])
    assert pair_energy(geo, "nonbonded", 0, 1, [d])[0] == pytest.approx(lj(4.0, 0.3), rel=1e-12)
    custom = _input(ATOMS_ABC, [{"nonbonded": {"default": [
        {"lennardjones": {"mixing": "LB", "custom": [{"A B": {"eps": 0.5, "sigma": 8}}]}}]}}])
    assert pair_energy(custom, "nonbonded", 0, 1, [d])[0] == pytest.approx(lj(8.0, 0.5), rel=1e-12)
    assert pair_energy(custom, "nonbonded", 0, 0, [d])[0] == pytest.approx(lj(2.0, 0.9), rel=1e-12)


def test_hard_sphere_doctest():
    """src/potentials.cpp:979-1015"""
    atoms = [{"A": {"sigma": 2}}, {"B": {"sigma": 8}}]
    hs = _input(atoms, [{"nonbonded": {"default": [{"hardsphere": {"mixing": "arithmetic"}}]}}])
    assert list(pair_energy(hs, "nonbonded", 0, 0, [2.01, 1.99])) == [0.0, np.inf]
    assert list(pair_energy(hs, "nonbonded", 0, 1, [5.01, 4.99])) == [0.0, np.inf]
    custom = _input(atoms, [{"nonbonded": {"default": [{"hardsphere": {"custom": [{"A B": {"sigma": 6}}]}}]}}])
    assert list(pair_energy(custom, "nonbonded", 0, 1, [6.01, 5.99])) == [0.0, np.inf]


def test_wca_closed_form():
    """src/potentials.h:151-160 (cut at r² > σ²·2^(1/3), shift ¼)"""
    wca = _input(ATOMS_ABC, [{"nonbonded": {"default": [{"wca": {"mixing": "LB"}}]}}])
    sigma, eps = 5.0, np.sqrt(0.9 * 0.1) * _kT_per_kJmol()
    r = np.array([4.0, 5.0, 5.6, 5.62, 7.0])
    expect = np.where(r * r > sigma ** 2 * 1.2599210498948732, 0.0,
                      4 * eps * ((sigma / r) ** 12 - (sigma / r) ** 6 + 0.25))
    np.testing.assert_allclose(pair_energy(wca, "nonbonded", 0, 1, r), expect, rtol=1e-9, atol=1e-14)


def test_functor_potential_doctest():
    """src/potentials.cpp:1333-1366: coulomb-plain ≡ Coulomb, "A B" override = coulomb + wca, HS for C-C"""
    atoms = [{"A": {"q": 1.0, "r": 1.1, "eps": 0.1}}, {"B": {"q": -1.0, "r": 2.0, "eps": 0.05}},
             {"C": {"r": 1.0}}]
    functor = _input(atoms, [{"nonbonded": {
        "default": [{"coulomb": {"epsr": 80.0, "type": "plain"}}],
        "A B": [{"coulomb": {"epsr": 80.0, "type": "plain"}}, {"wca": {"mixing": "LB"}}],
        "C C": [{"hardsphere": {}}]}}])
    lB = 7.0056973292
    r = 2.0
    assert pair_energy(functor, "nonbonded", 0, 0, [r])[0] == pytest.approx(lB / r, rel=1e-9)
    assert pair_energy(functor, "nonbonded", 1, 1, [r])[0] == pytest.approx(lB / r, rel=1e-9)
    sigma, eps = 0.5 * (2.2 + 4.0), np.sqrt(0.1 * 0.05) * _kT_per_kJmol()
    wca = 4 * eps * ((sigma / r) ** 12 - (sigma / r) ** 6 + 0.25)
    assert pair_energy(functor, "nonbonded", 0, 1, [r])[0] == pytest.approx(-lB / r + wca, rel=1e-9)
    assert list(pair_energy(functor, "nonbonded", 2, 2, [2.02, 1.98])) == [0.0, np.inf]
    pm = _input(atoms, [{"nonbonded_pmwca": {"coulomb": {"epsr": 80.0}, "wca": {"mixing": "LB"}}}])
    assert pair_energy(pm, "nonbonded_pmwca", 0, 1, [r])[0] == pytest.approx(-lB / r + wca, rel=1e-9)


def test_bulk_example(reference_values, bulk_input):
    """examples/bulk: state file energies within the reference's own 5 % tolerance
    (examples/CMakeLists.txt:243-249); nonbonded vs the golden final value, drift invariant"""
    ref = reference_values["bulk"]
    sim = oracle_sim(bulk_input)
    total, terms = sim.system_energy()
    assert len(terms) == 2  # [particle-self-energy, nonbonded]
    assert terms[1] == pytest.approx(ref["systemenergy_final"][1], rel=0.05)
    assert total == pytest.approx(ref["systemenergy_init"], rel=0.05)
    sim.trace_enable()
    sim.sweep(3)
    trace = sim.trace()
    assert len(trace["du"]) == 3 * 2304
    assert 0.3 < trace["accepted"].mean() < 0.5  # golden acceptance 0.393
    assert abs(sim.drift()) < 1e-9  # src/montecarlo.cpp:85-99 warns above 1e-9


def test_minimal_example(minimal_input):
    sim = oracle_sim(minimal_input)
    total, terms = sim.system_energy()
    assert np.isfinite(total) and len(terms) == 2
    sim.sweep(50)
    assert abs(sim.drift()) < 1e-9


def test_widom_example(reference_values, widom_input):
    """examples/widom: hard-sphere excess chemical potential, analytic −ln(1 − (r/R)³) = 0.13353;
    reference tolerance 1 % (examples/CMakeLists.txt:280-284)"""
    cfg = dict(widom_input)
    analysis = cfg.pop("analysis")[0]["widom"]
    sim = oracle_sim(cfg)
    wid = sim.widom_create(analysis)
    sim.widom_sample(wid, 100000)
    res = sim.widom_result(wid)
    mu = -np.log(res["sum_exp"] / res["count"])
    assert res["count"] == 100000 * analysis["ninsert"]
    assert mu == pytest.approx(reference_values["widom"]["mu_excess"], rel=0.01)
    assert mu == pytest.approx(-np.log(1 - (2.0 / 4.0) ** 3), rel=0.01)


def test_water_example_runs(water_input):
    """examples/water (ewald.yml, NOCHECKS in the reference): rigid-body + volume moves keep the drift small"""
    sim = oracle_sim(water_input)
    total, terms = sim.system_energy()
    assert len(terms) == 4  # [isobaric, self, nonbonded, ewald]
    sim.sweep(2)
    assert abs(sim.drift()) < 1e-8


# ---------------------------------------------------------------------------------------------------------
# The self energy of the cutoff schemes: a KNOWN, documented disagreement with the 2019 golden output
# ---------------------------------------------------------------------------------------------------------
def test_bulk_self_energy_follows_the_documented_definition(reference_values, bulk_input):
    """examples/bulk, term 0 (`particle-self-energy`). The reference defines (docs/_docs/energy.md:272-282)
    U_self = −½ Σ_i lim_{r→0}(u_ii − ũ_ii) = ½ S'(0) · lB · Σ q_i² / R_c; for Fanourgakis S(q) = 1 − 7/4 q + …
    (docs/_docs/energy.md:199) that is −7/8 · lB · N / R_c. The functor itself lives in the absent dependency
    (`pot.selfEnergyFunctor`, src/potentials.cpp:1599-1604): oracle and product implement the documented formula."""
    ref = reference_values["bulk"]
    _, terms = oracle_sim(bulk_input).system_energy()
    expected = -0.875 * ref["lB"] * 2304 / 14.0
    assert terms[0] == pytest.approx(expected, rel=1e-12)
    # the same definition gives the standard Wolf / Ewald self terms, which ARE pinned by the reference:
    # ewald: −α/√π · lB · Σq² (src/energy.cpp:514-517, reference doctest value −1.0092530088080642·lB for two unit charges)
    alpha = 0.894427190999916
    assert -alpha / np.sqrt(np.pi) * 2 == pytest.approx(-1.0092530088080642, rel=1e-12)


@pytest.mark.xfail(strict=True, reason=(
    "examples/bulk/bulk.out.json (git revision 76f393dc, 2019-10-31, when Faunus still carried its own CoulombGalore "
    "class) lists particle-self-energy = −24999.880050 = −1.0·lB·N/R_c, i.e. a Fanourgakis self prefactor of −1 "
    "instead of the −7/8 that the reference's documented definition (docs/_docs/energy.md:272-282, S'(0)/2) gives. "
    "The reference's own check cannot see the difference: scripts/jsoncompare.py skips lists of numbers (the `final` "
    "per-term energies) and compares `init` at 5 %, which −7/8 passes (3.2 %). The current functor is in the "
    "un-vendored mlund/coulombgalore@4055f58: unpinned. The term is constant at fixed N (cancels in every ΔU of "
    "a canonical move); it shifts absolute energies and the Widom ΔU of charged insertions by (1/8)·lB·q²/R_c."))
def test_bulk_self_energy_against_the_2019_golden(reference_values, bulk_input):
    ref = reference_values["bulk"]
    _, terms = oracle_sim(bulk_input).system_energy()
    assert terms[0] == pytest.approx(ref["systemenergy_final"][0], rel=1e-6)


# ---------------------------------------------------------------------------------------------------------
# RNG draw order of a translational displacement: which order is ASSUMED (unpinned by reference data)
# ---------------------------------------------------------------------------------------------------------
class _Mt19937:
    """std::mt19937 (default seed 5489) + libstdc++ uniform_real_distribution<double>(0, 1): generate_canonical
    takes two 32-bit draws, low word first; numpy's MT19937 legacy seeding is init_genrand, the same stream."""

    def __init__(self, seed=5489):
        self.rs = np.random.RandomState(seed)

    def raw(self):
        return int(self.rs.randint(0, 2 ** 32, dtype=np.uint64))

    def uniform(self):
        lo, hi = self.raw(), self.raw()
        return (lo + hi * 4294967296.0) / 18446744073709551616.0

    def unit_vector(self):  # src/core.cpp:249-261
        while True:
            p = np.array([self.uniform() - 0.5 for _ in range(3)])
            if p @ p <= 0.25:
                return p / np.sqrt(p @ p)


def test_displacement_draw_order_is_the_gcc_order():
    """`randomUnitVector(slump, dir) * dp * slump()` (src/move.cpp:229): C++ leaves the evaluation order of the two
    operands open. GCC evaluates the operands of the overloaded `operator*` right to left — the scalar `slump()` is
    drawn BEFORE the unit vector — clang left to right. The restated moves assume the GCC order (host/moves.hpp:8-10);
    nothing in the reference pins it (the goldens of the examples are compared at 1–5 %, and were written by a clang
    build). This test documents the assumption on ONE free particle: group pick and atom pick take one 32-bit draw
    each (ranges of one element), then the scalar, then the unit vector; ΔU = 0 is always accepted."""
    cfg = {
        "temperature": 298.15, "geometry": {"type": "cuboid", "length": 1000.0},
        "atomlist": [{"X": {"q": 0.0, "sigma": 1.0, "eps": 0.0, "dp": 2.5}}],
        "moleculelist": [{"gas": {"atoms": ["X"], "atomic": True}}],
        "groups": [{"id": 0, "size": 1, "cm": [0, 0, 0], "atomic": True, "compressible": False}],
        "particles": [{"id": 0, "pos": [1.0, 2.0, 3.0], "q": 0.0}],
        "energy": [{"nonbonded_coulomblj": {"lennardjones": {"mixing": "LB"}, "coulomb": {"type": "plain", "epsr": 80.0}}}],
        "moves": [{"transrot": {"molecule": "gas", "repeat": 1}}],
        "random": {"seed": "fixed"},
    }
    sim = oracle_sim(cfg)
    sim.trace_enable()
    sim.sweep(1)
    assert list(sim.trace()["accepted"]) == [1]
    moved = sim.particles()[0][0, :3] - np.array([1.0, 2.0, 3.0])

    def displacement(scalar_first):
        g = _Mt19937()
        g.raw(), g.raw()  # group pick, atom pick (uniform_int_distribution over one element: one draw each)
        if scalar_first:
            s = g.uniform()
            u = g.unit_vector()
        else:
            u = g.unit_vector()
            s = g.uniform()
        return u * 2.5 * s

    assert np.allclose(moved, displacement(scalar_first=True), rtol=0, atol=1e-12)
    assert not np.allclose(moved, displacement(scalar_first=False), rtol=0, atol=1e-6)


def test_space_update_particles_mass_centres_doctest():
    """src/space.cpp:627-682 (`[Faunus] Space::updateParticles`, "Group update"): two SPC/E-like waters of
    `SpaceFactory::makeWater` (:720-742; mw 15.999 / 1.007), positions (0,0,0), (3,3,3), (6,6,6) written over one group,
    the other, and across both: mass centres 0.5031366235, 0.1677122078, 5.8322877922"""
    atoms = [{"OW": {"sigma": 3.166, "eps": 0.65, "q": -0.8476, "mw": 15.999}},
             {"HW": {"sigma": 2.0, "eps": 0.0, "q": 0.4238, "mw": 1.007}}]
    structure = [{"OW": [2.3, 6.28, 1.13]}, {"HW": [1.37, 6.26, 1.5]}, {"HW": [2.31, 5.89, 0.21]}]
    # (the periodic mass centre is taken relative to the old one — the doctest sets it to x = −1 first — so the waters
    # start near x = −1: a molecule that jumps by more than half a cell is outside what the algorithm is meant for)
    start = [[-1.0, 3.0, 7.0], [-1.9, 3.0, 7.0], [-0.7, 3.9, 7.0], [-1.0, 3.0, 3.0], [-1.9, 3.0, 3.0], [-0.7, 3.9, 3.0]]
    cfg = {"temperature": 298.15, "random": {"seed": "fixed"}, "geometry": {"type": "cuboid", "length": 20},
           "atomlist": atoms, "moleculelist": [{"water": {"structure": structure}}],
           "groups": [{"id": 0, "size": 3, "cm": [-1.03, 3.05, 7.0], "atomic": False, "compressible": False},
                      {"id": 0, "size": 3, "cm": [-1.03, 3.05, 3.0], "atomic": False, "compressible": False}],
           "particles": [{"id": 0 if i % 3 == 0 else 1, "pos": p, "q": -0.8476 if i % 3 == 0 else 0.4238} for i, p in enumerate(start)],
           "energy": [{"nonbonded_coulomblj": {"coulomb": {"type": "plain", "epsr": 80}, "lennardjones": {"mixing": "LB"}}}],
           "moves": [{"moltransrot": {"molecule": "water", "dp": 0.1, "dprot": 0.1, "repeat": 1}}]}
    sim = oracle_sim(cfg)
    positions = [[0.0, 0.0, 0.0], [3.0, 3.0, 3.0], [6.0, 6.0, 6.0]]
    sim.trial_set(1, [0, 1, 2], positions, all=True)
    sim.trial_commit(True)
    assert sim.groups()[1][1][0] == pytest.approx(0.5031366235, rel=1e-9)
    sim.trial_set(0, [0, 1, 2], positions, all=True)
    sim.trial_commit(True)
    assert sim.groups()[1][0][0] == pytest.approx(0.5031366235, rel=1e-9)
    # "both groups affected": the three positions land on the atoms 1, 2 (first water) and 3 (oxygen of the second)
    sim.trial_set(0, [1, 2], positions[:2], all=False)
    sim.trial_commit(True)
    sim.trial_set(1, [0], positions[2:], all=False)
    sim.trial_commit(True)
    cm = sim.groups()[1]
    assert cm[0][0] == pytest.approx(0.1677122078, rel=1e-9) and cm[1][0] == pytest.approx(5.8322877922, rel=1e-9)
