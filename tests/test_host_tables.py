"""Host-side (configuration-time) lowering of the pair-potential JSON schema in the product library,
checked against the oracle on CPU: identical Coulomb spline tables, consistent mixing matrices."""
import ctypes as C
import json

import numpy as np
import pytest

import faunus_b200.native as native
from _oraclelib import c_double_p, oracle_lib, pair_energy
from test_oracle_golden import ATOMS_ABC, _input, _kT_per_kJmol

SCHEMES = [
    {"type": "plain", "epsr": 80}, {"type": "fanourgakis", "epsr": 1, "cutoff": 14},
    {"type": "ewald", "epsr": 78.7, "cutoff": 14, "alpha": 0.22, "ncutoff": 5},
    {"type": "ewald", "epsr": 78.7, "cutoff": 14, "alpha": 0.22, "ncutoff": 5, "debyelength": 30},
    {"type": "qpotential", "epsr": 80, "cutoff": 20, "order": 4}, {"type": "yukawa", "epsr": 80, "debyelength": 25},
    {"type": "yukawa", "epsr": 80, "debyelength": 25, "shift": True, "cutoff": 40},
    {"type": "poisson", "epsr": 80, "cutoff": 15, "C": 3, "D": 3},
    {"type": "poisson", "epsr": 80, "cutoff": 15, "C": 2, "D": 1, "debyelength": 20},
    {"type": "wolf", "epsr": 80, "cutoff": 12, "alpha": 0.2}, {"type": "zahn", "epsr": 80, "cutoff": 12, "alpha": 0.2},
    {"type": "fennell", "epsr": 80, "cutoff": 12, "alpha": 0.2},
    {"type": "zerodipole", "epsr": 80, "cutoff": 12, "alpha": 0.2},
    {"type": "reactionfield", "epsr": 80, "epsrf": 1000, "cutoff": 12},
]


def _table(fn, scheme, T):
    kn, co = np.zeros(2048), np.zeros(6 * 2048)
    vals = [C.c_double() for _ in range(4)]
    n = fn(json.dumps(scheme).encode(), T, kn.ctypes.data_as(c_double_p), co.ctypes.data_as(c_double_p), 2048,
           *[C.byref(v) for v in vals])
    assert n >= 2
    return n, kn[:n].copy(), co[:6 * (n - 1)].copy(), [v.value for v in vals]


@pytest.mark.parametrize("T", [298.15, 1100.0])
@pytest.mark.parametrize("scheme", SCHEMES, ids=lambda s: s["type"] + ("+k" if "debyelength" in s else ""))
def test_coulomb_tables_identical(scheme, T):
    lib = native.load()
    lib.fbh_coulomb_table.restype = C.c_int
    lib.fbh_coulomb_table.argtypes = [C.c_char_p, C.c_double, c_double_p, c_double_p, C.c_int] + [c_double_p] * 4
    a = _table(oracle_lib().fo_coulomb_table, scheme, T)
    b = _table(lib.fbh_coulomb_table, scheme, T)
    assert a[0] == b[0]
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert a[3] == b[3]  # lB, cutoff, kappa, self-energy prefactor


def _pair_tables(cfg, name):
    lib = native.load()
    lib.fbh_pair_tables_json.restype = C.c_int
    lib.fbh_pair_tables_json.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    text = json.dumps(cfg).encode()
    n = lib.fbh_pair_tables_json(text, name.encode(), None, 0)
    assert n > 0, native.sim_library().error()
    buf = C.create_string_buffer(n)
    lib.fbh_pair_tables_json(text, name.encode(), buf, n)
    return json.loads(buf.value.decode())


def test_mixing_matrices_match_oracle():
    """σ², 4ε matrices of the product reproduce the oracle's LJ/WCA energies (src/potentials.cpp:672-703)"""
    cfg = _input(ATOMS_ABC, [{"nonbonded": {"default": [
        {"lennardjones": {"mixing": "LB", "custom": [{"A C": {"eps": 0.5, "sigma": 8}}]}}]}}])
    t = _pair_tables(cfg, "nonbonded")
    n = t["n_types"]
    s2 = np.array(t["lj_s2"]).reshape(n, n)
    e4 = np.array(t["lj_e4"]).reshape(n, n)
    assert np.allclose(s2, s2.T) and np.allclose(e4, e4.T)
    r = np.array([3.0, 6.0, 9.0])
    for a in range(n):
        for b in range(n):
            x = (s2[a, b] / r ** 2) ** 3
            np.testing.assert_allclose(e4[a, b] * (x * x - x), pair_energy(cfg, "nonbonded", a, b, r), rtol=1e-12)
    assert e4[0, 2] == pytest.approx(4 * 0.5 * _kT_per_kJmol(), rel=1e-12)
    assert s2[0, 1] == pytest.approx(25.0)


def test_functor_flags_and_cutoffs():
    atoms = [{"A": {"q": 1.0, "sigma": 2.0, "eps": 0.1}}, {"B": {"q": -1.0, "sigma": 4.0, "eps": 0.05}}]
    cfg = _input(atoms, [{"nonbonded": {
        "default": [{"coulomb": {"epsr": 80.0, "type": "plain"}}, {"lennardjones": {"mixing": "LB"}}],
        "A B": [{"coulomb": {"epsr": 80.0, "type": "plain"}}, {"wca": {"mixing": "LB"}}],
        "cutoff_g2g": 12.0}}])
    t = _pair_tables(cfg, "nonbonded")
    flags = np.array(t["flags"], dtype=int).reshape(2, 2)
    assert flags[0, 0] == native.TERM_COULOMB_SPLINED | native.TERM_LJ
    assert flags[0, 1] == flags[1, 0] == native.TERM_COULOMB_SPLINED | native.TERM_WCA
    assert t["g2g_cutoff_squared"] == [144.0]


def test_splined_pair_tables_reproduce_exact():
    """nonbonded_splined: per-pair r² tables (src/potentials.cpp:1513-1595) stay within utol of the exact sum"""
    atoms = [{"A": {"q": 1.0, "sigma": 4.0, "eps": 0.5}}, {"B": {"q": -1.0, "sigma": 4.0, "eps": 0.5}}]
    cfg = _input(atoms, [{"nonbonded_splined": {"default": [
        {"lennardjones": {"mixing": "LB"}}, {"coulomb": {"type": "fanourgakis", "epsr": 80, "cutoff": 20}}]}}])
    t = _pair_tables(cfg, "nonbonded_splined")
    off = np.array(t["sp_offset"], dtype=int)
    knots, coeffs = np.array(t["sp_knots"]), np.array(t["sp_coeffs"])
    for pair in range(4):
        k = knots[off[pair]:off[pair + 1]]
        assert len(k) >= 2 and np.all(np.diff(k) > 0)
        assert k[0] == pytest.approx(t["sp_rmin2"][pair]) and k[-1] == pytest.approx(t["sp_rmax2"][pair])
    exact = _input(atoms, [{"nonbonded": cfg["energy"][0]["nonbonded_splined"]}])
    r = np.linspace(np.sqrt(t["sp_rmin2"][1]) + 0.05, np.sqrt(t["sp_rmax2"][1]) - 0.05, 50)
    u_exact = pair_energy(exact, "nonbonded", 0, 1, r)
    u_spline = pair_energy(cfg, "nonbonded_splined", 0, 1, r)
    assert np.max(np.abs(u_exact - u_spline)) < 2e-3  # utol = 1e-3 at the 11 check points per interval
    # the product's own table evaluates to the same numbers as the oracle's spline
    first = off[1]
    c = coeffs[6 * (first - 1):]  # pair 1: coefficient blocks of the preceding pair: (nk-1) blocks
    k = knots[off[1]:off[2]]
    pos = np.searchsorted(k, r * r, side="left") - 1
    dz = r * r - k[pos]
    val = np.zeros_like(r)
    for i in range(5, 0, -1):
        val = dz * (val + c[6 * pos + i])
    val = val + c[6 * pos]
    np.testing.assert_allclose(val, u_spline, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("block,expect", [
    ({"cutoff_g2g": 12.0}, [[12.0, 12.0], [12.0, 12.0]]),                                   # old style default
    ({"cutoff_g2g": {"default": 13.0}}, [[13.0, 13.0], [13.0, 13.0]]),                       # new style default
    ({"cutoff_g2g": {"default": 13.0, "M Q": 14.0}}, [[13.0, 14.0], [14.0, 13.0]]),          # custom
    ({"cutoff_g2g": {"M Q": 11.0}}, [[None, 11.0], [11.0, None]]),                           # custom, no default: none
])
def test_group_cutoff_doctest(block, expect):
    """src/energy.cpp:1955-2002 (`[Faunus] GroupCutoff`): the four ways to write `cutoff_g2g`, in the oracle's parser and
    in the product's table (the mass-centre cutoffs² that go into fb_config.g2g_cutoff_squared)"""
    cfg = {"temperature": 298.15,
           "atomlist": [{"A": {"sigma": 4.0}}, {"B": {"sigma": 2.4}}],
           "moleculelist": [{"M": {"structure": [{"A": [0.0, 0.0, 0.0]}, {"B": [1.0, 0.0, 0.0]}]}},
                            {"Q": {"structure": [{"A": [0.0, 0.0, 0.0]}, {"B": [1.0, 0.0, 0.0]}]}}],
           "geometry": {"type": "cuboid", "length": 100}, "groups": [], "particles": [], "moves": [],
           "energy": [{"nonbonded": dict({"default": [{"lennardjones": {"mixing": "LB"}}]}, **block)}]}
    none = 1.7976931348623157e308  # pc::max_value: "no cutoff" (the reference stores sqrt(max)² the same way)
    want = np.array([[none if v is None else v * v for v in row] for row in expect])
    lib = oracle_lib()
    lib.fo_group_cutoffs.restype = C.c_int
    lib.fo_group_cutoffs.argtypes = [C.c_char_p, C.c_char_p, c_double_p, C.c_int]
    got = np.zeros(4)
    n = lib.fo_group_cutoffs(json.dumps(cfg).encode(), json.dumps(cfg["energy"][0]["nonbonded"]).encode(),
                             got.ctypes.data_as(c_double_p), 4)
    assert n == 4
    np.testing.assert_allclose(got.reshape(2, 2), want, rtol=1e-15)
    product = np.array(_pair_tables(cfg, "nonbonded")["g2g_cutoff_squared"]).reshape(2, 2)
    np.testing.assert_allclose(product, want, rtol=1e-15)
