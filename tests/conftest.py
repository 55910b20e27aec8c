import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def reference_values():
    return load_golden("reference_values.json")


@pytest.fixture(scope="session")
def bulk_input():
    return load_golden("bulk_input.json")


@pytest.fixture(scope="session")
def minimal_input():
    return load_golden("minimal_input.json")


@pytest.fixture(scope="session")
def water_input():
    return load_golden("water_input.json")


@pytest.fixture(scope="session")
def widom_input():
    return load_golden("widom_input.json")


def nacl_pair_input(ewaldscheme="PBC"):
    """Two ions in a 10 Å box with the Ewald settings of the reference doctests (src/energy.cpp:665-762)"""
    return {
        "temperature": 298.15,
        "geometry": {"type": "cuboid", "length": 10},
        "atomlist": [{"Na": {"q": 1.0, "sigma": 0.1, "eps": 0.0, "dp": 1.0}},
                     {"Cl": {"q": -1.0, "sigma": 0.1, "eps": 0.0, "dp": 1.0}}],
        "moleculelist": [{"salt": {"atoms": ["Na", "Cl"], "atomic": True}}],
        "groups": [{"id": 0, "size": 2, "cm": [0, 0, 0], "atomic": True, "compressible": False}],
        "particles": [{"id": 0, "pos": [0, 0, 0], "q": 1.0}, {"id": 1, "pos": [1, 0, 0], "q": -1.0}],
        "energy": [{"nonbonded_coulomblj": {
            "lennardjones": {"mixing": "LB"},
            "coulomb": {"type": "ewald", "epsr": 1.0, "alpha": 0.894427190999916, "epss": 1.0,
                        "ncutoff": 11.0, "spherical_sum": True, "cutoff": 5.0, "ewaldscheme": ewaldscheme}}}],
        "moves": [{"transrot": {"molecule": "salt"}}],
    }


def small_electrolyte(n=400, seed=7, **kw):
    from faunus_b200.config import primitive_model
    return primitive_model(n=n, molarity=1.0, seed=seed, **kw)


def water_with_salt(water, n_pairs=8, seed=3, coulomb=None, volume_move=False):
    """examples/water plus an atomic NaCl group: `moltransrot` and `transrot` moves interleave"""
    import copy
    import numpy as np
    cfg = copy.deepcopy(water)
    rng = np.random.RandomState(seed)
    box = np.array(cfg["geometry"]["length"], dtype=float)
    pos = np.array([p["pos"] for p in cfg["particles"]])
    cfg["atomlist"] += [{"Na": {"q": 1.0, "sigma": 2.6, "eps": 0.3, "dp": 1.2, "mw": 22.99}},
                        {"Cl": {"q": -1.0, "sigma": 4.0, "eps": 0.3, "dp": 1.2, "mw": 35.45}}]
    cfg["moleculelist"] += [{"salt": {"atoms": ["Na", "Cl"] * n_pairs, "atomic": True}}]
    ions = []
    while len(ions) < 2 * n_pairs:
        x = (rng.uniform(size=3) - 0.5) * box
        d = pos - x
        d -= box * np.round(d / box)
        if (np.einsum("ij,ij->i", d, d) > 2.4 ** 2).all():
            pos = np.vstack([pos, x])
            ions.append({"id": 2 + len(ions) % 2, "pos": x.tolist(), "q": 1.0 if len(ions) % 2 == 0 else -1.0})
    cfg["groups"].append({"id": 1, "size": 2 * n_pairs, "cm": [0, 0, 0], "atomic": True, "compressible": False})
    cfg["particles"] += ions
    cfg["moves"] = [m for m in cfg["moves"] if volume_move or "volume" not in m] + [
        {"transrot": {"molecule": "salt", "repeat": 40}}]
    if coulomb is not None:
        for term in cfg["energy"]:
            for name, body in term.items():
                if name.startswith("nonbonded"):
                    body["coulomb"] = coulomb
    return cfg


def one_water_in_salt(water, n_pairs=40, coulomb=None):
    """ONE rigid water molecule in an NaCl solution (what `virtualtranslate` wants: exactly one active molecule)"""
    import copy
    single = copy.deepcopy(water)
    single["groups"] = single["groups"][:1]
    single["particles"] = single["particles"][:3]
    cfg = water_with_salt(single, n_pairs=n_pairs, coulomb=coulomb)
    cfg["moves"] = [{"moltransrot": {"molecule": "water", "dp": 0.4, "dprot": 0.4, "repeat": 5}},
                    {"transrot": {"molecule": "salt", "repeat": 60}}]
    cfg["energy"] = [t for t in cfg["energy"] if "isobaric" not in t]
    return cfg
