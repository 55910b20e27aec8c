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
