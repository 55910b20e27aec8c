"""Generates the committed fixtures under tests/golden/ from the reference tree (/root/reference, only
present in the development container) and the oracle. Run from the repo root:

    python tests/golden/make_fixtures.py

Fixtures:
  bulk_input.json      examples/bulk/bulk.yml + bulk.state.json merged into one input document
  minimal_input.json   examples/minimal/minimal.yml + minimal.state.json
  water_input.json     examples/water/ewald.yml with an oracle-generated start state (ewaldscheme PBC)
  widom_input.json     examples/widom/widom.yml (insertion handled by the driver)
  reference_values.json  numbers quoted from the reference's *.out.json and doctest known answers
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference/examples"

from faunus_b200.config import load_input, with_state  # noqa: E402
from _oraclelib import oracle_sim  # noqa: E402


def strip(cfg):
    for key in ("analysis", "mcloop"):
        cfg.pop(key, None)
    cfg["random"] = {"seed": "fixed"}
    return cfg


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f)
    print(name, os.path.getsize(os.path.join(HERE, name)), "bytes")


bulk = strip(with_state(load_input(f"{REF}/bulk/bulk.yml"), json.load(open(f"{REF}/bulk/bulk.state.json"))))
dump("bulk_input.json", bulk)

minimal = strip(with_state(load_input(f"{REF}/minimal/minimal.yml"),
                           json.load(open(f"{REF}/minimal/minimal.state.json"))))
dump("minimal_input.json", minimal)

water = strip(load_input(f"{REF}/water/ewald.yml"))
water["energy"][1]["nonbonded_coulomblj"]["coulomb"]["ewaldscheme"] = "PBC"
sim = oracle_sim(water)  # random insertion with the reference's RNG consumption
state = sim.state_json()
water = with_state(water, state)
dump("water_input.json", water)

widom = load_input(f"{REF}/widom/widom.yml")
widom_analysis = widom["analysis"]
widom = strip(widom)
widom["analysis"] = widom_analysis
dump("widom_input.json", widom)

bulk_out = json.load(open(f"{REF}/bulk/bulk.out.json"))
widom_out = json.load(open(f"{REF}/widom/widom.out.json"))
minimal_out = json.load(open(f"{REF}/minimal/minimal.out.json"))


def find(obj, key):
    if isinstance(obj, dict):
        if key in obj:
            return obj[key]
        for v in obj.values():
            r = find(v, key)
            if r is not None:
                return r
    elif isinstance(obj, list):
        for v in obj:
            r = find(v, key)
            if r is not None:
                return r
    return None


values = {
    "source": "quoted from mlund/faunus examples/*.out.json and doctest known answers (file:line in comments of tests)",
    "bulk": {
        "lB": find(bulk_out, "lB"),
        "systemenergy_init": find(find(bulk_out, "systemenergy"), "init"),
        "systemenergy_final": find(find(bulk_out, "systemenergy"), "final"),
        "relative drift": find(bulk_out, "relative drift"),
        "acceptance": find(find(bulk_out, "transrot"), "acceptance"),
    },
    "widom": {"mu_excess": find(widom_out, "excess")},
    "minimal": {"lB": find(minimal_out, "lB"), "coulomb_cutoff": find(minimal_out, "cutoff")},
    "ewald_doctest": {  # src/energy.cpp:74-99, 249-305, 665-762
        "K_pbc": 2975, "K_ipbc": 846,
        "self_over_lB": -1.0092530088080642, "surface_over_lB": 0.0020943951023931952,
        "reciprocal_over_lB": 0.21303063979675319, "reciprocal_ipbc_over_lB": 0.0865107467,
        "energy_after_move": 103.7300260099, "energy_change": -16.8380445846,
    },
    "andrea_doctest": {  # src/tabulate.h:313-365
        "knots": 19, "coeffs": 108, "r2_1": 0.212991, "r2_2": 0.782554, "c_back": -0.0441931,
    },
    "potentials_doctest": {  # src/potentials.cpp:660-661, 778, 1625
        "lB_epsr80_T298.15": 7.0056973292, "lj_force_x": 0.0142838474, "galore_plain_force_z": 0.1429734149,
    },
}
dump("reference_values.json", values)
