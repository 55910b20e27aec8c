"""Input handling: Faunus YAML/JSON input documents and the synthetic benchmark systems.

The reference converts YAML input to JSON with ``scripts/yason.py`` before piping it into the
binary (``examples/CMakeLists.txt``); this module does the same conversion in-process and builds
the synthetic systems defined in SURVEY.md §8(d) (S1 "pm-1e5", S2 "pm-1e6", Widom variant).
"""
from __future__ import annotations

import json
from typing import Optional

import numpy as np
import yaml


def load_input(path: str, **template) -> dict:
    """YAML or JSON input file → dict; ``{{name}}`` placeholders are filled from ``template``."""
    text = open(path).read()
    for key, value in template.items():
        text = text.replace("{{" + key + "}}", str(value))
    if path.endswith(".json"):
        return json.loads(text)
    return yaml.safe_load(text)


def with_state(config: dict, state: dict) -> dict:
    """Merge a reference ``state.json`` (geometry, groups, particles) into an input document."""
    out = dict(config)
    out["geometry"] = state["geometry"]
    out["groups"] = state["groups"]
    out["particles"] = [{"id": p["id"], "pos": p["pos"], "q": p["q"]} for p in state["particles"]]
    out.pop("insertmolecules", None)
    return out


def lattice_positions(n: int, box: float, jitter: float, seed: int = 5489) -> np.ndarray:
    """Ions on a randomly filled simple-cubic lattice with uniform jitter (vectorised; N = 1e6 in < 1 s).

    Sites are visited in a random permutation so that the alternating Na/Cl ids are spatially mixed.
    """
    rng = np.random.RandomState(seed)
    m = int(np.ceil(n ** (1.0 / 3.0)))
    spacing = box / m
    sites = rng.permutation(m ** 3)[:n]
    ijk = np.stack(np.unravel_index(sites, (m, m, m)), axis=1).astype(np.float64)
    pos = (ijk + 0.5) * spacing - 0.5 * box
    pos += (rng.random_sample((n, 3)) - 0.5) * 2.0 * min(jitter, 0.45 * spacing)
    return pos


def electrolyte_positions(n: int, box: float, min_distance: float, seed: int = 5489) -> np.ndarray:
    """Uniform random ion positions in a cubic box, rejecting overlaps below ``min_distance``.

    Pure-Python rejection sampling with a cell grid; use for small systems (tests).
    Deterministic for a given seed.
    """
    rng = np.random.RandomState(seed)
    ncell = max(1, int(box / max(min_distance, 1e-9)))
    cell = box / ncell
    grid: dict = {}
    pos = np.empty((n, 3))
    count = 0
    md2 = min_distance * min_distance
    while count < n:
        cand = (rng.random_sample((max(1024, n - count), 3)) - 0.5) * box
        for p in cand:
            c = tuple(((p + 0.5 * box) / cell).astype(int) % ncell)
            ok = True
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    for dz in (-1, 0, 1):
                        key = ((c[0] + dx) % ncell, (c[1] + dy) % ncell, (c[2] + dz) % ncell)
                        for q in grid.get(key, ()):
                            d = np.abs(p - q)
                            d -= box * (d > 0.5 * box)
                            if d @ d < md2:
                                ok = False
                                break
                        if not ok:
                            break
                    if not ok:
                        break
                if not ok:
                    break
            if ok:
                grid.setdefault(c, []).append(p)
                pos[count] = p
                count += 1
                if count == n:
                    break
    return pos


def primitive_model(n: int = 100_000, molarity: float = 1.0, coulomb: Optional[dict] = None,
                    energy_name: str = "nonbonded_coulombwca", sigma: float = 4.0, eps: float = 0.2,
                    dp: float = 4.0, temperature: float = 298.15, seed: int = 5489,
                    ghost_pairs: int = 0, min_distance: float = 3.5, summation_policy: str = "serial",
                    extra_nonbonded: Optional[dict] = None, placement: str = "auto", moves_per_sweep: int = 1) -> dict:
    """restricted primitive model 1:1 electrolyte, one atomic group.

    N ions (alternating Na+/Cl−, as ``atoms: [Na, Cl]`` insertion gives) in a cubic PBC box at
    ``molarity`` (1.0 M, N = 1e5 → L = 436.2 Å), WCA cores, splined Coulomb of the given type.
    ``ghost_pairs`` appends that many inactive Na/Cl pairs to the group capacity for Widom insertion.
    """
    if n % 2:
        raise ValueError("n must be even")
    avogadro = 6.022137e23
    npairs = n // 2
    volume = npairs / (molarity * avogadro / 1e27)
    box = volume ** (1.0 / 3.0)
    if coulomb is None:
        coulomb = {"type": "ewald", "epsr": 78.7, "cutoff": 14.0, "alpha": 0.22, "ncutoff": 30}
    if placement == "auto":
        placement = "random" if n <= 5000 else "lattice"
    if placement == "random":
        pos = electrolyte_positions(n, box, min_distance, seed)
    else:
        pos = lattice_positions(n, box, jitter=2.0, seed=seed)
    particles = [{"id": i % 2, "pos": pos[i].tolist(), "q": 1.0 if i % 2 == 0 else -1.0} for i in range(n)]
    groups = [{"id": 0, "size": n, "cm": [0.0, 0.0, 0.0], "atomic": True, "compressible": False}]
    moleculelist = [{"salt": {"atoms": ["Na", "Cl"], "atomic": True}}]
    if ghost_pairs > 0:
        moleculelist.append({"ghost": {"atoms": ["Na", "Cl"], "atomic": True}})
        for _ in range(ghost_pairs):
            particles.append({"id": 0, "pos": [0.0, 0.0, 0.0], "q": 1.0})
            particles.append({"id": 1, "pos": [0.0, 0.0, 0.0], "q": -1.0})
        groups.append({"id": 1, "size": 0, "capacity": 2 * ghost_pairs, "cm": [0.0, 0.0, 0.0],
                       "atomic": True, "compressible": False})
    nonbonded = {"coulomb": dict(coulomb), "summation_policy": summation_policy}
    if energy_name == "nonbonded_coulombwca":
        nonbonded["wca"] = {"mixing": "LB"}
    elif energy_name == "nonbonded_coulomblj":
        nonbonded["lennardjones"] = {"mixing": "LB"}
    elif energy_name == "nonbonded_pm":
        nonbonded = {"coulomb": {"epsr": coulomb["epsr"]}, "hardsphere": {"mixing": "arithmetic"},
                     "summation_policy": summation_policy}
    elif energy_name == "nonbonded_pmwca":
        nonbonded = {"coulomb": {"epsr": coulomb["epsr"]}, "wca": {"mixing": "LB"},
                     "summation_policy": summation_policy}
    if extra_nonbonded:
        nonbonded.update(extra_nonbonded)
    return {
        "temperature": temperature,
        "random": {"seed": "fixed"},
        "geometry": {"type": "cuboid", "length": [box, box, box]},
        "atomlist": [{"Na": {"q": 1.0, "sigma": sigma, "eps": eps, "dp": dp}},
                     {"Cl": {"q": -1.0, "sigma": sigma, "eps": eps, "dp": dp}}],
        "moleculelist": moleculelist,
        "groups": groups,
        "particles": particles,
        "energy": [{energy_name: nonbonded}],
        "moves": [{"transrot": {"molecule": "salt", "repeat": moves_per_sweep}}],
    }
