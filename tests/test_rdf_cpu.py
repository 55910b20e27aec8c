"""The oracle's restatement of AtomRDF (src/analysis.cpp:1556-1600, src/aux/equidistant_table.h:32-40) against an
independent numpy evaluation of the same definitions. The reference ships no rdf fixture for these inputs
(examples/bulk/rdf.dat is produced by its test run, not committed), so this pins the restatement to the written
semantics only."""
import numpy as np
import pytest

from _oraclelib import oracle_sim


def numpy_rdf(xyz, ids, box, id1, id2, dr, slicedir=None, thickness=0.0):
    a, b = xyz[ids == id1], xyz[ids == id2]
    d = a[:, None, :] - b[None, :, :]
    d = np.where(d > box / 2, d - box, np.where(d < -box / 2, d + box, d))   # Geometry::vdist, src/geometry.h:429-458
    keep = np.ones(d.shape[:2], dtype=bool)
    if id1 == id2:
        keep = np.triu(keep, k=1)
    if slicedir is not None and sum(slicedir) > 0:
        s = np.array(slicedir, dtype=float)
        keep &= np.linalg.norm(d * (1 - s), axis=2) < thickness
        r = np.linalg.norm(d * s, axis=2)
    else:
        r = np.sqrt(d[..., 0] ** 2 + d[..., 1] ** 2 + d[..., 2] ** 2)
    bins = np.floor(r[keep] * (1.0 / dr)).astype(int)
    return np.bincount(bins)


@pytest.mark.parametrize("names,extra", [(("Na", "Cl"), {}), (("Na", "Na"), {}), (("Cl", "Na"), {"dr": 0.25}),
                                         (("Na", "Cl"), {"slicedir": [0, 0, 1], "thickness": 5.0})])
def test_oracle_rdf_matches_numpy(bulk_input, names, extra):
    sim = oracle_sim(bulk_input)
    sim.sweep(1)
    cfg = {"name1": names[0], "name2": names[1], "dr": 0.1, "file": "rdf.dat", **extra}
    rid = sim.rdf_create(cfg)
    sim.rdf_sample(rid)
    r, pairs, g = sim.rdf_result(rid)
    xyzq, ids = sim.particles()
    type_id = {"Na": 0, "Cl": 1}
    box = np.array(bulk_input["geometry"]["length"], dtype=float) * np.ones(3)
    want = numpy_rdf(xyzq[:, :3], ids, box, type_id[names[0]], type_id[names[1]], cfg["dr"], extra.get("slicedir"),
                     extra.get("thickness", 0.0))
    assert len(pairs) == len(want)
    assert np.array_equal(pairs, want.astype(np.uint64))
    n1, n2 = (ids == type_id[names[0]]).sum(), (ids == type_id[names[1]]).sum()
    if "slicedir" not in extra:
        assert pairs.sum() == (n1 * (n1 - 1) // 2 if names[0] == names[1] else n1 * n2)
        # g(r) = N <V> / (4 pi r^2 dr sum N): about 1 where the shell lies inside the cell (r < L/2)
        shell = (r > 12) & (r < 20)
        assert abs(g[shell].mean() - 1.0) < 0.05
    assert r[1] == pytest.approx(cfg["dr"]) and r[0] == 0.0
    # a second sample accumulates
    sim.rdf_sample(rid)
    assert np.array_equal(sim.rdf_result(rid)[1], 2 * pairs)


def test_oracle_molecule_rdf_matches_numpy(water_input):
    """MoleculeRDF (src/analysis.cpp:1607-1658): mass centres of the water molecules, distance = sqrt(sqdist)"""
    sim = oracle_sim(water_input)
    sim.sweep(1)
    rid = sim.rdf_create({"type": "molrdf", "name1": "water", "name2": "water", "dr": 0.1, "file": "rdf.dat"})
    sim.rdf_sample(rid)
    r, pairs, g = sim.rdf_result(rid)
    _, cm = sim.groups()
    box = np.array(sim.state_json()["geometry"]["length"], dtype=float) * np.ones(3)   # after a volume move
    d = np.abs(cm[:, None, :] - cm[None, :, :])
    d = d - box * (d > box / 2)                                   # Geometry::sqdist, src/geometry.h:460-470
    dist = np.sqrt((d ** 2).sum(axis=2))[np.triu_indices(len(cm), k=1)]
    want = np.bincount(np.floor(dist * (1.0 / 0.1)).astype(int))
    assert np.array_equal(pairs, want.astype(np.uint64))
    assert pairs.sum() == len(cm) * (len(cm) - 1) // 2
    with pytest.raises(RuntimeError):
        oracle_sim(water_input).rdf_create({"type": "molrdf", "name1": "water", "name2": "nosuch", "dr": 0.1, "file": "x"})
