"""The oracle against an evaluation that shares NO code with it: plain numpy on the positions it reports.

`oracle/*.hpp` includes the host headers of the product (`Space`, geometry, moves, the MC engine), so traces and
energies compared between oracle and device share that scaffolding. Here the energies of configurations the oracle's
engine produced — after sweeps of accepted and rejected moves, boundary wraps included — are recomputed from scratch:
minimum image by `numpy.round`, Lennard-Jones with Lorentz-Berthelot mixing and plain Coulomb written out, the Ewald
reciprocal sum from `updateBox` / `updateComplex` / `reciprocalEnergy` / `surfaceEnergy` as the formulas stand in
SURVEY Appendix A (src/energy.cpp:133-186, 191-206, 466-482, 524-531; constants src/units.h). CPU only."""
import numpy as np
import pytest

from _oraclelib import oracle_sim

# src/units.h:19-24, 37-41 (the reference's own, non-CODATA values)
E0, KB, EC, NA = 8.85419e-12, 1.380658e-23, 1.602177e-19, 6.022137e23


def bjerrum_length(epsr, T):
    return EC * EC / (4 * np.pi * E0 * epsr * 1e-10 * KB * T)


def kT_per_kJmol(T):
    return 1e3 / (KB * T * NA)


def min_image(d, box):
    return d - box * np.round(d / box)


def pair_sum(pos, box, fn):
    total = 0.0
    for i in range(len(pos) - 1):
        d = min_image(pos[i + 1:] - pos[i], box)
        total += fn(i, np.arange(i + 1, len(pos)), np.sqrt((d * d).sum(axis=1))).sum()
    return total


def test_minimal_example_energies_and_trace(minimal_input):
    """examples/minimal (40 ions, `nonbonded`: Lennard-Jones LB + plain Coulomb): the oracle's system energy before and
    after 30 sweeps equals the numpy sum over all pairs, and the accepted energy changes of its trace add up to the
    difference — every du of the trace is a difference of true energies"""
    T = minimal_input["temperature"]
    box = np.array(minimal_input["geometry"]["length"], dtype=float)
    atoms = [list(a.values())[0] for a in minimal_input["atomlist"]]
    sigma = np.array([a["sigma"] for a in atoms])
    eps = np.array([a["eps"] for a in atoms]) * kT_per_kJmol(T)
    lB = bjerrum_length(80.0, T)

    def energy(xyzq, ids):
        def u(i, j, r):
            s = 0.5 * (sigma[ids[i]] + sigma[ids[j]])
            e = np.sqrt(eps[ids[i]] * eps[ids[j]])
            x = (s / r) ** 6
            return 4 * e * (x * x - x) + lB * xyzq[i, 3] * xyzq[j, 3] / r
        return pair_sum(xyzq[:, :3], box, u)

    sim = oracle_sim(minimal_input)
    before = energy(*sim.particles())
    total, terms = sim.system_energy()
    assert total == pytest.approx(before, rel=1e-11)
    sim.trace_enable()
    sim.sweep(30)
    xyzq, ids = sim.particles()
    assert np.all(np.abs(xyzq[:, :3]) <= box / 2)  # Geometry::boundary after every move
    after = energy(xyzq, ids)
    assert sim.system_energy()[0] == pytest.approx(after, rel=1e-11)
    trace = sim.trace()
    assert trace["accepted"].sum() > 100
    assert trace["du"][trace["accepted"] != 0].sum() == pytest.approx(after - before, rel=1e-9, abs=1e-9)


def ewald_numpy(pos, q, L, alpha, ncutoff, lB, epss):
    """PolicyIonIon::updateBox / updateComplex / reciprocalEnergy / surfaceEnergy for a cubic cell, tinfoil unless
    epss ≥ 1, spherical sum"""
    nc = int(np.ceil(ncutoff))
    n = np.array([(x, y, z) for x in range(0, nc + 1) for y in range(-nc, nc + 1) for z in range(-nc, nc + 1)
                  if (x, y, z) != (0, 0, 0) and (x * x + y * y + z * z) / ncutoff ** 2 <= 1.0], dtype=float)
    k = 2 * np.pi * n / L
    k2 = (k * k).sum(axis=1)
    factor = np.where(n[:, 0] > 0, 2.0, 1.0)
    A = factor * np.exp(-k2 / (4 * alpha * alpha)) / k2
    phase = pos @ k.T
    Q = (q[:, None] * np.exp(1j * phase)).sum(axis=0)
    V = L ** 3
    reciprocal = 2 * np.pi * lB / V * (A * np.abs(Q) ** 2).sum()
    surface = 0.0
    if epss >= 1:
        mu = (q[:, None] * pos).sum(axis=0)
        surface = 2 * np.pi / ((2 * epss + 1) * V) * (mu @ mu) * lB
    return len(k), reciprocal, surface


@pytest.mark.parametrize("epss", [0.0, 80.0])
def test_ewald_term_against_numpy_after_moves(epss):
    """40 ions, Ewald with and without the surface term: K, reciprocal + surface energy of the oracle's term after
    sweeps (Q(k) kept current by partial updates) equal the numpy evaluation from the final positions"""
    from test_forces_cpu import _salt_input
    cfg = _salt_input(epss)
    coulomb = cfg["energy"][0]["nonbonded_coulombwca"]["coulomb"]
    sim = oracle_sim(cfg)
    for sweeps in (0, 40):
        if sweeps:
            sim.sweep(sweeps)
        xyzq, _ = sim.particles()
        K, reciprocal, surface = ewald_numpy(xyzq[:, :3], xyzq[:, 3], 20.0, coulomb["alpha"], coulomb["ncutoff"],
                                             bjerrum_length(coulomb["epsr"], cfg["temperature"]), epss)
        terms = sim.system_energy()[1]
        assert sim.info()["energy"][-1]["ewald"]["wavefunctions"] == K
        assert terms[2] == pytest.approx(reciprocal + surface, rel=1e-10)
