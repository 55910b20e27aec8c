"""Full-size checks (BASELINE.json sizes).

The headline configuration S1 (N = 1e5, exactly `bench.workload()`) is compared with the ORACLE move by move
(`test_s1_matches_oracle`: per-term system energies, u_new / u_old of every trial move, the accept/reject
sequence and the final positions, for one move per launch, windows walked on the host and runs walked on the
device), and S2 (N = 1e6, device cell list) with the oracle's brute-force pair sums (`test_s2_pair_part_matches_oracle`).
The oracle's start-up passes over all pairs / all (k, particle) run in its large-system mode (all host threads,
identical repeats answered from the stored result, oracle/oracle.cpp `fo_set_large_system_mode`); its per-move
path is the serial reference restatement. The remaining tests use size-independent properties:

  * energy bookkeeping: E_final − (E_init + Σ accepted ΔU) ≈ 0 (the reference's own drift check,
    src/montecarlo.cpp:85-99) after thousands of windowed moves,
  * the windowed evaluation (fb_batch_trial) and the one-move-per-launch protocol give the same
    accept/reject trace and energies,
  * sharded = unsharded (Widom slices, system-energy shares),
  * streaming full-energy kernel (all-atomic fast path) == tiled general kernel (forced by a molecular dummy).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EWALD = {"type": "ewald", "epsr": 78.7, "cutoff": 28.0, "alpha": 0.12, "ncutoff": 30, "ewaldscheme": "PBC"}


def s1(moves_per_sweep, **kw):
    from faunus_b200.config import primitive_model
    return primitive_model(n=100_000, molarity=1.0, seed=5489, moves_per_sweep=moves_per_sweep, coulomb=EWALD, **kw)


def sim(cfg, window=None, run=None, run_min=None):
    from faunus_b200.native import B200Simulation
    return B200Simulation(cfg, window=window, run=run, run_min=run_min)


@pytest.fixture(scope="module")
def s1_oracle():
    """256 trial moves of bench.workload() by the oracle: energies per term before and after, trace, positions"""
    import bench
    from _oraclelib import oracle_lib, oracle_sim
    lib = oracle_lib()
    lib.fo_set_large_system_mode(1)
    try:
        o = oracle_sim(bench.workload(moves_per_step=256))
        before = o.system_energy()[1]
        o.trace_enable()
        o.sweep(1)
        after = o.system_energy()[1]
        out = {"before": before, "after": after, "trace": o.trace(), "positions": o.particles()[0], "drift": o.drift()}
        o.close()
    finally:
        lib.fo_set_large_system_mode(0)
    return out


@pytest.mark.parametrize("mode", ["single", "windows", "runs"])
def test_s1_matches_oracle(s1_oracle, mode):
    """The headline workload against the oracle (src/montecarlo.cpp:139-187 move by move). Tolerance: 1e-10 relative
    to the largest energy of the comparison (FP64, summation order only); decisions and positions identical."""
    import bench
    kw = {"single": dict(window=0), "windows": dict(window=64, run=0), "runs": dict(window=64, run=512, run_min=1)}[mode]
    g = sim(bench.workload(moves_per_step=256), **kw)
    assert g.num_particles == 100_000
    before = g.system_energy()[1]
    g.trace_enable()
    g.sweep(1)
    after = g.system_energy()[1]
    o, t = s1_oracle["trace"], g.trace()
    assert len(t["du"]) == len(o["du"]) == 256
    if mode == "runs":
        assert g.run_stats()["windows"] > 0
    scale = np.abs(s1_oracle["before"]).max()
    assert np.abs(before - s1_oracle["before"]).max() <= 1e-10 * scale
    assert np.abs(after - s1_oracle["after"]).max() <= 1e-10 * scale
    assert np.array_equal(t["accepted"], o["accepted"])
    move_scale = max(np.abs(o["u_new"]).max(), np.abs(o["u_old"]).max())
    assert np.abs(t["u_new"] - o["u_new"]).max() <= 1e-10 * move_scale
    assert np.abs(t["u_old"] - o["u_old"]).max() <= 1e-10 * move_scale
    assert np.abs(t["du"] - o["du"]).max() <= 1e-10 * move_scale
    assert np.array_equal(g.particles()[0], s1_oracle["positions"])
    assert abs(g.drift()) < 1e-9 and abs(s1_oracle["drift"]) < 1e-9


def test_s2_pair_part_matches_oracle():
    """S2 (N = 1e6, L = 939.8 A): 24 trial moves through the device cell list against the oracle's
    brute-force pair sums, cutoff scheme without k-space. The oracle skips the Σ_{i<j} over all 5e11 pairs of the
    start configuration (large-system mode 2): only the per-move energies and decisions are compared."""
    from faunus_b200.config import primitive_model
    from _oraclelib import oracle_lib, oracle_sim
    cfg = primitive_model(n=1_000_000, molarity=1.0, seed=5489, moves_per_sweep=24,
                          coulomb={"type": "fanourgakis", "epsr": 78.7, "cutoff": 28.0})
    lib = oracle_lib()
    lib.fo_set_large_system_mode(2)
    try:
        o = oracle_sim(cfg)
        o.trace_enable()
        o.sweep(1)
        to, xo = o.trace(), o.particles()[0]
        o.close()
    finally:
        lib.fo_set_large_system_mode(0)
    g = sim(cfg, window=64)
    g.trace_enable()
    g.sweep(1)
    tg = g.trace()
    assert len(tg["du"]) == len(to["du"]) == 24
    assert np.array_equal(tg["accepted"], to["accepted"])
    # the oracle's trace holds Hamiltonian energies: self-energy term (identical on both sides) + pair sums
    scale = max(np.abs(to["u_new"]).max(), np.abs(to["u_old"]).max())
    assert np.abs(tg["u_new"] - to["u_new"]).max() <= 1e-10 * scale
    assert np.abs(tg["u_old"] - to["u_old"]).max() <= 1e-10 * scale
    assert np.array_equal(g.particles()[0], xo)
    assert abs(g.drift()) < 1e-9


def test_s1_drift_invariant_windowed():
    g = sim(s1(3000), window=64)
    g.trace_enable()
    g.sweep(2)
    tr = g.trace()
    assert len(tr["du"]) == 6000
    assert 0.05 < tr["accepted"].mean() < 0.95
    assert abs(g.drift()) < 1e-9


def test_s1_device_walk_equals_host_walk():
    """N = 1e5: runs of ~400 proposals (6 windows per round trip) decided on the device == windows walked on the host"""
    a, b = sim(s1(3000), window=64, run=0), sim(s1(3000), window=64, run=512)
    for s in (a, b):
        s.trace_enable()
        s.sweep(1)
    ta, tb = a.trace(), b.trace()
    assert len(ta["du"]) == 3000
    assert np.array_equal(ta["accepted"], tb["accepted"])
    # the same corrections in the same order, but the windows are cut differently (proposals on a pending atom
    # end a run where the host walk had already decided the earlier move): equal to rounding
    scale = np.abs(ta["u_new"]).max()
    for key in ("u_new", "u_old"):
        assert np.abs(ta[key] - tb[key]).max() <= 1e-12 * scale
    xa, _ = a.particles()
    xb, _ = b.particles()
    assert np.array_equal(xa, xb)
    wa, wb = a.window_time_ms(), b.window_time_ms()
    assert wb["round_trips"] < wa["round_trips"] / 3
    assert abs(b.drift()) < 1e-9


def test_s1_windowed_equals_single_moves():
    a, b, c = sim(s1(400), window=0), sim(s1(400), window=32), sim(s1(400), window=7)
    for s in (a, b, c):
        s.trace_enable()
        s.sweep(1)
    ta, tb, tc = a.trace(), b.trace(), c.trace()
    assert np.array_equal(ta["accepted"], tb["accepted"]) and np.array_equal(ta["accepted"], tc["accepted"])
    scale = np.abs(ta["u_new"]).max()
    for t in (tb, tc):
        assert np.abs(ta["u_new"] - t["u_new"]).max() <= 1e-10 * scale
        assert np.abs(ta["u_old"] - t["u_old"]).max() <= 1e-10 * scale
    xa, _ = a.particles()
    xb, _ = b.particles()
    assert np.array_equal(xa, xb)
    ea, eb = a.system_energy()[1], b.system_energy()[1]
    assert np.abs(ea - eb).max() <= 1e-10 * np.abs(ea).max()


@pytest.mark.parametrize("window", [32, 64])
def test_s1_cell_list_equals_brute_force(window):
    """N = 1e5: windows through the device cell list (forced; the default starts at 200 000 particles) vs brute
    force; window 64: inside runs walked on the device (the cell list is updated from the device-side commit list)"""
    a, b = sim(s1(600), window=window), sim(s1(600), window=window)
    a.configure_cells(0)
    b.configure_cells(-1)
    for s in (a, b):
        s.trace_enable()
        s.sweep(1)
    ta, tb = a.trace(), b.trace()
    assert np.array_equal(ta["accepted"], tb["accepted"])
    scale = np.abs(ta["u_new"]).max()
    assert np.abs(ta["u_new"] - tb["u_new"]).max() <= 1e-10 * scale
    assert abs(a.drift()) < 1e-9


def test_s1_shards_add_up():
    g = sim(s1(10))
    _, terms = g.system_energy()
    parts = np.array([g.system_energy_shard(r, 4) for r in range(4)])
    scale = np.abs(terms).max()
    assert abs(parts[:, 0].sum() - terms[1]) <= 1e-10 * scale
    assert abs(parts[:, 1].sum() - terms[2]) <= 1e-10 * scale


def test_s1_widom_slices():
    import ctypes as C
    from faunus_b200.config import primitive_model
    cfg = primitive_model(n=100_000, molarity=1.0, seed=5489, moves_per_sweep=10, ghost_pairs=1,
                          coulomb={"type": "fanourgakis", "epsr": 78.7, "cutoff": 28.0})
    a, b = sim(cfg), sim(cfg)
    analysis = {"molecule": "ghost", "ninsert": 1000}
    wa, wb = a.widom_create(analysis), b.widom_create(analysis)
    a.widom_sample(wa, 1)
    n = b.api.widom_prepare(b.handle, wb)
    assert n == 1000
    du = np.zeros(n)
    for first, count in ((0, 333), (333, 333), (666, 334)):
        part = np.zeros(count)
        assert b.api.widom_evaluate_slice(b.handle, wb, first, count, part.ctypes.data_as(C.POINTER(C.c_double))) == 0
        du[first:first + count] = part
    assert b.api.widom_collect(b.handle, wb, du.ctypes.data_as(C.POINTER(C.c_double)), n) == 0
    ra, rb = a.widom_result(wa), b.widom_result(wb)
    assert np.array_equal(ra["last_du"], rb["last_du"])
    assert ra["sum_exp"] == rb["sum_exp"] and ra["sum_exp"] > 0


def test_window_edge_cases():
    """window of one move, capacity one, a two-atom system (every second proposal hits a pending atom)"""
    from conftest import nacl_pair_input, small_electrolyte
    from _oraclelib import oracle_sim
    cfg = nacl_pair_input()
    cfg["energy"][0]["nonbonded_coulomblj"]["coulomb"]["epss"] = 0.0  # tinfoil: windowed path eligible
    cfg["moves"] = [{"transrot": {"molecule": "salt", "repeat": 50}}]
    o = oracle_sim(cfg)
    for window in (1, 16):
        g = sim(cfg, window=window)
        assert g.window == window
        for s in (o, g) if window == 1 else (g,):
            s.trace_enable()
            s.sweep(2)
        assert np.array_equal(o.trace()["accepted"], g.trace()["accepted"])
        assert np.abs(o.trace()["du"] - g.trace()["du"]).max() <= 1e-9 * np.abs(o.trace()["u_new"]).max()
    # ragged sweep length (not a multiple of the window) and inactive particles (ghost group) in the mirror
    cfg = small_electrolyte(n=150, moves_per_sweep=37, ghost_pairs=2,
                            coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 9.0, "alpha": 0.35, "ncutoff": 5})
    o, g = oracle_sim(cfg), sim(cfg, window=16)
    for s in (o, g):
        s.trace_enable()
        s.sweep(5)
    assert np.array_equal(o.trace()["accepted"], g.trace()["accepted"])
    assert abs(g.drift()) < 1e-9


def test_s1_atom_rdf_pair_count():
    """N = 1e5: every Na-Cl pair (2.5e9) and every Na-Na pair (1.25e9) lands in exactly one bin; equal to the
    oracle's loop on the first 4000 ions of the same configuration"""
    cfg = s1(100)
    g = sim(cfg, window=64)
    n = g.num_particles
    for names, expected in ((("Na", "Cl"), (n // 2) ** 2), (("Na", "Na"), (n // 2) * (n // 2 - 1) // 2)):
        rid = g.rdf_create({"name1": names[0], "name2": names[1], "dr": 0.1, "file": "rdf.dat"})
        g.rdf_sample(rid)
        r, pairs, gr = g.rdf_result(rid)
        assert int(pairs.sum()) == expected
        shell = (r > 50) & (r < 200)
        assert abs(gr[shell].mean() - 1.0) < 0.01   # an ideal-gas-like start configuration
