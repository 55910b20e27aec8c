"""Parity of the B200 path (through the C ABI / adaptor terms) with the oracle on identical inputs.
Tolerance: 1e-10 relative on energies (FP64, summation order only; BASELINE.json north_star), identical
accept/reject traces for a fixed seed."""
import numpy as np
import pytest

from _oraclelib import oracle_sim
from conftest import nacl_pair_input, small_electrolyte

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def b200_sim(cfg, window=None, run=None):
    from faunus_b200.native import B200Simulation
    return B200Simulation(cfg, window=window, run=run, run_min=1)   # run_min=1: runs even for a single window


def assert_close(a, b, rtol=RTOL, scale=None):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape
    inf = np.isinf(a) | np.isinf(b)
    assert np.array_equal(a[inf], b[inf])
    ref = np.abs(a[~inf]) if scale is None else scale
    assert np.all(np.abs(a[~inf] - b[~inf]) <= rtol * np.maximum(ref, 1e-300) + 1e-300), \
        f"max abs diff {np.abs(a[~inf] - b[~inf]).max()} (ref scale {np.max(ref) if np.size(ref) else 0})"


def pair_of_sims(cfg, window=None):
    return oracle_sim(cfg), b200_sim(cfg, window)


def electrolyte_variants():
    ew = {"type": "ewald", "epsr": 78.7, "cutoff": 9.0, "alpha": 0.35, "ncutoff": 6}
    return {
        "coulombwca_ewald": small_electrolyte(coulomb=ew),
        "coulombwca_ewald_surface": small_electrolyte(coulomb=dict(ew, epss=80.0)),
        "coulomblj_fanourgakis": small_electrolyte(energy_name="nonbonded_coulomblj",
                                                   coulomb={"type": "fanourgakis", "epsr": 78.7, "cutoff": 12.0}),
        "coulombwca_yukawa": small_electrolyte(coulomb={"type": "yukawa", "epsr": 78.7, "debyelength": 9.0}),
        "coulombwca_qpot": small_electrolyte(coulomb={"type": "qpotential", "epsr": 78.7, "cutoff": 12.0, "order": 3}),
        "coulombwca_wolf": small_electrolyte(coulomb={"type": "wolf", "epsr": 78.7, "cutoff": 12.0, "alpha": 0.2}),
        "coulombwca_zerodipole": small_electrolyte(coulomb={"type": "zerodipole", "epsr": 78.7, "cutoff": 12.0, "alpha": 0.2}),
        "coulombwca_reactionfield": small_electrolyte(coulomb={"type": "reactionfield", "epsr": 78.7, "epsrf": 40.0,
                                                                "cutoff": 12.0}),
        "coulombwca_poisson": small_electrolyte(coulomb={"type": "poisson", "epsr": 78.7, "cutoff": 12.0, "C": 4, "D": 3}),
        "pm": small_electrolyte(energy_name="nonbonded_pm", coulomb={"epsr": 78.7}, sigma=3.0),
        "pmwca": small_electrolyte(energy_name="nonbonded_pmwca", coulomb={"epsr": 78.7}),
    }


def functor_variants():
    base = small_electrolyte(coulomb={"type": "plain", "epsr": 78.7})
    functor = dict(base)
    functor["energy"] = [{"nonbonded": {
        "default": [{"lennardjones": {"mixing": "LB"}}, {"coulomb": {"type": "fanourgakis", "epsr": 78.7, "cutoff": 12}}],
        "Na Cl": [{"coulomb": {"type": "fanourgakis", "epsr": 78.7, "cutoff": 12}}, {"wca": {"mixing": "LB"}}],
        "Cl Cl": [{"coulomb": {"type": "fanourgakis", "epsr": 78.7, "cutoff": 12}}, {"hardsphere": {"custom": [{"Cl Cl": {"sigma": 3.2}}]}}]}}]
    splined = dict(base)
    splined["energy"] = [{"nonbonded_splined": {
        "default": [{"lennardjones": {"mixing": "LB"}}, {"coulomb": {"type": "fanourgakis", "epsr": 78.7, "cutoff": 12}}]}}]
    cached = dict(base)  # NonbondedCached<SplinedPotential> (src/energy.cpp:1311-1315): the energies of `nonbonded_splined`
    cached["energy"] = [{"nonbonded_cached": splined["energy"][0]["nonbonded_splined"]}]
    return {"functor": functor, "splined": splined, "cached": cached}


ALL_VARIANTS = {**electrolyte_variants(), **functor_variants()}


#: 0 = one move per launch (updateState/energy/sync protocol); > 0 = windowed evaluation (fb_batch_trial);
#: 64 = runs of windows walked on the device (fb_run_submit)
WINDOWS = [0, 32, 64]


@pytest.mark.parametrize("window", WINDOWS)
@pytest.mark.parametrize("name", sorted(ALL_VARIANTS))
def test_system_energy_and_moves(name, window):
    """Full energy per term, then 300 single-ion trial moves: u_new/u_old per move and the trace"""
    cfg = ALL_VARIANTS[name]
    o, g = pair_of_sims(cfg, window)
    if window and "surface" not in name:
        assert g.window == window
        assert (g.run > 0) == (window == 64)
    eo, to = o.system_energy()
    eg, tg = g.system_energy()
    assert len(to) == len(tg)
    assert_close(to, tg, scale=np.abs(to).max())
    for s in (o, g):
        s.trace_enable()
        s.sweep(300)
    a, b = o.trace(), g.trace()
    assert np.array_equal(a["accepted"], b["accepted"])
    scale = np.abs(a["u_new"][np.isfinite(a["u_new"])]).max()
    assert_close(a["u_new"], b["u_new"], scale=scale)
    assert_close(a["u_old"], b["u_old"], scale=scale)
    assert abs(g.drift()) < 1e-9
    assert_close(o.system_energy()[1], g.system_energy()[1], scale=np.abs(to).max())
    assert g.launch_count > 0


@pytest.mark.parametrize("window", [0, 7, 64])
def test_bulk_example_trace(bulk_input, window):
    """examples/bulk state: identical accept/reject sequence over 3 sweeps (6912 moves)"""
    o, g = pair_of_sims(bulk_input, window)
    assert_close(o.system_energy()[1], g.system_energy()[1], scale=7e4)
    for s in (o, g):
        s.trace_enable()
        s.sweep(3)
    a, b = o.trace(), g.trace()
    assert len(a["du"]) == 3 * 2304
    assert np.array_equal(a["accepted"], b["accepted"])
    assert_close(a["u_new"], b["u_new"], scale=np.abs(a["u_new"]).max())
    assert abs(g.drift()) < 1e-9


def test_bulk_example_trace_1e5_moves(bulk_input):
    """BASELINE.json north star: with the same seed the accept/reject sequence over 10^5 moves is identical
    on the bundled example (examples/bulk: N = 2304, Fanourgakis + LJ; 44 sweeps = 101 376 moves, windowed)"""
    o, g = pair_of_sims(bulk_input, 64)
    for s in (o, g):
        s.trace_enable()
        s.sweep(44)
    a, b = o.trace(), g.trace()
    assert len(a["du"]) == 44 * 2304 > 100_000
    assert np.array_equal(a["accepted"], b["accepted"])
    finite = np.isfinite(a["du"])
    assert np.array_equal(finite, np.isfinite(b["du"]))
    assert np.abs(a["du"][finite] - b["du"][finite]).max() <= 1e-10 * np.abs(a["u_new"][np.isfinite(a["u_new"])]).max()
    xo, _ = o.particles()
    xg, _ = g.particles()
    assert np.array_equal(xo, xg)
    assert abs(g.drift()) < 1e-9


@pytest.mark.parametrize("window", [0, 16])
def test_minimal_example(minimal_input, window):
    o, g = pair_of_sims(minimal_input, window)
    assert_close(o.system_energy()[1], g.system_energy()[1], scale=abs(o.system_energy()[0]))
    for s in (o, g):
        s.trace_enable()
        s.sweep(100)
    assert np.array_equal(o.trace()["accepted"], g.trace()["accepted"])


@pytest.mark.parametrize("window", [0, 64])
def test_water_example(water_input, window):
    """examples/water: rigid SPC/E, mass-centre cutoff, Ewald partial updates for 3-atom moves,
    volume moves (full updates + box change); one move per launch and windows of rigid-molecule moves"""
    o, g = pair_of_sims(water_input, window)
    eo, to = o.system_energy()
    eg, tg = g.system_energy()
    assert len(to) == 4
    assert_close(to, tg, scale=np.abs(to).max())
    for s in (o, g):
        s.trace_enable()
        s.sweep(3)
    a, b = o.trace(), g.trace()
    assert len(a["du"]) > 700
    assert np.array_equal(a["move_id"], b["move_id"])
    assert np.array_equal(a["accepted"], b["accepted"])
    scale = np.abs(to).max()
    assert_close(a["du"], b["du"], rtol=1e-9, scale=scale)  # ΔU of full energies: cancellation of ~1e4 kT sums
    assert (a["move_id"] == 1).any(), "no volume move sampled"
    assert_close(o.system_energy()[1], g.system_energy()[1], scale=scale)
    xo, _ = o.particles()
    xg, _ = g.particles()
    assert np.array_equal(xo, xg)
    t = g.window_time_ms()
    assert (t["windows"] > 20 and t["moves"] > 600) if window else t["windows"] == 0


@pytest.mark.parametrize("coulomb", [None, {"type": "fanourgakis", "epsr": 1, "cutoff": 9}])
@pytest.mark.parametrize("window", [5, 64])
def test_water_with_salt_windows(water_input, coulomb, window):
    """rigid-molecule windows and single-atom windows interleave (a window is of one kind), with and without
    k-space; short windows (5) exercise atom-capacity splits"""
    from conftest import water_with_salt
    cfg = water_with_salt(water_input, coulomb=coulomb)
    o, g = pair_of_sims(cfg, window)
    to = o.system_energy()[1]
    scale = np.abs(to).max()
    assert_close(to, g.system_energy()[1], scale=scale)
    for s in (o, g):
        s.trace_enable()
        s.sweep(4)
    a, b = o.trace(), g.trace()
    assert len(a["du"]) > 900 and set(a["move_id"]) == {0, 1}
    assert np.array_equal(a["move_id"], b["move_id"])
    assert np.array_equal(a["accepted"], b["accepted"])
    per_move = np.maximum(np.abs(a["u_new"]), np.abs(a["u_old"]))  # the example starts with overlapping molecules
    finite = np.isfinite(per_move)
    assert np.all(np.abs(a["du"] - b["du"])[finite] <= RTOL * per_move[finite])
    assert 0.05 < a["accepted"].mean() < 0.95
    assert_close(o.system_energy()[1], g.system_energy()[1], scale=scale)
    xo, _ = o.particles()
    xg, _ = g.particles()
    assert np.array_equal(xo, xg)
    assert g.window_time_ms()["moves"] > 900


def test_ewald_doctest_values(reference_values):
    """The reference's Energy::Ewald known answers on the device (src/energy.cpp:665-762)"""
    ref = reference_values["ewald_doctest"]
    lB = 560.4557863339663
    for scheme in ("PBC", "PBCEigen"):
        g = b200_sim(nacl_pair_input(scheme))
        terms = g.system_energy()[1]
        assert terms[2] == pytest.approx((ref["surface_over_lB"] + ref["reciprocal_over_lB"]) * lB, rel=1e-10)
    for use_all in (True, False):
        g = b200_sim(nacl_pair_input())
        before = g.system_energy()[1][2]
        g.trial_set(0, [0], [[0.1, 0.1, 0.1]], all=use_all, internal=True)
        g.trial_commit(True)
        after = g.system_energy()[1][2]
        assert after == pytest.approx(ref["energy_after_move"], rel=1e-9)
        assert after - before == pytest.approx(ref["energy_change"], rel=1e-9)
    g = b200_sim(nacl_pair_input("IPBC"))
    o = oracle_sim(nacl_pair_input("IPBC"))
    assert g.system_energy()[1][2] == pytest.approx(o.system_energy()[1][2], rel=1e-10)


@pytest.mark.parametrize("scheme,ncutoff,n,alpha", [("PBC", 6, 400, 0.35), ("PBCEigen", 6, 400, 0.35), ("PBC", 11, 700, 0.5),
                                                    ("PBCEigen", 13.5, 70, 0.5), ("PBC", 34, 130, 1.2),
                                                    ("IPBC", 6, 400, 0.35), ("IPBC", 19, 90, 0.8)])
def test_full_q_matrix_product(scheme, ncutoff, n, alpha):
    """fb_ewald_update_full (ewaldFullGemmKernel, fb_fullq.cuh: Q = [X·Y]·[Z] on the FP64 tensor path) against
    numpy's Σ_j q_j e^{ik·r_j} over the downloaded k-vectors (src/energy.cpp:191-206; PBCEigen sums the imaginary part
    WITHOUT the charges, :208-217; IPBC is the real product of the cosines), tiles of 1 … 8 column groups, two z windows at
    ncutoff 34, ragged particle ranges"""
    from faunus_b200 import native
    lib = native.load()
    cfg = small_electrolyte(n=n, coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 9.0, "alpha": alpha, "ncutoff": ncutoff,
                                          "ewaldscheme": scheme})
    g = b200_sim(cfg)
    xyzq, _ = g.particles()
    assert lib.fb_ewald_update_full(g.ctx, 0) == 0
    kmax = (int(np.ceil(ncutoff)) + 1) * (2 * int(np.ceil(ncutoff)) + 1) ** 2
    q, kv = np.zeros(2 * kmax), np.zeros(3 * kmax)
    assert lib.fb_ewald_download(g.ctx, 0, q.ctypes.data_as(native.c_double_p), kv.ctypes.data_as(native.c_double_p), None) == 0
    kv = kv.reshape(-1, 3)
    K = int(np.flatnonzero(np.abs(kv).sum(axis=1) > 0).max()) + 1
    assert K > 100
    Q = q[:2 * K:2] + 1j * q[1:2 * K:2]
    ref = np.zeros(K, dtype=complex)
    for j0 in range(0, len(xyzq), 64):
        ph = kv[:K] @ xyzq[j0:j0 + 64, :3].T
        w = xyzq[j0:j0 + 64, 3]
        if scheme == "IPBC":  # q cos(kx x) cos(ky y) cos(kz z), src/energy.cpp:414-430: the real product of the same kernel
            ref += np.cos(kv[:K, None, :] * xyzq[None, j0:j0 + 64, :3]).prod(axis=2) @ w
            continue
        ref += np.cos(ph) @ w + 1j * (np.sin(ph).sum(axis=1) if scheme == "PBCEigen" else np.sin(ph) @ w)
    assert np.abs(Q - ref).max() <= 1e-12 * n
    # slabs of tile columns (fb_system_energy_shard: the same product over a range of tiles, Q not stored) add up to the
    # reciprocal energy of the numpy sum, whatever the number of slabs
    if scheme != "PBCEigen":
        aks = np.zeros(kmax)
        assert lib.fb_ewald_download(g.ctx, 0, None, None, aks.ctypes.data_as(native.c_double_p)) == 0
        box = np.array(cfg["geometry"]["length"], dtype=float) * np.ones(3)
        full = g.system_energy_shard(0, 1)[1]
        expect = (aks[:K] * np.abs(ref) ** 2).sum()
        for size in (2, 5):
            parts = [g.system_energy_shard(r, size)[1] for r in range(size)]
            assert abs(sum(parts) - full) <= 1e-11 * abs(full)
            assert sum(1 for x in parts if x != 0.0) >= 2
        prefactor = full / expect          # 2π lB / V
        assert prefactor == pytest.approx(2 * np.pi * 7.1 / box.prod(), rel=0.05)   # (lB ≈ 7.1 Å at 78.7, 298 K)


def test_full_q_matrix_product_cuboid_and_inactive():
    """The same product in a box with three different side lengths (non-spherical sum: every |n| ≤ n_cutoff) and with
    inactive slots (a ghost group of capacity 4, empty): inactive particles carry no weight in either part"""
    from faunus_b200 import native
    lib = native.load()
    for scheme in ("PBC", "PBCEigen"):
        cfg = small_electrolyte(n=300, ghost_pairs=2,
                                coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 9.0, "alpha": 0.4, "ncutoff": 7,
                                         "spherical_sum": False, "ewaldscheme": scheme})
        scale = np.array([1.3, 0.8, 1.0])
        cfg["geometry"]["length"] = (np.array(cfg["geometry"]["length"]) * scale).tolist()
        for p in cfg["particles"]:
            p["pos"] = (np.array(p["pos"]) * scale).tolist()
        g = b200_sim(cfg)
        xyzq, _ = g.particles()
        assert len(xyzq) == 304
        assert lib.fb_ewald_update_full(g.ctx, 0) == 0
        kmax = 8 * 15 * 15
        q, kv = np.zeros(2 * kmax), np.zeros(3 * kmax)
        assert lib.fb_ewald_download(g.ctx, 0, q.ctypes.data_as(native.c_double_p), kv.ctypes.data_as(native.c_double_p), None) == 0
        kv = kv.reshape(-1, 3)
        K = int(np.flatnonzero(np.abs(kv).sum(axis=1) > 0).max()) + 1
        assert K == 8 * 15 * 15 - 1       # nx = 0 … 7, ny, nz = −7 … 7, without k = 0
        Q = q[:2 * K:2] + 1j * q[1:2 * K:2]
        ph = kv[:K] @ xyzq[:300, :3].T    # the four inactive slots are left out
        w = xyzq[:300, 3]
        ref = np.cos(ph) @ w + 1j * (np.sin(ph).sum(axis=1) if scheme == "PBCEigen" else np.sin(ph) @ w)
        assert np.abs(Q - ref).max() <= 1e-12 * 300


def test_full_q_and_full_pair_paths_agree(monkeypatch):
    """The kernels the product and the screened pair sum replaced stay selectable for comparisons (FAUNUS_B200_FULLQ=cells:
    one block per k-cell; FAUNUS_B200_FULLPAIR=fp64: all-FP64 pair sum; read at fb_create): same Q(k), same energies, same
    slabs"""
    from faunus_b200 import native
    lib = native.load()
    cfg = small_electrolyte(n=1200, coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 9.0, "alpha": 0.35, "ncutoff": 9})
    results = []
    for fullq, fullpair in (("gemm", "screen"), ("cells", "fp64")):
        monkeypatch.setenv("FAUNUS_B200_FULLQ", fullq)
        monkeypatch.setenv("FAUNUS_B200_FULLPAIR", fullpair)
        g = b200_sim(cfg)
        assert lib.fb_ewald_update_full(g.ctx, 0) == 0
        q = np.zeros(2 * 10 * 19 * 19)
        assert lib.fb_ewald_download(g.ctx, 0, q.ctypes.data_as(native.c_double_p), None, None) == 0
        results.append((q, np.array(g.system_energy()[1]), np.array([g.system_energy_shard(r, 3) for r in range(3)])))
    (qa, ea, sa), (qb, eb, sb) = results
    assert np.abs(qa).max() > 1 and np.abs(qa - qb).max() <= 1e-12 * np.abs(qb).max()
    assert_close(ea, eb, scale=np.abs(eb).max())
    assert_close(sa.sum(axis=0), sb.sum(axis=0), scale=np.abs(eb).max())


def test_reject_restores_state():
    """trial → reject → the next evaluation sees the accepted state again (sync direction)"""
    cfg = small_electrolyte(coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 9.0, "alpha": 0.35, "ncutoff": 6})
    o, g = pair_of_sims(cfg)
    rng = np.random.RandomState(3)
    xyzq, _ = g.particles()
    for step in range(40):
        i = int(rng.randint(0, len(xyzq)))
        new = xyzq[i, :3] + rng.uniform(-2, 2, 3)
        uo = o.trial_set(0, [i], [new])
        ug = g.trial_set(0, [i], [new])
        assert_close(uo, ug, scale=max(abs(uo[0]), 1.0))
        accept = bool(step % 3 == 0)
        o.trial_commit(accept)
        g.trial_commit(accept)
        if accept:
            xyzq[i, :3] = new
    assert_close(o.system_energy()[1], g.system_energy()[1], scale=abs(o.system_energy()[0]))


def test_energy_without_update_state():
    """Callers like SystemEnergy / Widom call energy(change) on the accepted Hamiltonian with no
    updateState and no sync: the term must be idempotent"""
    cfg = small_electrolyte(coulomb={"type": "fanourgakis", "epsr": 78.7, "cutoff": 12.0})
    o, g = pair_of_sims(cfg)
    for idx in ([5], [1, 7, 20], []):
        kw = dict(group=0, internal=True, indices=idx, all=not idx)
        e1 = g.energy(0, **kw)
        e2 = g.energy(0, **kw)
        assert e1 == e2
        assert_close([o.energy(0, **kw)], [e1])


def test_widom_batched_matches_sequential():
    """Batched ghost insertions (one launch) == sequential energy(change) calls == oracle"""
    cfg = small_electrolyte(coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 9.0, "alpha": 0.35, "ncutoff": 6},
                            ghost_pairs=1)
    analysis = {"molecule": "ghost", "ninsert": 64}
    o, g, g_seq = oracle_sim(cfg), b200_sim(cfg), b200_sim(cfg)
    wo = o.widom_create(analysis)
    wg = g.widom_create(analysis)
    ws = g_seq.widom_create(dict(analysis, batched=False))
    for s, w in ((o, wo), (g, wg), (g_seq, ws)):
        s.widom_sample(w, 3)
    ro, rg, rs = o.widom_result(wo), g.widom_result(wg), g_seq.widom_result(ws)
    assert ro["count"] == rg["count"] == rs["count"] == 3 * 64
    scale = np.abs(ro["last_du"]).max()
    assert_close(ro["last_du"], rg["last_du"], scale=scale)
    assert_close(ro["last_du"], rs["last_du"], scale=scale)
    assert rg["sum_exp"] == pytest.approx(ro["sum_exp"], rel=1e-9)


def test_widom_example(reference_values, widom_input):
    """examples/widom on the device: hard-sphere μ_ex within the reference's 1 % of 0.13353"""
    cfg = dict(widom_input)
    analysis = dict(cfg.pop("analysis")[0]["widom"], ninsert=4096)
    o, g = pair_of_sims(cfg)
    wo, wg = o.widom_create(analysis), g.widom_create(analysis)
    o.widom_sample(wo, 50)
    g.widom_sample(wg, 50)
    ro, rg = o.widom_result(wo), g.widom_result(wg)
    assert np.array_equal(ro["last_du"], rg["last_du"])  # 0 or +inf
    assert ro["sum_exp"] == rg["sum_exp"]
    mu = -np.log(rg["sum_exp"] / rg["count"])
    assert mu == pytest.approx(reference_values["widom"]["mu_excess"], rel=0.01)


def test_window_matches_single_moves():
    """fb_batch_trial through the raw C ABI: window energies + corrections reproduce one-at-a-time
    fb_trial_energy / fb_trial_commit on the same proposals (Ewald electrolyte, every 2nd move accepted)"""
    import ctypes as C
    import faunus_b200.native as native
    lib = native.load()
    cfg = small_electrolyte(n=600, coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 9.0, "alpha": 0.35, "ncutoff": 6})
    ga, gb = b200_sim(cfg, 0), b200_sim(cfg, 0)
    xyzq, ids = ga.particles()
    rng = np.random.RandomState(11)
    n = 40
    picks = rng.choice(len(xyzq), n, replace=False)
    newpos = xyzq[picks, :3] + rng.uniform(-1.5, 1.5, (n, 3))
    accept = [(m % 2 == 0) for m in range(n)]
    # reference: one move at a time on context A
    ref = []
    mv = native.FbTrialMove()
    for m in range(n):
        mv.group_index, mv.n_atoms, mv.internal, mv.with_ewald = 0, 1, 1, 1
        mv.rel_index[0], mv.atom_id[0] = int(picks[m]), int(ids[picks[m]])
        for d in range(3):
            mv.xyzq[0][d] = newpos[m, d]
        mv.xyzq[0][3] = xyzq[picks[m], 3]
        out = [C.c_double() for _ in range(4)]
        assert lib.fb_trial_energy(ga.ctx, C.byref(mv), *[C.byref(x) for x in out]) == 0, lib.fb_last_error(ga.ctx)
        ref.append([x.value for x in out])
        assert lib.fb_trial_commit(ga.ctx, int(accept[m])) == 0
    ref = np.array(ref)
    # window on context B
    moves = (native.FbBatchMove * n)()
    for m in range(n):
        moves[m].group_index, moves[m].rel_index, moves[m].atom_id = 0, int(picks[m]), int(ids[picks[m]])
        for d in range(3):
            moves[m].xyzq[d] = newpos[m, d]
        moves[m].xyzq[3] = xyzq[picks[m], 3]
        moves[m].old_atom_id = int(ids[picks[m]])
        for d in range(4):
            moves[m].old_xyzq[d] = xyzq[picks[m], d]
    res = native.FbBatchResult()
    assert lib.fb_batch_trial(gb.ctx, n, moves, 1, C.byref(res)) == 0, lib.fb_last_error(gb.ctx)
    S = res.stride
    assert S == 64 and res.n_moves == n
    arr = lambda p, k: np.ctypeslib.as_array(p, shape=(k,)).copy()
    u_new, u_old, rec = arr(res.u_new, S), arr(res.u_old, S), arr(res.rec_delta, S)
    cn, co, g = (arr(p, S * S).reshape(S, S) for p in (res.cross_new, res.cross_old, res.rec_cross))
    run = res.rec_start
    scale = np.abs(ref[:, :2]).max()
    for m in range(n):
        acc = [a for a in range(m) if accept[a]]
        un = u_new[m] + sum(cn[m, a] for a in acc)   # matrices are stored [m][a]
        uo = u_old[m] + sum(co[m, a] for a in acc)
        dr = rec[m] + 2 * sum(g[m, a] for a in acc)
        assert abs(un - ref[m, 0]) <= RTOL * scale
        assert abs(uo - ref[m, 1]) <= RTOL * scale
        assert abs(res.rec_prefactor * run - ref[m, 3]) <= RTOL * abs(ref[m, 3])
        assert abs(res.rec_prefactor * (run + dr) - ref[m, 2]) <= RTOL * abs(ref[m, 2])
        if accept[m]:
            run += dr
    flags = (C.c_ubyte * n)(*[int(a) for a in accept])
    assert lib.fb_batch_commit(gb.ctx, n, flags) == 0
    # both contexts now hold the same state: full energies agree, and a second window starts from it
    # (the Simulation's host Space was bypassed; compare the device mirrors directly)
    xa = np.zeros((len(xyzq), 4)); xb = np.zeros((len(xyzq), 4))
    ia = np.zeros(len(xyzq), dtype=np.int32); ib = np.zeros(len(xyzq), dtype=np.int32)
    for ctx, x, i in ((ga.ctx, xa, ia), (gb.ctx, xb, ib)):
        assert lib.fb_download_space(ctx, 0, x.ctypes.data_as(native.c_double_p), i.ctypes.data_as(native.c_int_p), None) == 0
    assert np.array_equal(xa, xb) and np.array_equal(ia, ib)
    qa, qb = np.zeros(8000), np.zeros(8000)  # K < 4000 here; Q(k) after the commits: same to rounding
    assert lib.fb_ewald_download(ga.ctx, 0, qa.ctypes.data_as(native.c_double_p), None, None) == 0
    assert lib.fb_ewald_download(gb.ctx, 0, qb.ctypes.data_as(native.c_double_p), None, None) == 0
    assert np.allclose(qa, qb, rtol=0, atol=1e-10 * np.abs(qa).max())


def test_device_walk_equals_host_walk():
    """runs (the device walks the windows, fb_run_submit) against windows walked on the host (fb_batch_submit):
    the same additions in the same order, windows cut differently — identical decisions, energies equal to
    rounding; hard spheres (`pm`: infinite pair
    energies stop windows early) and Ewald"""
    for name in ("pm", "coulombwca_ewald"):
        cfg = dict(ALL_VARIANTS[name])
        host, dev = b200_sim(cfg, 64, run=0), b200_sim(cfg, 64, run=512)
        assert host.run == 0 and dev.run == 512
        for s in (host, dev):
            s.trace_enable()
            s.sweep(600)
        a, b = host.trace(), dev.trace()
        assert len(a["du"]) > 500
        assert np.array_equal(a["accepted"], b["accepted"])
        finite = np.isfinite(a["u_new"])
        assert np.array_equal(finite, np.isfinite(b["u_new"]))
        scale = np.abs(a["u_new"][finite]).max()
        for key in ("u_new", "u_old"):
            assert np.abs(a[key][finite] - b[key][finite]).max() <= 1e-12 * scale, key
        xa, _ = host.particles()
        xb, _ = dev.particles()
        assert np.array_equal(xa, xb)
        assert host.system_energy()[0] == pytest.approx(dev.system_energy()[0], rel=1e-12)
        th, td = host.window_time_ms(), dev.window_time_ms()
        assert td["moves"] == th["moves"] and td["round_trips"] <= th["round_trips"]
        stats = dev.run_stats()
        assert stats["moves"] == td["moves"] and stats["rounds"] >= stats["windows"] > 0


def test_runs_with_pair_sums_ahead(bulk_input):
    """fb_configure_runs(1): pair sums of a window evaluated one window ahead and corrected for the moves accepted
    since (batchPairFixKernel; predictions fail at conditional proposals and early stops and fall back) — same
    trace as the oracle on examples/bulk and on a hard-sphere electrolyte (infinite corrections → redo)"""
    for cfg, sweeps in ((bulk_input, 3), (ALL_VARIANTS["pm"], 600), (ALL_VARIANTS["coulombwca_ewald"], 600)):
        o, g = pair_of_sims(cfg, 64)
        g.configure_runs(True)
        for s in (o, g):
            s.trace_enable()
            s.sweep(sweeps)
        a, b = o.trace(), g.trace()
        assert len(a["du"]) > 500
        assert np.array_equal(a["accepted"], b["accepted"])
        finite = np.isfinite(a["u_new"])
        scale = np.abs(a["u_new"][finite]).max()
        assert_close(a["u_new"], b["u_new"], scale=scale)
        assert_close(a["u_old"], b["u_old"], scale=scale)
        xo, _ = o.particles()
        xg, _ = g.particles()
        assert np.array_equal(xo, xg)
        assert abs(g.drift()) < 1e-9
        assert g.run_stats()["windows"] > 0


@pytest.mark.parametrize("collide", [False, True])
def test_runs_through_the_c_abi(collide):
    """fb_run_submit / fb_run_wait through the raw C ABI: two runs queued behind each other, with conditional
    proposals (a second move on an atom whose first move is still undecided, in the same run and across the two
    runs), against one-move-at-a-time fb_trial_energy / fb_trial_commit with the Metropolis rule applied here.
    collide: move 22 lands 0.01 Å from where the (accepted) move 21 has just put another ion — the correction of its
    energies would cancel 1e28 kT, so the window stops there, the first run needs a window more than scheduled, the
    run queued behind it halts on the device and is launched again."""
    import ctypes as C
    import faunus_b200.native as native
    lib = native.load()
    cfg = small_electrolyte(n=600, coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 9.0, "alpha": 0.35, "ncutoff": 6})
    ga, gb = b200_sim(cfg, 0), b200_sim(cfg, 0)
    xyzq, ids = ga.particles()
    box = np.array(cfg["geometry"]["length"], dtype=float) * np.ones(3)
    rng = np.random.RandomState(5)
    n1, n2 = (128, 90) if collide else (150, 90)
    n = n1 + n2
    atoms = rng.choice(len(xyzq), n, replace=False)
    # repeats: (later move, earlier move) on the same atom — inside run 1 (one of them within one window), inside
    # run 2, and from run 2 back into run 1
    repeats = [(40, 10), (100, 20), (n1 + 30, n1 + 5), (n1 + 80, n1 + 70), (n1 + 10, 120), (n1 + 60, 60)]
    if not collide:
        repeats.append((149, 140))
    for later, earlier in repeats:
        atoms[later] = atoms[earlier]
    disp = rng.uniform(-1.2, 1.2, (n, 3))
    uniform = rng.uniform(size=n)
    wrap = lambda x: x - box * np.round(x / box)
    if collide:
        disp[21] = (1e-3, 0.0, 0.0)   # practically always accepted
        uniform[21] = 0.5
        target = wrap(xyzq[atoms[21], :3] + disp[21]) + np.array([0.01, 0.0, 0.0])
        disp[22] = target - xyzq[atoms[22], :3]

    # reference: one move at a time on context A, Hamiltonian = [non-bonded, Ewald]
    pos = xyzq[:, :3].copy()
    ref_acc, ref_new, ref_old = [], [], []
    mv = native.FbTrialMove()
    for m in range(n):
        a = int(atoms[m])
        new = wrap(pos[a] + disp[m])
        mv.group_index, mv.n_atoms, mv.internal, mv.with_ewald = 0, 1, 1, 1
        mv.rel_index[0], mv.atom_id[0] = a, int(ids[a])
        for d in range(3):
            mv.xyzq[0][d] = new[d]
        mv.xyzq[0][3] = xyzq[a, 3]
        out = [C.c_double() for _ in range(4)]
        assert lib.fb_trial_energy(ga.ctx, C.byref(mv), *[C.byref(x) for x in out]) == 0, lib.fb_last_error(ga.ctx)
        u_new, u_old, ew_new, ew_old = (x.value for x in out)
        t_new, t_old = u_new + ew_new, u_old + ew_old
        accept = bool(uniform[m] <= np.exp(-(t_new - t_old)))
        assert lib.fb_trial_commit(ga.ctx, int(accept)) == 0
        if accept:
            pos[a] = new
        ref_acc.append(accept); ref_new.append(t_new); ref_old.append(t_old)

    # the same proposals as two runs on context B
    def fill(rec, a, start, new):
        rec.group_index, rec.rel_index, rec.atom_id, rec.old_atom_id = 0, a, int(ids[a]), int(ids[a])
        for d in range(3):
            rec.xyzq[d], rec.old_xyzq[d] = new[d], start[d]
        rec.xyzq[3] = rec.old_xyzq[3] = xyzq[a, 3]

    def pack(first, count):
        block = (native.FbRunMove * count)()
        for i in range(count):
            m = first + i
            a = int(atoms[m])
            earlier = [k for k in range(m) if atoms[k] == a]
            r = block[i]
            r.uniform, r.host_new, r.host_old, r.flags = uniform[m], 0.0, 0.0, 0
            start = xyzq[a, :3]
            if not earlier:
                r.depends_on = -1
                fill(r.move, a, start, wrap(start + disp[m]))
            else:
                (q,) = earlier
                r.depends_on = q - first if q >= first else (1 << 30) | q
                q_new = wrap(start + disp[q])
                fill(r.move, a, q_new, wrap(q_new + disp[m]))   # the earlier move is accepted
                fill(r.alt, a, start, wrap(start + disp[m]))     # ... rejected
        return block

    cfg_run = native.FbRunConfig(max_energy=float("inf"), cancellation_limit=1e4)
    b1, b2 = pack(0, n1), pack(n1, n2)
    assert lib.fb_run_submit(gb.ctx, n1, b1, 1, C.byref(cfg_run)) == 0, lib.fb_last_error(gb.ctx)
    assert lib.fb_run_submit(gb.ctx, n2, b2, 1, C.byref(cfg_run)) == 0, lib.fb_last_error(gb.ctx)   # queued behind
    assert lib.fb_run_submit(gb.ctx, n2, b2, 1, C.byref(cfg_run)) != 0                               # two at most
    got_acc, got_new, got_old, windows = [], [], [], []
    for count in (n1, n2):
        res = native.FbRunResult()
        assert lib.fb_run_wait(gb.ctx, C.byref(res)) == 0, lib.fb_last_error(gb.ctx)
        assert res.n_moves == count
        got_acc += [bool(res.accepted[i]) for i in range(count)]
        got_new += [res.u_new[i] for i in range(count)]
        got_old += [res.u_old[i] for i in range(count)]
        windows.append(res.n_windows)
    if collide:   # 128 moves: 22 (stopped at the collision) + 64 + 42, one more than the two scheduled
        assert windows[0] == 3 and ref_acc[21] and not ref_acc[22]
    else:         # 150 moves: 64 + 64 + 21 (cut before move 149, which depends on 140) + 1
        assert windows[0] == 4 and windows[1] >= 2
    assert got_acc == ref_acc
    assert 0.2 < np.mean(ref_acc) < 0.9
    ordinary = np.abs(ref_new) < 1e6   # the collision itself: 1e28 kT, compared relatively below
    scale = np.abs(np.array(ref_new)[ordinary]).max()
    assert np.abs(np.array(got_new) - ref_new)[ordinary].max() <= RTOL * scale
    assert np.abs(np.array(got_old) - ref_old)[ordinary].max() <= RTOL * scale
    assert np.allclose(np.array(got_new)[~ordinary], np.array(ref_new)[~ordinary], rtol=1e-9)
    # both contexts end in the same state (fb_download_space applies what is still pending)
    xa, xb = np.zeros((len(xyzq), 4)), np.zeros((len(xyzq), 4))
    ia, ib = np.zeros(len(xyzq), dtype=np.int32), np.zeros(len(xyzq), dtype=np.int32)
    for ctx, x, i in ((ga.ctx, xa, ia), (gb.ctx, xb, ib)):
        assert lib.fb_download_space(ctx, 0, x.ctypes.data_as(native.c_double_p), i.ctypes.data_as(native.c_int_p), None) == 0
    assert np.array_equal(xa, xb)
    assert np.array_equal(xa[:, :3], pos)


def test_system_energy_shards_add_up():
    """fb_system_energy_shard: tile rows / k-vector slabs dealt to 3 'GPUs' add up to the full energies"""
    cfg = small_electrolyte(n=900, coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 9.0, "alpha": 0.35, "ncutoff": 6})
    o, g = pair_of_sims(cfg)
    _, terms = o.system_energy()   # [self, nonbonded, ewald (tinfoil: reciprocal only)]
    for size in (1, 3):
        parts = np.array([g.system_energy_shard(r, size) for r in range(size)])
        assert_close([parts[:, 0].sum()], [terms[1]], scale=np.abs(terms).max())
        assert_close([parts[:, 1].sum()], [terms[2]], scale=np.abs(terms).max())
    assert abs(g.system_energy_shard(1, 3)[0]) > 0


def test_system_energy_shards_with_particle_splits():
    """a slab of few k-cells also splits the particles (shares of 8192, ewaldFullCellKernel + ewaldCellEnergyKernel):
    N = 20 000 over 1, 2 and 8 'GPUs' against the oracle"""
    cfg = small_electrolyte(n=20000, coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 9.0, "alpha": 0.35, "ncutoff": 9})
    from _oraclelib import oracle_lib
    oracle_lib().fo_set_parallel_ewald_init(1)
    try:
        o, g = pair_of_sims(cfg)
        _, terms = o.system_energy()
    finally:
        oracle_lib().fo_set_parallel_ewald_init(0)
    for size in (1, 2, 8):
        parts = np.array([g.system_energy_shard(r, size) for r in range(size)])
        assert_close([parts[:, 0].sum()], [terms[1]], scale=np.abs(terms).max())
        assert_close([parts[:, 1].sum()], [terms[2]], scale=np.abs(terms).max())


def test_widom_sharded_matches_unsharded():
    """Widom insertions split over two 'ranks' (two contexts on this GPU) == the unsharded batched run"""
    import ctypes as C
    cfg = small_electrolyte(coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 9.0, "alpha": 0.35, "ncutoff": 6},
                            ghost_pairs=1)
    analysis = {"molecule": "ghost", "ninsert": 101}
    ref, a, b = b200_sim(cfg), b200_sim(cfg), b200_sim(cfg)
    wr, wa, wb = (s.widom_create(analysis) for s in (ref, a, b))
    ref.widom_sample(wr, 1)
    n = a.api.widom_prepare(a.handle, wa)
    assert n == b.api.widom_prepare(b.handle, wb) == 101
    bounds = [0, 50, 101]
    parts = []
    for rank, (s, w) in enumerate(((a, wa), (b, wb))):
        local = np.zeros(bounds[rank + 1] - bounds[rank])
        assert s.api.widom_evaluate_slice(s.handle, w, bounds[rank], len(local),
                                          local.ctypes.data_as(C.POINTER(C.c_double))) == 0
        parts.append(local)
    everyone = np.concatenate(parts)
    for s, w in ((a, wa), (b, wb)):
        assert s.api.widom_collect(s.handle, w, everyone.ctypes.data_as(C.POINTER(C.c_double)), n) == 0
    rr, ra = ref.widom_result(wr), a.widom_result(wa)
    assert np.array_equal(rr["last_du"], ra["last_du"])
    assert rr["sum_exp"] == ra["sum_exp"] == b.widom_result(wb)["sum_exp"]
    # size 1 through the public call
    c = b200_sim(cfg)
    wc = c.widom_create(analysis)
    assert c.widom_sample_sharded(wc, 0, 1) == 101
    assert c.widom_result(wc)["sum_exp"] == rr["sum_exp"]


def test_tempering_replicas_on_device():
    """Hamiltonian parallel tempering (examples/temper semantics, src/move.cpp:844-968) with the replicas'
    energies on the B200 terms: same exchanges and final states as the oracle replicas"""
    import faunus_b200.native as native
    from faunus_b200.replica import run_local_replicas
    from _oraclelib import oracle_api
    from test_tempering_cpu import electrolyte_replicas
    cfgs = electrolyte_replicas(2)
    want = run_local_replicas(oracle_api(), cfgs, sweeps=20)
    got = run_local_replicas(native.sim_library(), cfgs, sweeps=20)
    assert all(r["error"] == "" for r in got)
    for w, g in zip(want, got):
        assert np.array_equal(np.array(w["xyzq"]), np.array(g["xyzq"]))   # identical trajectories
        assert g["energy"] == pytest.approx(w["energy"], rel=1e-10)
        assert abs(g["drift"]) < 1e-9
        tw = [m["temper"] for m in w["moves"] if "temper" in m][0]
        tg = [m["temper"] for m in g["moves"] if "temper" in m][0]
        assert tw["exchange"] == tg["exchange"]


@pytest.mark.parametrize("window", [32, 64])
@pytest.mark.parametrize("capacity", [None, 2])
def test_window_cell_list(capacity, window):
    """pair part of the windows through the device cell list (forced on a small system): same trace as the
    oracle's brute-force sums; capacity 2 makes buckets run full, which must fall back and grow (window 64:
    inside a device-decided run)"""
    import faunus_b200.native as native
    cfg = small_electrolyte(n=500, moves_per_sweep=200,
                            coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 9.0, "alpha": 0.35, "ncutoff": 6})
    o, g = pair_of_sims(cfg, window)
    g.configure_cells(0)
    if capacity:
        assert native.load().fb_debug_set_cell_capacity(g.ctx, capacity) == 0
    launches0 = g.launch_count
    for s in (o, g):
        s.trace_enable()
        s.sweep(3)
    a, b = o.trace(), g.trace()
    assert np.array_equal(a["accepted"], b["accepted"])
    scale = np.abs(a["u_new"][np.isfinite(a["u_new"])]).max()
    assert_close(a["u_new"], b["u_new"], scale=scale)
    assert_close(a["u_old"], b["u_old"], scale=scale)
    assert abs(g.drift()) < 1e-9
    assert g.launch_count > launches0


@pytest.mark.parametrize("window", [16, 64])
@pytest.mark.parametrize("cells", [0, -1])
def test_window_with_volume_moves(window, cells):
    """NPT electrolyte: runs of windowed `transrot` moves interleaved with `volume` moves (everything changes:
    box, k-vectors, Q(k), cell list) — the queue is drained, the other move runs one at a time, windows resume.
    cells = −1: brute-force windows, whose launches are replayed as CUDA graphs — every accepted volume move changes
    the launch arguments (box lengths), so the graphs have to be captured anew"""
    cfg = small_electrolyte(n=300, moves_per_sweep=60,
                            coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 9.0, "alpha": 0.35, "ncutoff": 5})
    cfg["energy"] = [{"isobaric": {"P/mM": 2000.0}}] + cfg["energy"]
    cfg["moves"].append({"volume": {"dV": 0.03, "repeat": 6}})
    o = oracle_sim(cfg)
    g = b200_sim(cfg, window)
    assert g.window == window   # the isobaric term is a per-atom host term: windows stay eligible
    g.configure_cells(cells)
    for s in (o, g):
        s.trace_enable()
        s.sweep(6)
    a, b = o.trace(), g.trace()
    assert (a["move_id"] == 1).sum() > 5, "no volume moves sampled"
    assert np.array_equal(a["move_id"], b["move_id"])
    assert np.array_equal(a["accepted"], b["accepted"])
    scale = np.abs(o.system_energy()[1]).max()
    assert_close(a["du"], b["du"], rtol=1e-9, scale=scale)   # volume moves: differences of full energies
    assert_close(o.system_energy()[1], g.system_energy()[1], scale=scale)
    xo, _ = o.particles()
    xg, _ = g.particles()
    assert np.array_equal(xo, xg)


@pytest.mark.parametrize("case", ["bulk_na_cl", "bulk_na_na", "bulk_slice", "water_ow_hw", "sphere_pm"])
def test_atom_rdf_counts_are_exact(bulk_input, water_input, case):
    """fb_atom_rdf (pair histogram on the device mirror) against the oracle's AtomRDF pair loop: the same integer
    counts in every bin, accumulated over samples taken between sweeps (the mirror follows windows, runs, volume
    moves); periodic cuboid, NPT water (box changes), and a non-periodic cell"""
    if case.startswith("bulk"):
        cfg = bulk_input
        rdf = {"bulk_na_cl": {"name1": "Na", "name2": "Cl", "dr": 0.1},
               "bulk_na_na": {"name1": "Na", "name2": "Na", "dr": 0.05},
               "bulk_slice": {"name1": "Cl", "name2": "Na", "dr": 0.2, "slicedir": [1, 0, 0], "thickness": 4.0}}[case]
    elif case == "water_ow_hw":
        cfg, rdf = water_input, {"name1": "OW", "name2": "HW", "dr": 0.1}
    else:
        cfg = small_electrolyte(n=300, energy_name="nonbonded_pm", coulomb={"epsr": 78.7}, sigma=3.0)
        cfg["geometry"] = {"type": "sphere", "radius": 60.0}   # holds the corners of the 63 Å cube the ions start in
        rdf = {"name1": "Na", "name2": "Cl", "dr": 0.5}
    o, g = pair_of_sims(cfg, 64)
    ro, rg = o.rdf_create(dict(rdf, file="rdf.dat")), g.rdf_create(dict(rdf, file="rdf.dat"))
    for _ in range(3):
        for s, r in ((o, ro), (g, rg)):
            s.sweep(1)
            s.rdf_sample(r)
    (r_o, pairs_o, g_o), (r_g, pairs_g, g_g) = o.rdf_result(ro), g.rdf_result(rg)
    n = min(len(pairs_o), len(pairs_g))   # the device histogram is sized for the cell, the oracle's grows on demand
    assert pairs_o[n:].sum() == 0 and pairs_g[n:].sum() == 0
    assert pairs_o.sum() > 1000
    assert np.array_equal(pairs_o[:n], pairs_g[:n])
    assert np.array_equal(r_o[:n], r_g[:n])
    assert np.allclose(g_o[:n], g_g[:n], rtol=1e-12, atol=0)


def test_atom_rdf_shards_add_up(bulk_input):
    """fb_atom_rdf split over three 'GPUs' (tile rows dealt round robin): the integer histograms add up exactly"""
    g = b200_sim(bulk_input, 64)
    g.sweep(1)
    cfg = {"name1": "Na", "name2": "Na", "dr": 0.1, "file": "rdf.dat"}
    whole = g.rdf_create(cfg)
    g.rdf_sample(whole)
    parts = []
    for rank in range(3):
        rid = g.rdf_create(cfg)
        g.rdf_sample_shard(rid, rank, 3)
        parts.append(g.rdf_result(rid)[1])
    assert all(p.sum() > 0 for p in parts)
    assert np.array_equal(sum(parts), g.rdf_result(whole)[1])


@pytest.mark.parametrize("case", ["bulk", "ewald", "water"])
def test_virtual_volume_move(bulk_input, water_input, case):
    """VirtualVolumeMove (src/analysis.cpp:825-843): energy(change = everything + volume) on the accepted
    Hamiltonian before and after scaling the Space, no updateState, no sync — the adaptor terms refresh their mirror
    (and, like the reference, the Ewald term keeps the k-vectors of the unscaled cell); then MC goes on as before"""
    cfg = {"bulk": bulk_input, "water": water_input,
           "ewald": small_electrolyte(n=400, coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 9.0, "alpha": 0.35,
                                                      "ncutoff": 6})}[case]
    o, g = pair_of_sims(cfg, 64)
    vo, vg = o.virtualvolume_create({"dV": 20.0}), g.virtualvolume_create({"dV": 20.0})
    scale = np.abs(o.system_energy()[1]).max()
    for _ in range(2):
        for s, v in ((o, vo), (g, vg)):
            s.sweep(1)
            s.virtualvolume_sample(v)
    ro, rg = o.virtualvolume_result(vo), g.virtualvolume_result(vg)
    assert ro["count"] == rg["count"]
    assert abs(ro["last_du"] - rg["last_du"]) <= 1e-9 * scale   # a difference of two full energies
    if ro["count"]:
        assert rg["excess_pressure_kT_per_A3"] == pytest.approx(ro["excess_pressure_kT_per_A3"], rel=1e-6, abs=1e-9 * scale)
    for s in (o, g):   # the analysis left both sides where they were
        s.trace_enable()
        s.sweep(1)
    assert np.array_equal(o.trace()["accepted"], g.trace()["accepted"])
    assert_close(o.system_energy()[1], g.system_energy()[1], scale=scale)


@pytest.mark.parametrize("coulomb", [None, {"type": "fanourgakis", "epsr": 1, "cutoff": 9}])
def test_virtual_translate(water_input, coulomb):
    """VirtualTranslate (src/analysis.cpp:2794-2860) between sweeps: a group moved, evaluated and moved back
    without updateState / sync — same ΔU as the oracle, and the mirror is back in step afterwards (identical
    traces); with and without Ewald (whose Q(k), like the reference's, does not follow the virtual move)"""
    from conftest import one_water_in_salt
    cfg = one_water_in_salt(water_input, coulomb=coulomb)
    o, g = pair_of_sims(cfg, 64)
    scale = np.abs(o.system_energy()[1]).max()
    vo, vg = (s.virtualtranslate_create({"molecule": "water", "dL": 0.25, "dir": [0, 0, 1]}) for s in (o, g))
    for s in (o, g):
        s.trace_enable()
    for _ in range(3):
        for s, v in ((o, vo), (g, vg)):
            s.sweep(1)
            s.virtualtranslate_sample(v)
    ro, rg = o.virtualtranslate_result(vo), g.virtualtranslate_result(vg)
    assert ro["count"] == rg["count"] == 3
    assert abs(ro["last_du"] - rg["last_du"]) <= 1e-10 * scale
    assert rg["sum_exp"] == pytest.approx(ro["sum_exp"], rel=1e-9)
    a, b = o.trace(), g.trace()
    assert np.array_equal(a["move_id"], b["move_id"]) and np.array_equal(a["accepted"], b["accepted"])
    assert_close(o.system_energy()[1], g.system_energy()[1], scale=scale)
    xo, _ = o.particles()
    xg, _ = g.particles()
    assert np.array_equal(xo, xg)


def test_molecule_rdf_counts_are_exact(water_input):
    """fb_molecule_rdf (mass centres from the mirror, which follows rigid-molecule windows and volume moves) against
    the oracle's MoleculeRDF loop: equal counts; sharded in three, the histograms add up"""
    cfg = {"type": "molrdf", "name1": "water", "name2": "water", "dr": 0.1, "file": "rdf.dat"}
    o, g = pair_of_sims(water_input, 64)
    ro, rg = o.rdf_create(cfg), g.rdf_create(cfg)
    for _ in range(3):
        for s, r in ((o, ro), (g, rg)):
            s.sweep(1)
            s.rdf_sample(r)
    pairs_o, pairs_g = o.rdf_result(ro)[1], g.rdf_result(rg)[1]
    n = min(len(pairs_o), len(pairs_g))
    assert pairs_o[n:].sum() == 0 and pairs_g[n:].sum() == 0 and pairs_o.sum() == 3 * 256 * 255 // 2
    assert np.array_equal(pairs_o[:n], pairs_g[:n])
    whole = g.rdf_create(cfg)
    g.rdf_sample(whole)
    parts = []
    for rank in range(3):
        rid = g.rdf_create(cfg)
        g.rdf_sample_shard(rid, rank, 3)
        parts.append(g.rdf_result(rid)[1])
    assert np.array_equal(sum(parts), g.rdf_result(whole)[1])


def test_restore_on_the_device(water_input):
    """`restore` (src/montecarlo.cpp:118-137) of a state file written by the ORACLE after two sweeps — positions,
    group records and both generators — into a fresh device simulation: the mirror is re-uploaded, Q(k) rebuilt, and
    the continued run equals the oracle's own continuation (windows and one move per launch)."""
    from conftest import water_with_salt
    cfg = water_with_salt(water_input, n_pairs=6)
    o = oracle_sim(cfg)
    o.sweep(2)
    state = o.state_json()
    o.trace_enable()
    o.sweep(2)
    ref = o.trace()
    for window in (0, 64):
        g = b200_sim(cfg, window)
        g.sweep(1)  # the device has moved on before the restore
        g.restore(state)
        g.trace_enable()
        g.sweep(2)
        got = g.trace()
        assert np.array_equal(ref["move_id"], got["move_id"]) and np.array_equal(ref["accepted"], got["accepted"])
        scale = np.abs(ref["u_new"]).max()
        assert np.abs(ref["u_new"] - got["u_new"]).max() <= 1e-10 * scale
        assert np.array_equal(o.particles()[0], g.particles()[0])
        assert abs(g.drift()) < 1e-9


def test_restore_on_the_device_from_a_binary_state_file(water_input, tmp_path):
    """the same through files: the oracle writes `state.ubj` (Universal Binary JSON, src/analysis.cpp:656-668), the device
    simulation starts from it (`--state state.ubj`, src/faunus.cpp:430-455) and writes one itself that the oracle reads"""
    from conftest import water_with_salt
    cfg = water_with_salt(water_input, n_pairs=6)
    o = oracle_sim(cfg)
    o.sweep(2)
    first = str(tmp_path / "oracle.ubj")
    o.save_state(first)
    g = b200_sim(cfg, 64)
    g.load_state(first)
    for s in (o, g):
        s.trace_enable()
        s.sweep(1)
    assert np.array_equal(o.trace()["accepted"], g.trace()["accepted"])
    second = str(tmp_path / "device.ubj")
    g.save_state(second)
    o2 = oracle_sim(cfg)
    o2.load_state(second)
    for s in (o2, g):
        s.trace_enable()
        s.sweep(1)
    assert np.array_equal(o2.trace()["accepted"][-len(g.trace()["accepted"]):], g.trace()["accepted"]) or \
        np.array_equal(o2.trace()["accepted"], g.trace()["accepted"][-len(o2.trace()["accepted"]):])
    assert np.array_equal(o2.particles()[0], g.particles()[0])


def test_packed_state_round_trip(water_input):
    """fb_export_state / fb_import_state (+ fb_upload_groups): the packed mirror of one simulation imported into the
    mirror of another gives the energy of the original (src/mpicontroller.cpp:192-219: what a replica exchange
    ships). Molecular groups (mass centres follow by fb_upload_groups) and a box that changed (NPT)."""
    import ctypes as C
    from conftest import water_with_salt
    from faunus_b200.native import FbGroup, load, make_change
    lib = load()
    dp = C.POINTER(C.c_double)
    cfg = water_with_salt(water_input, n_pairs=6, volume_move=True)
    a, b = b200_sim(cfg, 0), b200_sim(cfg, 0)
    a.sweep(3)  # a has moved on (its box too), b still holds the start configuration
    n_groups = a.api.sim_num_groups(a.handle)
    n = lib.fb_state_doubles(a.ctx)
    assert n == 3 + n_groups + 5 * a.num_particles
    buf = np.zeros(n)
    assert lib.fb_export_state_host(a.ctx, 0, buf.ctypes.data_as(dp)) == 0, lib.fb_last_error(a.ctx)
    xyzq, ids = a.particles()
    packed = buf[3 + n_groups:].reshape(-1, 5)
    assert np.array_equal(packed[:, :4], xyzq) and np.array_equal(packed[:, 4], ids)
    rec, cm = a.groups()
    assert np.array_equal(buf[3:3 + n_groups], rec[:, 1])
    groups = (FbGroup * n_groups)()
    for g in range(n_groups):
        groups[g].begin, groups[g].size, groups[g].capacity, groups[g].molid = (int(v) for v in rec[g])
        for k in range(3):
            groups[g].cm[k] = cm[g, k]
    everything = make_change(everything=True)
    want = C.c_double()
    assert lib.fb_nonbonded_energy(a.ctx, 0, C.byref(everything), C.byref(want)) == 0, lib.fb_last_error(a.ctx)
    before = C.c_double()
    assert lib.fb_nonbonded_energy(b.ctx, 0, C.byref(everything), C.byref(before)) == 0
    assert abs(before.value - want.value) > 1e-6 * abs(want.value)  # b really is somewhere else
    for slot in (0, 1):
        assert lib.fb_import_state_host(b.ctx, slot, buf.ctypes.data_as(dp)) == 0, lib.fb_last_error(b.ctx)
        assert lib.fb_upload_groups(b.ctx, slot, groups, n_groups) == 0, lib.fb_last_error(b.ctx)
        got = C.c_double()
        assert lib.fb_nonbonded_energy(b.ctx, slot, C.byref(everything), C.byref(got)) == 0, lib.fb_last_error(b.ctx)
        assert abs(got.value - want.value) <= 1e-12 * abs(want.value)
    # and back out again: the packed state of b is that of a, bit for bit
    again = np.zeros(n)
    assert lib.fb_export_state_host(b.ctx, 1, again.ctypes.data_as(dp)) == 0
    assert np.array_equal(again, buf)


def test_tempering_packed_exchange_equals_message_exchange():
    """In-process replicas on the device: the exchange of the packed mirrors (fb_export_state → fb_import_state, the
    Spaces follow; the host logic of the NCCL communicator) reproduces the exchange of the reference's three
    messages bit for bit, and both reproduce the oracle's in-process run."""
    from faunus_b200.config import primitive_model
    from faunus_b200.native import sim_library
    from faunus_b200.replica import run_local_replicas
    from _oraclelib import oracle_api
    cfgs = []
    for r in range(3):
        cfg = primitive_model(n=300, seed=11, moves_per_sweep=30,
                              coulomb={"type": "ewald", "epsr": 60.0 + 12.0 * r, "cutoff": 10.0, "alpha": 0.3, "ncutoff": 5})
        cfg["moves"].append({"temper": {"format": "xyzqi"}})
        cfgs.append(cfg)
    sweeps = 25
    plain = run_local_replicas(sim_library(), cfgs, sweeps)
    packed = run_local_replicas(sim_library(), cfgs, sweeps, packed=True)
    oracle = run_local_replicas(oracle_api(), cfgs, sweeps)
    accepted = 0.0
    for a, b, o in zip(plain, packed, oracle):
        assert a["error"] == b["error"] == o["error"] == ""
        ta = [m["temper"] for m in a["moves"] if "temper" in m][0]["exchange"]
        tb = [m["temper"] for m in b["moves"] if "temper" in m][0]["exchange"]
        to = [m["temper"] for m in o["moves"] if "temper" in m][0]["exchange"]
        assert ta == tb == to
        assert a["xyzq"] == b["xyzq"] == o["xyzq"]
        assert a["energy"] == b["energy"]
        assert abs(b["drift"]) < 1e-9
        accepted += sum(s["acceptance"] * s["attempts"] for s in tb.values())
    assert accepted > 0


@pytest.mark.parametrize("coulomb", [{"type": "fanourgakis", "epsr": 78.7, "cutoff": 10.0},
                                     {"type": "ewald", "epsr": 78.7, "cutoff": 9.0, "alpha": 0.35, "ncutoff": 5}])
@pytest.mark.parametrize("window", [0, 64])
def test_matter_change_atomic(coulomb, window):
    """Change::matter_change on the device (GroupPairing::accumulateSpeciation, src/energy.h:1390-1435; Ewald partial
    update with particles that appear / disappear, src/energy.cpp:219-247; translational-entropy bias,
    src/montecarlo.cpp:271-374): ions of an atomic group are activated and deactivated between sweeps of ordinary
    moves; energies of both states equal the oracle's, and so do the traces of the sweeps in between."""
    cfg = small_electrolyte(n=150, ghost_pairs=3, moves_per_sweep=60, coulomb=coulomb)
    o, g = pair_of_sims(cfg, window)
    rec, _ = o.groups()
    ghost = int(np.argmax(rec[:, 2] - rec[:, 1]))
    pos = [[3.0, -7.0, 11.0], [-9.0, 2.5, -4.0], [12.0, 12.0, -12.0], [1.0, 1.0, 1.5]]
    steps = [
        ([{"index": ghost, "size": 2, "atoms": [0, 1], "pos": pos[:2], "dNatomic": True}], 1),
        ([{"index": ghost, "size": 4, "atoms": [2, 3], "pos": pos[2:], "dNatomic": True}], 0),   # rejected
        ([{"index": ghost, "size": 3, "atoms": [2], "pos": pos[2:3], "dNatomic": True}], 1),
        ([{"index": ghost, "size": 1, "atoms": [1, 2], "dNatomic": True}], 1),                   # two removed
        ([{"index": ghost, "size": 0, "atoms": [0], "dNatomic": True}], 2),                       # Metropolis
    ]
    for s in (o, g):
        s.trace_enable()
    scale = np.abs(o.system_energy()[1]).max()
    for groups, mode in steps:
        ro, rg = o.matter_change(groups, mode), g.matter_change(groups, mode)
        assert ro["accepted"] == rg["accepted"]
        assert abs(ro["u_new"] - rg["u_new"]) <= 1e-10 * scale and abs(ro["u_old"] - rg["u_old"]) <= 1e-10 * scale
        assert ro["bias"] == rg["bias"]
        assert_close(o.system_energy()[1], g.system_energy()[1], scale=scale)
        for s in (o, g):
            s.sweep(1)
        assert np.array_equal(o.trace()["accepted"], g.trace()["accepted"])
    assert_close(o.trace()["du"], g.trace()["du"], scale=scale)
    assert np.array_equal(o.particles()[0], g.particles()[0])
    assert abs(g.drift()) < 1e-9


def test_matter_change_molecule(water_input):
    """a whole rigid molecule disappears and comes back (`all`, every atom listed), next to an atomic salt group:
    group-group pairs with the mass-centre cutoff, no internal energy for rigid bodies, k-space follows"""
    from conftest import water_with_salt
    cfg = water_with_salt(water_input, n_pairs=6)
    o, g = pair_of_sims(cfg, 64)
    for s in (o, g):
        s.trace_enable()
        s.sweep(1)
    scale = np.abs(o.system_energy()[1]).max()
    molecule = 17
    for size, mode in ((0, 1), (3, 1), (0, 0)):
        groups = [{"index": molecule, "size": size, "atoms": [0, 1, 2], "all": True}]
        ro, rg = o.matter_change(groups, mode), g.matter_change(groups, mode)
        assert ro["accepted"] == rg["accepted"]
        assert abs(ro["u_new"] - rg["u_new"]) <= 1e-10 * scale and abs(ro["u_old"] - rg["u_old"]) <= 1e-10 * scale
        assert ro["bias"] == rg["bias"]
        assert_close(o.system_energy()[1], g.system_energy()[1], scale=scale)
        for s in (o, g):
            s.sweep(1)
        assert np.array_equal(o.trace()["accepted"], g.trace()["accepted"])
    assert np.array_equal(o.particles()[0], g.particles()[0])


def test_particle_and_group_pair_energies(water_input):
    """NonbondedBase::particleParticleEnergy / groupGroupEnergy (src/energy.h:1498-1503) through the C ABI against the
    oracle's pair functor: explicit particle pairs, and pairs of rigid molecules with the mass-centre cutoff."""
    import ctypes as C
    from _oraclelib import pair_energy
    from faunus_b200.native import load
    lib = load()
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    g = b200_sim(water_input, 0)
    xyzq, ids = g.particles()
    rec, cm = g.groups()
    box = np.array(water_input["geometry"]["length"], dtype=float)
    name = [k for e in water_input["energy"] for k in e if k.startswith("nonbonded")][0]

    def r_min(a, b):
        d = np.abs(a - b)
        d -= box * (d > box / 2)
        return np.sqrt((d * d).sum())

    # explicit pairs: the first 40 atoms against atoms further down the list
    n = 40
    a, b = np.ascontiguousarray(xyzq[:n]), np.ascontiguousarray(xyzq[100:100 + n])
    ia, ib = np.ascontiguousarray(ids[:n]), np.ascontiguousarray(ids[100:100 + n])
    out = np.zeros(n)
    assert lib.fb_particle_pair_energy(g.ctx, 0, n, a.ctypes.data_as(dp), ia.ctypes.data_as(ip), b.ctypes.data_as(dp),
                                       ib.ctypes.data_as(ip), out.ctypes.data_as(dp)) == 0, lib.fb_last_error(g.ctx)
    want = np.array([pair_energy(water_input, name, int(ia[k]), int(ib[k]), [r_min(a[k, :3], b[k, :3])])[0] for k in range(n)])
    assert np.abs(out - want).max() <= 1e-10 * np.abs(want).max()
    # molecule pairs: inside and outside the mass-centre cutoff
    cut = 10.0
    seen = set()
    for g1 in range(0, 12):
        for g2 in range(g1 + 1, 40):
            beyond = r_min(cm[g1], cm[g2]) >= cut
            if (beyond in seen) and len(seen) == 2 and g2 > g1 + 3:
                continue
            seen.add(beyond)
            u = C.c_double()
            assert lib.fb_group_group_energy(g.ctx, 0, g1, g2, C.byref(u)) == 0, lib.fb_last_error(g.ctx)
            expected = 0.0
            if not beyond:
                for i in range(rec[g1][0], rec[g1][0] + rec[g1][1]):
                    for j in range(rec[g2][0], rec[g2][0] + rec[g2][1]):
                        expected += pair_energy(water_input, name, int(ids[i]), int(ids[j]), [r_min(xyzq[i, :3], xyzq[j, :3])])[0]
            assert abs(u.value - expected) <= 1e-10 * max(1.0, abs(expected))
    assert seen == {True, False}


FORCE_VARIANTS = ["coulombwca_ewald_surface", "coulombwca_ewald", "coulomblj_fanourgakis", "coulombwca_yukawa", "coulombwca_qpot",
                  "coulombwca_wolf", "coulombwca_zerodipole", "coulombwca_reactionfield", "coulombwca_poisson"]


@pytest.mark.parametrize("name", FORCE_VARIANTS)
def test_forces_match_oracle(name):
    """Nonbonded::force (src/energy.h:1584-1597), Ewald::force (src/energy.cpp:596-629) and Hamiltonian::force
    (src/energy.cpp:1162-1166) on the device against the oracle, term by term, before and after moves; 1e-10 of the
    largest force component of the term."""
    o, g = pair_of_sims(electrolyte_variants()[name], 64)
    n_terms = len(o.system_energy()[1])
    for sweep in range(2):
        for term in range(n_terms):
            want, got = o.forces(term=term), g.forces(term=term)
            assert_close(want, got, scale=max(np.abs(want).max(), 1e-300))
        want, got = o.forces(), g.forces()
        assert np.abs(want).max() > 0
        assert_close(want, got, scale=np.abs(want).max())
        for s in (o, g):
            s.sweep(1)
        assert np.array_equal(o.trace()["accepted"], g.trace()["accepted"])
    # the trial state's terms see the trial Space
    assert_close(o.forces(which=1), g.forces(which=1), scale=np.abs(o.forces(which=1)).max())


def test_forces_of_the_bulk_example(bulk_input):
    """examples/bulk (N = 2304, fanourgakis + Lennard-Jones, no Ewald term): the pair forces of all 2.65e6 pairs"""
    o, g = pair_of_sims(bulk_input, 64)
    want, got = o.forces(), g.forces()
    assert np.abs(want.sum(axis=0)).max() <= 1e-9 * np.abs(want).max()  # Newton's third law
    assert_close(want, got, scale=np.abs(want).max())


@pytest.mark.parametrize("name", ["pm", "pmwca", "functor", "splined"])
def test_no_forces_where_the_reference_has_none(name):
    """PairPotential::force throws for plain Coulomb, hard spheres and the functor potentials
    (src/potentials.cpp:246-251): same message from the oracle and from the device library"""
    o, g = pair_of_sims(ALL_VARIANTS[name], 0)
    for s in (o, g):
        with pytest.raises(RuntimeError, match="Force computation not implemented"):
            s.forces()


def test_forces_include_inactive_particles():
    """the reference's stub sums over the whole particle vector, inactive ghosts included (src/energy.h:1590-1596)"""
    cfg = small_electrolyte(n=60, coulomb={"type": "fanourgakis", "epsr": 78.7, "cutoff": 9.0}, ghost_pairs=1)
    cfg["particles"][-2]["pos"] = [1.0, 2.0, 3.0]  # the inactive pair: somewhere, apart from each other
    cfg["particles"][-1]["pos"] = [-4.0, 2.5, 0.5]
    o, g = pair_of_sims(cfg, 0)
    want, got = o.forces(), g.forces()
    assert want.shape[0] == 62 and np.abs(want[-2:]).max() > 0 and np.all(np.isfinite(want))
    assert_close(want, got, scale=np.abs(want).max())


@pytest.mark.parametrize("geometry", [{"type": "sphere", "radius": 60.0}, {"type": "slit", "length": [63.0, 63.0, 80.0]}])
def test_forces_without_full_periodicity(geometry):
    """`Chameleon::vdist` folds periodic axes only (src/geometry.h:429-458): a sphere (no fold) and a slit (x, y)"""
    cfg = small_electrolyte(n=300, energy_name="nonbonded_coulomblj", coulomb={"type": "yukawa", "epsr": 78.7, "debyelength": 12.0})
    cfg["geometry"] = geometry
    o, g = pair_of_sims(cfg, 0)
    terms = len(o.system_energy()[1])
    nonbonded = [t for t in range(terms) if np.abs(o.forces(term=t)).max() > 0]
    assert len(nonbonded) == 1
    want, got = o.forces(term=nonbonded[0]), g.forces(term=nonbonded[0])
    assert_close(want, got, scale=np.abs(want).max())
