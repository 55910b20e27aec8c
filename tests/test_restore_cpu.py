"""State files: `savestate` with the generators (src/analysis.cpp:656-682) → `restore` (src/montecarlo.cpp:118-137)
continues the SAME proposal stream: a restarted run reproduces the uninterrupted one move by move."""
import numpy as np

from _oraclelib import oracle_sim
from conftest import small_electrolyte


def _continue_equals_restart(make, sweeps_before=2, sweeps_after=2):
    a = make()
    a.sweep(sweeps_before)
    state = a.state_json()
    assert "random-move" in state and "random-global" in state and "particles" in state
    a.trace_enable()
    a.sweep(sweeps_after)
    b = make()
    b.restore(state)
    b.trace_enable()
    b.sweep(sweeps_after)
    ta, tb = a.trace(), b.trace()
    assert len(ta["du"]) == len(tb["du"]) > 0
    assert np.array_equal(ta["accepted"], tb["accepted"])
    # Q(k) of the restarted run is rebuilt from the positions, that of the uninterrupted run updated incrementally
    scale = max(np.abs(ta["u_new"]).max(), np.abs(ta["u_old"]).max())
    assert np.abs(ta["u_new"] - tb["u_new"]).max() <= 1e-10 * scale
    assert np.abs(ta["u_old"] - tb["u_old"]).max() <= 1e-10 * scale
    assert np.array_equal(a.particles()[0], b.particles()[0])
    # without the generators in the file the restart draws different proposals
    stripped = {k: v for k, v in state.items() if not k.startswith("random-")}
    c = make()
    c.restore(stripped)
    c.trace_enable()
    c.sweep(sweeps_after)
    assert not np.array_equal(c.particles()[0], a.particles()[0])


def test_restore_continues_the_proposal_stream_atomic():
    cfg = small_electrolyte(n=120, moves_per_sweep=50,
                            coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 9.0, "alpha": 0.35, "ncutoff": 5})
    _continue_equals_restart(lambda: oracle_sim(cfg))


def test_restore_continues_the_proposal_stream_water(water_input):
    """rigid molecules: `moltransrot` picks the molecule with the GLOBAL generator (src/move.cpp:1622)"""
    import copy
    cfg = copy.deepcopy(water_input)
    cfg["moves"] = [m for m in cfg["moves"] if "volume" not in m]
    for m in cfg["moves"]:
        for body in m.values():
            body["repeat"] = 20
    _continue_equals_restart(lambda: oracle_sim(cfg), 1, 1)
