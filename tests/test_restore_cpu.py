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


# ---- Universal Binary JSON state files (`savestate` with a .ubj name, src/analysis.cpp:646-668; `--state x.ubj`,
# src/faunus.cpp:430-455). The reference writes them with nlohmann's `json::to_ubjson` defaults: no container size or
# type optimisation, integers in the smallest of i / U / I / l / L, every other number a big-endian 'D'.
def _libraries():
    import faunus_b200.native as native
    from _oraclelib import oracle_api
    return {"oracle": oracle_api(), "product": native.sim_library()}


def test_ubjson_known_encodings():
    """byte strings spelled out from the UBJSON specification (and nlohmann's documented example
    {"compact": true, "schema": 0} → {i\\x07compactTi\\x06schemai\\x00})"""
    import struct
    for lib in _libraries().values():
        assert lib.to_ubjson({"compact": True, "schema": 0}) == b"{i\x07compactTi\x06schemai\x00}"
        assert lib.to_ubjson([1, 2, 3]) == b"[i\x01i\x02i\x03]"
        assert lib.to_ubjson(None) == b"Z" and lib.to_ubjson(False) == b"F"
        assert lib.to_ubjson(-128) == b"i\x80" and lib.to_ubjson(127) == b"i\x7f"
        assert lib.to_ubjson(128) == b"U\x80" and lib.to_ubjson(255) == b"U\xff"
        assert lib.to_ubjson(256) == b"I\x01\x00" and lib.to_ubjson(-129) == b"I\xff\x7f"
        assert lib.to_ubjson(32768) == b"l\x00\x00\x80\x00" and lib.to_ubjson(-40000) == b"l" + struct.pack(">i", -40000)
        assert lib.to_ubjson(2 ** 31) == b"L" + struct.pack(">q", 2 ** 31)
        assert lib.to_ubjson(3.14) == b"D" + struct.pack(">d", 3.14)
        assert lib.to_ubjson("hello") == b"Si\x05hello"
        assert lib.to_ubjson("x" * 300) == b"SI\x01\x2c" + b"x" * 300
        assert lib.to_ubjson({"pos": [0.5, -1.25, 2.0], "id": 3}) == \
            b"{i\x03pos[" + b"".join(b"D" + struct.pack(">d", v) for v in (0.5, -1.25, 2.0)) + b"]i\x02idi\x03}"


def test_ubjson_reader_accepts_what_other_writers_produce():
    """optimised containers ('#' count, '$' type), 'd', 'C'; truncated input is an error, not a crash"""
    import struct
    import pytest
    for lib in _libraries().values():
        assert lib.from_ubjson(b"[#i\x03i\x01i\x02i\x03") == [1, 2, 3]
        assert lib.from_ubjson(b"[$U#i\x02\x07\xff") == [7, 255]
        assert lib.from_ubjson(b"[$D#i\x01" + struct.pack(">d", -2.5)) == [-2.5]
        assert lib.from_ubjson(b"{#i\x01i\x01aT") == {"a": True}
        assert lib.from_ubjson(b"{$i#i\x02i\x01a\x01i\x01b\x02") == {"a": 1, "b": 2}
        assert lib.from_ubjson(b"d" + struct.pack(">f", 0.5)) == 0.5 and lib.from_ubjson(b"CA") == "A"
        value = {"groups": [{"id": 0, "size": 2}], "particles": [{"pos": [1e-3, -7.25, 1e300], "q": -1.0}], "n": 70000,
                 "text": "a\"b", "none": None}
        assert lib.from_ubjson(lib.to_ubjson(value)) == value
        for bad in (b"[i\x01", b"Si\x05hel", b"{i\x03po", b"D\x00\x00", b"?", b"i\x01i\x02"):
            with pytest.raises(RuntimeError, match="ubjson"):
                lib.from_ubjson(bad)


def test_ubjson_state_file_continues_the_run(tmp_path):
    """save to .ubj and to .json in the middle of a run, start two new simulations from the files: both continue the
    uninterrupted run move by move, and the two files hold the same document"""
    import json
    import pytest
    cfg = small_electrolyte(n=120, moves_per_sweep=50,
                            coulomb={"type": "ewald", "epsr": 78.7, "cutoff": 9.0, "alpha": 0.35, "ncutoff": 5})
    a = oracle_sim(cfg)
    a.sweep(2)
    ubj, txt = str(tmp_path / "state.ubj"), str(tmp_path / "state.json")
    a.save_state(ubj)
    a.save_state(txt)
    a.trace_enable()
    a.sweep(2)
    lib = _libraries()["oracle"]
    document = lib.from_ubjson(open(ubj, "rb").read())
    assert document == json.load(open(txt)) and "random-move" in document
    assert open(ubj, "rb").read()[:1] == b"{" and open(ubj, "rb").read() == lib.to_ubjson(document)
    for filename in (ubj, txt):
        b = oracle_sim(cfg)
        b.load_state(filename)
        b.trace_enable()
        b.sweep(2)
        assert np.array_equal(a.trace()["accepted"], b.trace()["accepted"])
        assert np.array_equal(a.particles()[0], b.particles()[0])
    with pytest.raises(RuntimeError, match="unknown file extension"):
        a.save_state(str(tmp_path / "state.xyz"))
    with pytest.raises(RuntimeError, match="state file error"):
        a.load_state(str(tmp_path / "missing.ubj"))
    # saverandom: false leaves the generators out (src/analysis.cpp:661-664)
    a.save_state(ubj, save_random=False)
    assert "random-move" not in lib.from_ubjson(open(ubj, "rb").read())
