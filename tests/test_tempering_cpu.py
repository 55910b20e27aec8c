"""Parallel tempering host logic on CPU: the restated `temper` move (src/move.cpp:844-968) with the
in-process communicator and, across processes, over torch.distributed gloo (world_size 2)."""
import json
import os
import socket
import sys

import numpy as np
import pytest

from _oraclelib import ORACLE_SO, oracle_api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def temper_config(scale, micro_repeat=10):
    """examples/temper/temper.yml: one particle in the hard-coded 1D `example2d` potential"""
    return {
        "temperature": 300, "random": {"seed": "fixed"},
        "geometry": {"type": "cuboid", "length": [4, 4, 4]},
        "atomlist": [{"A": {"dp": 0.1}}],
        "moleculelist": [{"mygroup": {"atoms": ["A"], "atomic": True, "insdir": [1, 1, 0]}}],
        "insertmolecules": [{"mygroup": {"N": 1}}],
        "energy": [{"example2d": {"scale": scale, "2D": False}}],
        "moves": [{"transrot": {"molecule": "mygroup", "dir": [1, 1, 0], "repeat": micro_repeat}},
                  {"temper": {"format": "xyz", "volume_scale": "isotropic"}}],
    }


def electrolyte_replicas(n_replicas):
    """Hamiltonian tempering of a small electrolyte: replicas differ in eps_r"""
    from faunus_b200.config import primitive_model
    cfgs = []
    for r in range(n_replicas):
        cfg = primitive_model(n=60, seed=11, moves_per_sweep=20,
                              coulomb={"type": "fanourgakis", "epsr": 60.0 + 15.0 * r, "cutoff": 10.0})
        cfg["moves"].append({"temper": {"format": "xyzqi"}})
        cfgs.append(cfg)
    return cfgs


def test_temper_example_local():
    from faunus_b200.replica import run_local_replicas
    scales = np.geomspace(1.0, 0.1, 4)
    res = run_local_replicas(oracle_api(), [temper_config(float(s)) for s in scales], sweeps=400)
    assert len(res) == 4 and all(r["error"] == "" for r in res)
    attempts = 0
    for r in res:
        temper = [m["temper"] for m in r["moves"] if "temper" in m][0]
        assert temper["replicas"] == 4
        for stat in temper["exchange"].values():
            attempts += stat["attempts"]
            assert 0.0 <= stat["acceptance"] <= 1.0
        x = r["xyzq"][0]
        assert -2.0 <= x <= 2.0
    assert attempts > 0
    # partners take the same decision: the two ends of a pair report identical statistics
    t0 = [m["temper"] for m in res[0]["moves"] if "temper" in m][0]["exchange"]["0 <-> 1"]
    t1 = [m["temper"] for m in res[1]["moves"] if "temper" in m][0]["exchange"]["0 <-> 1"]
    assert t0 == t1


def test_temper_electrolyte_local_energy_consistency():
    from faunus_b200.replica import run_local_replicas
    res = run_local_replicas(oracle_api(), electrolyte_replicas(2), sweeps=30)
    assert all(r["error"] == "" for r in res)
    for r in res:
        assert abs(r["drift"]) < 1e-9  # ΣΔU bookkeeping survives the state swaps


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, configs, sweeps, out_dir):
    import ctypes as C
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from faunus_b200._simapi import SimLibrary
    from faunus_b200.replica import ReplicaSimulation, TorchReplicaComm
    dist.init_process_group("gloo", rank=rank, world_size=world)
    api = SimLibrary(C.CDLL(ORACLE_SO), "fo")
    comm = TorchReplicaComm()
    sim = ReplicaSimulation(api, configs[rank], comm)
    sim.sweep(sweeps)
    xyzq, _ = sim.particles()
    total, _ = sim.system_energy()
    json.dump({"energy": total, "xyzq": xyzq.reshape(-1).tolist(), "drift": sim.drift(), "info": sim.info(),
               "bytes": comm.bytes_exchanged}, open(os.path.join(out_dir, f"rank{rank}.json"), "w"))
    sim.close()
    dist.destroy_process_group()


def test_temper_gloo_matches_local(tmp_path):
    """world_size-2 gloo run == in-process run, bit for bit (same moves, same exchanges)"""
    import torch.multiprocessing as mp
    from faunus_b200.replica import run_local_replicas
    configs = electrolyte_replicas(2)
    sweeps = 25
    local = run_local_replicas(oracle_api(), configs, sweeps=sweeps)
    mp.spawn(_gloo_worker, args=(2, _free_port(), configs, sweeps, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        got = json.load(open(tmp_path / f"rank{rank}.json"))
        assert got["energy"] == local[rank]["energy"]
        assert got["xyzq"] == local[rank]["xyzq"]
        assert abs(got["drift"]) < 1e-9
        assert got["bytes"] > 0
        temper = [m["temper"] for m in got["info"]["moves"] if "temper" in m][0]
        assert temper["replicas"] == 2


def test_temper_local_replicas_at_different_temperatures(tmp_path):
    """Temperature tempering inside ONE process: the input temperature is a process-global in the reference
    (`pc::temperature`, src/faunus.cpp:106; one MPI process per replica). The in-process replicas are constructed
    one at a time so that each builds its Bjerrum length and kJ/mol tables at its OWN temperature: equal, bit for
    bit, to the run with one process per replica."""
    import torch.multiprocessing as mp
    from faunus_b200.replica import run_local_replicas
    configs = electrolyte_replicas(2)
    configs[0]["temperature"], configs[1]["temperature"] = 300.0, 345.0
    for cfg in configs:  # same dielectric constant: the replicas differ in temperature only
        cfg["energy"][0]["nonbonded_coulombwca"]["coulomb"]["epsr"] = 70.0
    sweeps = 12
    for attempt in range(3):  # a construction race would show up as run-to-run differences
        local = run_local_replicas(oracle_api(), configs, sweeps=sweeps)
        assert all(r["error"] == "" for r in local)
        if attempt == 0:
            first = local
        else:
            assert [r["energy"] for r in local] == [r["energy"] for r in first]
    mp.spawn(_gloo_worker, args=(2, _free_port(), configs, sweeps, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        got = json.load(open(tmp_path / f"rank{rank}.json"))
        assert got["energy"] == local[rank]["energy"]
        assert got["xyzq"] == local[rank]["xyzq"]
