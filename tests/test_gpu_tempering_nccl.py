"""Parallel tempering over the device library's own NCCL communicator (fb_nccl_*; needs ≥ 2 GPUs): the packed
mirror goes GPU to GPU and is imported on the device. Must reproduce, bit for bit, the run in which the same
replicas exchange the reference's messages through the launcher's callbacks (torch.distributed), which in turn is
pinned against the in-process run by tests/test_tempering_cpu.py."""
import json
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _configs(world):
    from faunus_b200.config import primitive_model
    cfgs = []
    for r in range(world):
        cfg = primitive_model(n=600, seed=11, moves_per_sweep=40,
                              coulomb={"type": "ewald", "epsr": 60.0 + 12.0 * r, "cutoff": 12.0, "alpha": 0.25, "ncutoff": 6})
        cfg["moves"].append({"temper": {"format": "xyzqi"}})
        cfgs.append(cfg)
    return cfgs


def _worker(rank, world, port, sweeps, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import faunus_b200.native as native
    from faunus_b200.replica import NcclReplicaSimulation, ReplicaSimulation, TorchReplicaComm
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    dist.barrier(device_ids=[rank])
    out = {}
    for mode in ("callbacks", "nccl"):
        if mode == "nccl":
            sim = NcclReplicaSimulation(_configs(world)[rank], device=rank)
        else:
            native.load().fbh_set_device(rank)
            comm = TorchReplicaComm()
            sim = ReplicaSimulation(native.sim_library(), _configs(world)[rank], comm)
        native.load().fbh_sim_set_window(sim.handle, 64)
        sim.sweep(sweeps)
        temper = [m["temper"] for m in sim.info()["moves"] if "temper" in m][0]
        out[mode] = {"energy": sim.system_energy()[0], "xyzq": sim.particles()[0].reshape(-1).tolist(),
                     "drift": sim.drift(), "exchange": temper["exchange"],
                     "stats": sim.exchange_stats() if mode == "nccl" else None}
        sim.close()
        dist.barrier(device_ids=[rank])
    json.dump(out, open(os.path.join(out_dir, f"rank{rank}.json"), "w"))
    dist.destroy_process_group()


def test_nccl_exchange_equals_callback_exchange(tmp_path):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    sweeps = 30
    mp.spawn(_worker, args=(world, port, sweeps, str(tmp_path)), nprocs=world, join=True)
    accepted = 0
    for rank in range(world):
        got = json.load(open(tmp_path / f"rank{rank}.json"))
        a, b = got["callbacks"], got["nccl"]
        assert a["exchange"] == b["exchange"]
        assert a["xyzq"] == b["xyzq"]          # the same exchanges, the same moves, the same positions
        assert a["energy"] == b["energy"]
        assert abs(b["drift"]) < 1e-9
        assert b["stats"]["messages"] > 0 and b["stats"]["bytes"] > 5 * 8 * 600
        accepted += sum(s["acceptance"] * s["attempts"] for s in b["exchange"].values())
    assert accepted > 0  # some exchanges were accepted: imported states were really used
