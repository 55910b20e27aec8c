"""Parallel tempering across processes: one replica per process / GPU, exchange over torch.distributed.

The C++ ``ParallelTempering`` move (faunus_b200/csrc/host/moves.hpp, restating src/move.cpp:844-968) talks to a
``ReplicaComm``; this module supplies the callback flavour implemented with ``torch.distributed``
point-to-point operations: NCCL send/recv between GPUs (NVLink/NVSwitch) when the process group uses the
``nccl`` backend, gloo on CPU (tests). Messages are the reference's: 8-byte energy and volume
exchanges, group sizes, and the XYZQI particle buffer (src/mpicontroller.cpp:94-246).
"""
from __future__ import annotations

import ctypes as C
import json

import numpy as np

from ._simapi import SimLibrary, Simulation, c_double_p

_BARRIER = C.CFUNCTYPE(None, C.c_void_p)
_SENDRECV = C.CFUNCTYPE(None, C.c_void_p, c_double_p, C.c_size_t, C.c_int)
_GATHER = C.CFUNCTYPE(None, C.c_void_p, C.c_double, c_double_p)


class FbReplicaCallbacks(C.Structure):
    _fields_ = [("rank", C.c_int), ("size", C.c_int), ("user", C.c_void_p), ("barrier", _BARRIER),
                ("sendrecv_replace", _SENDRECV), ("gather", _GATHER)]


class TorchReplicaComm:
    """Exchange primitives over the default ``torch.distributed`` process group."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.size = dist.get_rank(), dist.get_world_size()
        self.device = device if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
        self.bytes_exchanged = 0
        self.exchanges = 0
        self._cb = FbReplicaCallbacks(self.rank, self.size, None, _BARRIER(self._barrier),
                                      _SENDRECV(self._sendrecv), _GATHER(self._gather))

    @property
    def callbacks(self) -> FbReplicaCallbacks:
        return self._cb

    def _barrier(self, _user):
        if self.device.type == "cuda":
            self.dist.barrier(device_ids=[self.device.index])
        else:
            self.dist.barrier()

    def _sendrecv(self, _user, data, n, partner):
        torch, dist = self.torch, self.dist
        host = np.ctypeslib.as_array(data, shape=(n,))
        send = torch.from_numpy(host.copy()).to(self.device)
        recv = torch.empty_like(send)
        ops = [dist.P2POp(dist.isend, send, partner), dist.P2POp(dist.irecv, recv, partner)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        host[:] = recv.cpu().numpy()
        self.bytes_exchanged += 8 * n
        self.exchanges += 1

    def _gather(self, _user, value, out):
        torch, dist = self.torch, self.dist
        mine = torch.tensor([value], dtype=torch.float64, device=self.device)
        everyone = [torch.empty_like(mine) for _ in range(self.size)]
        dist.all_gather(everyone, mine)
        for i, t in enumerate(everyone):
            out[i] = float(t.item())


class ReplicaSimulation(Simulation):
    """A simulation whose ``temper`` move exchanges with the other ranks of the process group."""

    def __init__(self, api: SimLibrary, config, comm: TorchReplicaComm):
        self.comm = comm  # keep the ctypes callbacks alive
        fn = getattr(api.lib, f"{api.prefix}_sim_create_replica")
        fn.restype = C.c_void_p
        fn.argtypes = [C.c_char_p, C.POINTER(FbReplicaCallbacks)]
        self.api = api
        text = config if isinstance(config, str) else json.dumps(config)
        self.handle = fn(text.encode(), C.byref(comm.callbacks))
        if not self.handle:
            raise RuntimeError(f"{api.prefix}_sim_create_replica: {api.error()}")


class NcclReplicaSimulation(Simulation):
    """A replica whose ``temper`` move exchanges over the device library's OWN NCCL communicator (``fb_nccl_*``):
    the packed mirror of the trial state goes GPU to GPU (ncclSend/ncclRecv on device buffers, imported on the
    device), nothing of the exchange passes through Python. ``torch.distributed`` is only used once, to hand the
    128-byte NCCL id of rank 0 to the other ranks."""

    def __init__(self, config, device: int):
        import torch
        import torch.distributed as dist
        from . import native
        lib = native.load()
        native.require_device()
        lib.fbh_set_device(device)
        rank, size = dist.get_rank(), dist.get_world_size()
        uid = C.create_string_buffer(128)
        if rank == 0 and lib.fb_nccl_unique_id(uid) != 0:
            raise RuntimeError("fb_nccl_unique_id: " + lib.fb_last_error(None).decode())
        box = [uid.raw if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        fn = lib.fbh_sim_create_replica_nccl
        fn.restype = C.c_void_p
        fn.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        self.api = native.sim_library()
        text = config if isinstance(config, str) else json.dumps(config)
        self.handle = fn(text.encode(), box[0], rank, size)
        if not self.handle:
            raise RuntimeError(f"fbh_sim_create_replica_nccl: {self.api.error()}")
        lib.fbh_sim_exchange_stats.restype = C.c_int
        lib.fbh_sim_exchange_stats.argtypes = [C.c_void_p, c_double_p]
        self._lib = lib

    def exchange_stats(self) -> dict:
        out = (C.c_double * 2)()
        self._lib.fbh_sim_exchange_stats(self.handle, out)
        return {"messages": int(out[0]), "bytes": int(out[1])}


def run_local_replicas(api: SimLibrary, configs, sweeps: int, packed: bool = False):
    """All replicas in this process, one thread each (``<prefix>_temper_run_local``). ``packed`` (B200 library only):
    the state exchange ships the packed device mirrors (fb_export_state → fb_import_state), the Spaces follow."""
    fn = getattr(api.lib, f"{api.prefix}_temper_run_local" + ("_packed" if packed else ""))
    fn.restype = C.c_int
    fn.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    text = json.dumps(list(configs)).encode()
    size = 1 << 16
    while True:
        buf = C.create_string_buffer(size)
        n = fn(text, sweeps, buf, size)
        if n < 0:
            raise RuntimeError(f"{api.prefix}_temper_run_local: {api.error()}")
        if n <= size:
            return json.loads(buf.value.decode())
        size = n + 16
        # results are deterministic, so re-running with a larger buffer returns the same data


def torch_all_gather(device=None):
    """``all_gather(local, counts)`` for :meth:`Simulation.widom_sample_sharded` over the default process group
    (NCCL between GPUs, gloo on CPU): ragged slices are padded to the longest one."""
    import torch
    import torch.distributed as dist
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")

    def gather(local: np.ndarray, counts):
        width = max(counts)
        mine = torch.zeros(width, dtype=torch.float64, device=device)
        mine[: len(local)] = torch.from_numpy(np.ascontiguousarray(local)).to(device)
        everyone = [torch.empty_like(mine) for _ in counts]
        dist.all_gather(everyone, mine)
        return np.concatenate([t[:c].cpu().numpy() for t, c in zip(everyone, counts)])

    return gather


def reduce_in_rank_order(values, device=None):
    """Sum of per-rank partial energies in rank order on every rank (deterministic): one gather into
    one device tensor, one copy to the host, the ranks added in order."""
    import torch
    import torch.distributed as dist
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(values), dtype=torch.float64, device=device)
    everyone = torch.empty(dist.get_world_size() * len(mine), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(everyone, mine)
    rows = everyone.cpu().numpy().reshape(dist.get_world_size(), len(mine))
    total = np.zeros(len(mine))
    for row in rows:
        total += row
    return total


def all_reduce_sum(device=None):
    """``all_reduce(values) -> np.ndarray``: sum of a few doubles over the default process group (the two numbers of
    :meth:`Simulation.widom_sample_fast`); the ranks are added in rank order, so every rank holds the same bits."""
    def reduce(values):
        return reduce_in_rank_order(values, device)
    return reduce


def all_reduce_pair_counts(counts, device=None):
    """Sum of the per-rank pair-distance histograms of :meth:`Simulation.rdf_sample_shard` over the default process
    group (an integer all-reduce: exact, whatever the order; NCCL between GPUs, gloo on CPU). Histograms that grew to
    different lengths on the ranks are padded to the longest one."""
    import torch
    import torch.distributed as dist
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    length = torch.tensor([len(counts)], dtype=torch.int64, device=device)
    dist.all_reduce(length, op=dist.ReduceOp.MAX)
    total = torch.zeros(int(length.item()), dtype=torch.int64, device=device)
    total[: len(counts)] = torch.from_numpy(np.ascontiguousarray(counts).astype(np.int64)).to(device)
    dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return total.cpu().numpy().astype(np.uint64)
