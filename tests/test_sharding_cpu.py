"""Host logic of the multi-GPU sharding on CPU: Widom insertions split over ranks with
torch.distributed gloo (world_size 2) give bit-identical averages to an unsharded run. The energy
evaluation is the oracle's here; on the GPU box the same driver code runs on the B200 terms
(tests/test_gpu_parity.py::test_widom_sharded_matches_unsharded)."""
import json
import os
import socket
import sys

import numpy as np

from _oraclelib import ORACLE_SO, oracle_sim
from conftest import small_electrolyte

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ANALYSIS = {"molecule": "ghost", "ninsert": 37}  # odd on purpose: ragged slices
RDF = {"name1": "Na", "name2": "Cl", "dr": 0.2, "file": "rdf.dat"}


def widom_config():
    return small_electrolyte(n=120, coulomb={"type": "fanourgakis", "epsr": 78.7, "cutoff": 10.0}, ghost_pairs=1,
                             moves_per_sweep=30)


def test_sharded_slices_in_one_process():
    """two 'ranks' in one process: slices evaluated separately, collected in order == plain sample()"""
    ref, a, b = oracle_sim(widom_config()), oracle_sim(widom_config()), oracle_sim(widom_config())
    wr, wa, wb = (s.widom_create(ANALYSIS) for s in (ref, a, b))
    for event in range(3):
        for s in (ref, a, b):
            s.sweep(1)
        ref.widom_sample(wr, 1)
        mailbox = {}

        def gather_for(rank):
            def gather(local, counts):
                mailbox[rank] = local.copy()
                return None
            return gather

        # emulate the collective: first pass records both slices, second pass collects
        n = a.api.widom_prepare(a.handle, wa)
        assert n == b.api.widom_prepare(b.handle, wb) == ANALYSIS["ninsert"]
        bounds = [n * r // 2 for r in range(3)]
        parts = []
        for rank, (s, w) in enumerate(((a, wa), (b, wb))):
            local = np.zeros(bounds[rank + 1] - bounds[rank])
            import ctypes as C
            rc = s.api.widom_evaluate_slice(s.handle, w, bounds[rank], len(local),
                                            local.ctypes.data_as(C.POINTER(C.c_double)))
            assert rc == 0
            parts.append(local)
        everyone = np.concatenate(parts)
        for s, w in ((a, wa), (b, wb)):
            import ctypes as C
            assert s.api.widom_collect(s.handle, w, everyone.ctypes.data_as(C.POINTER(C.c_double)), n) == 0
    rr, ra, rb = ref.widom_result(wr), a.widom_result(wa), b.widom_result(wb)
    assert rr["count"] == ra["count"] == rb["count"] == 3 * ANALYSIS["ninsert"]
    assert rr["sum_exp"] == ra["sum_exp"] == rb["sum_exp"]
    assert np.array_equal(rr["last_du"], ra["last_du"])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, out_dir):
    import ctypes as C
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from faunus_b200._simapi import SimLibrary, Simulation
    from faunus_b200.replica import torch_all_gather
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sim = Simulation(SimLibrary(C.CDLL(ORACLE_SO), "fo"), widom_config())
    w = sim.widom_create(ANALYSIS)
    gather = torch_all_gather()
    for _ in range(3):
        sim.sweep(1)
        assert sim.widom_sample_sharded(w, rank, world, gather) == ANALYSIS["ninsert"]
    res = sim.widom_result(w)
    # pair-distance histogram: every rank counts its share of the pairs, an integer all-reduce adds them up
    from faunus_b200.replica import all_reduce_pair_counts
    rdf = sim.rdf_create(RDF)
    sim.rdf_sample_shard(rdf, rank, world)
    local = sim.rdf_result(rdf)[1]
    local_pairs = int(local.sum())
    counts = all_reduce_pair_counts(local)
    json.dump({"sum_exp": res["sum_exp"], "count": res["count"], "last_du": res["last_du"].tolist(),
               "rdf": counts.tolist(), "rdf_local_pairs": local_pairs},
              open(os.path.join(out_dir, f"rank{rank}.json"), "w"))
    sim.close()
    dist.destroy_process_group()


def test_widom_sharded_gloo_world2(tmp_path):
    """world_size-2 gloo run: every rank ends with the averages of the unsharded run, bit for bit"""
    import torch.multiprocessing as mp
    ref = oracle_sim(widom_config())
    wr = ref.widom_create(ANALYSIS)
    for _ in range(3):
        ref.sweep(1)
        ref.widom_sample(wr, 1)
    want = ref.widom_result(wr)
    rdf = ref.rdf_create(RDF)
    ref.rdf_sample(rdf)
    want_rdf = ref.rdf_result(rdf)[1]
    mp.spawn(_gloo_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        got = json.load(open(tmp_path / f"rank{rank}.json"))
        assert got["count"] == want["count"]
        assert got["sum_exp"] == want["sum_exp"]
        assert got["last_du"] == want["last_du"].tolist()
        n = max(len(got["rdf"]), len(want_rdf))
        padded = lambda a: np.pad(np.asarray(a, dtype=np.int64), (0, n - len(a)))
        assert np.array_equal(padded(got["rdf"]), padded(want_rdf.astype(np.int64)))
        assert 0 < got["rdf_local_pairs"] < int(want_rdf.sum())


def _gloo_fast_worker(rank, world, port, out_dir):
    import ctypes as C
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from faunus_b200._simapi import SimLibrary, Simulation
    from faunus_b200.replica import all_reduce_sum
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sim = Simulation(SimLibrary(C.CDLL(ORACLE_SO), "fo"), widom_config())
    w = sim.widom_create({"molecule": "ghost", "ninsert": 600 // world})  # this rank's share of the insertions
    sim.seed_global(1000 + rank)                                          # … drawn from its own generator
    combined = sim.widom_sample_fast(w, 4, all_reduce_sum())
    local = sim.widom_result(w, max_du=1)
    json.dump({"combined": combined, "local_sum": local["sum_exp"], "local_count": local["count"]},
              open(os.path.join(out_dir, f"fast{rank}.json"), "w"))
    sim.close()
    dist.destroy_process_group()


def test_widom_fast_mode_gloo_world2(tmp_path):
    """fast mode: per-rank ghosts from per-rank generators, the averages combined as
    Average::operator+ (src/average.h:61-76) — value sums and sample counts add. Every rank ends with the same
    combined numbers; they equal the sums of the per-rank averages; the excess chemical potential agrees with a
    single-process run of the same total number of insertions within the statistical error."""
    import torch.multiprocessing as mp
    mp.spawn(_gloo_fast_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    got = [json.load(open(tmp_path / f"fast{rank}.json")) for rank in range(2)]
    assert got[0]["combined"] == got[1]["combined"]
    sum_exp, count, mu = got[0]["combined"]
    assert count == got[0]["local_count"] + got[1]["local_count"] == 4 * 600
    assert sum_exp == got[0]["local_sum"] + got[1]["local_sum"]
    assert got[0]["local_sum"] != got[1]["local_sum"]  # different ghosts on the two ranks
    ref = oracle_sim(widom_config())
    wr = ref.widom_create({"molecule": "ghost", "ninsert": 600})
    ref.widom_sample(wr, 4)
    r = ref.widom_result(wr, max_du=1)
    mu_ref = -np.log(r["sum_exp"] / r["count"])
    assert abs(mu - mu_ref) < 0.15  # two independent estimates from 2400 insertions each
