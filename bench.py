#!/usr/bin/env python
"""Benchmark of the B200 ΔU hot path (driver contract: one JSON line on stdout from rank 0).

Workload (BASELINE.json metric "MC trial moves/s ... at N=1e5", SURVEY §8(d) S1 "pm-1e5"):
restricted-primitive-model 1:1 electrolyte, N = 100 000 ions in one atomic group, 1.0 M, cubic PBC box
L = 436.25 Å, T = 298.15 K, eps_r = 78.7, `nonbonded_coulombwca` with Ewald real space (alpha 0.12,
cutoff 28 Å) + reciprocal space (ncutoff 30 → K = 56 k k-vectors, policy PBC), single-ion `transrot`
moves (dp = 4 Å), fixed seed. One "step" = one sweep of MOVES_PER_STEP trial moves through the
Metropolis engine on the `Energy::EnergyTerm` adaptor terms; runs of `transrot` proposals travel to the
device a run at a time (fb_run_submit: windows of 64 evaluated speculatively and walked in order on the
device), everything else through updateState → energy(trial) → energy(accepted) → sync.

  value  : trial moves/s counting only device time (inputs resident in HBM: CUDA-event time from the first
           to the last kernel of every run / window)
  e2e    : trial moves/s end to end through the engine (proposals drawn on the host from the host Space →
           H2D per run → kernels → decisions and energies D2H → replay into the host Spaces),
           wall/CUDA-event bracketed, L2 flushed between steps
  --impl reference : the same moves by the CPU restatement of the reference path (oracle, built with the
           reference's Release flags + OpenMP) on the host cores, bounded sample.

N > 1 GPUs: single-move ΔU is not split across GPUs (replicas only); every rank runs an
independent replica of the workload with its own seed offset and `value` is the sum (weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# moves per step (sweep): a fifth of the reference's own sweep for this system (`repeat: N`); FAUNUS_B200_MOVES_PER_STEP
# for side-by-side runs (the queue of proposals drains at the end of every sweep)
MOVES_PER_STEP = int(os.environ.get("FAUNUS_B200_MOVES_PER_STEP", "20000"))
N_IONS = 100_000
FLOP_PER_PAIR = 49          # splined Coulomb + WCA, pair within the cutoffs, SURVEY §8(d)
FLOP_PER_FAR_PAIR = 25      # pair beyond both cutoffs: min-image r² (20) + sqrt, +eps, r<Rc test (3) + WCA cut test (2)
FLOP_PER_K_MOVE = 40        # per (k-vector, move): 2 factorised phases (24) + δ (4) + A_k(2 Re(conj Q δ)+|δ|²) (12)
FLOP_PER_K_CROSS = 4        # per (k-vector, ordered pair of moves): A_k Re(conj δ_a δ_m) = 2 FMA
BYTES_PER_PARTICLE = 36     # x, y, z, q doubles + int32 id
BYTES_PER_KVECTOR = 64      # k (24) + A_k (8) + Q read (16) + Q write (16), k-vector components read


#: the workloads of SURVEY §8(d). "s1" is the headline line; the others are extra measurements (--workload)
WORKLOADS = {
    "s1": {"n": 100_000, "coulomb": {"type": "ewald", "epsr": 78.7, "cutoff": 28.0, "alpha": 0.12, "ncutoff": 30,
                                      "ewaldscheme": "PBC"},
           "name": "pm-1e5: RPM 1:1 electrolyte N=100000, 1.0 M, L=436.25 A, nonbonded_coulombwca, "
                   "Ewald alpha=0.12 Rc=28 ncutoff=30 (K=56k, PBC), single-ion transrot dp=4"},
    # the regime in which k-space is HBM-bound: short real-space cutoff, K = 2.4e6 k-vectors
    "s1-largeK": {"n": 100_000, "coulomb": {"type": "ewald", "epsr": 78.7, "cutoff": 14.0, "alpha": 0.22, "ncutoff": 104,
                                             "ewaldscheme": "PBC"},
                  "name": "pm-1e5-largeK: RPM 1:1 electrolyte N=100000, 1.0 M, L=436.25 A, nonbonded_coulombwca, "
                          "Ewald alpha=0.22 Rc=14 ncutoff=104 (K=2.4M, PBC), single-ion transrot dp=4"},
    # N = 1e6: pair part through the device cell list
    "s2": {"n": 1_000_000, "coulomb": {"type": "ewald", "epsr": 78.7, "cutoff": 28.0, "alpha": 0.12, "ncutoff": 65,
                                        "ewaldscheme": "PBC"},
           "name": "pm-1e6: RPM 1:1 electrolyte N=1000000, 1.0 M, L=939.9 A, nonbonded_coulombwca, device cell list, "
                   "Ewald alpha=0.12 Rc=28 ncutoff=65 (K=5.7e5, PBC), single-ion transrot dp=4"},
}
ACTIVE_WORKLOAD = "s1"


def workload(moves_per_step=MOVES_PER_STEP, n=None, seed=5489, summation_policy="serial", which=None):
    from faunus_b200.config import primitive_model
    w = WORKLOADS[which or ACTIVE_WORKLOAD]
    return primitive_model(n=n or w["n"], molarity=1.0, seed=seed, moves_per_sweep=moves_per_step,
                           summation_policy=summation_policy, coulomb=dict(w["coulomb"]))


WORKLOAD_NAME = WORKLOADS["s1"]["name"]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}",
                 "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.samples.append((time.perf_counter(), parts))

    def stop(self, t_begin=None, t_end=None):
        """median SM clock / reasons of the samples that arrived in [t_begin, t_end] (the window under load)"""
        if self.proc:
            self.proc.terminate()
        window = [p for t, p in self.samples if (t_begin is None or t >= t_begin) and (t_end is None or t <= t_end)]
        sm = sorted(int(s[0]) for s in window if s[0].isdigit())
        mx = [int(s[1]) for s in window if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in window for i in range(4) if s[2 + i].lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm),
                "window": "warm-up + timed steps + per-kernel timing pass (same workload), nvidia-smi every 50 ms"}


def dist_setup(n_gpus: int, backend: str):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist_mod.init_process_group(backend=backend)
        dist = dist_mod
    return rank, world, local, dist


def build_oracle_native():
    """(Re)build the timed CPU baseline with -march=native on THIS box"""
    out = os.path.join(ROOT, "oracle", "_build", "native")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libfaunus_oracle_fast.so")
    flags = ["-std=c++20", "-fopenmp", "-fPIC", "-fvisibility=hidden", "-fno-gnu-unique", "-O3", "-ffast-math",
             "-fno-finite-math-only", "-march=native", "-shared"]
    try:
        subprocess.check_call(["/usr/bin/g++", *flags, "-o", so, os.path.join(ROOT, "oracle", "oracle.cpp")],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return so, "native"
    except (OSError, subprocess.CalledProcessError):
        return os.path.join(ROOT, "oracle", "_build", "libfaunus_oracle_fast.so"), "x86-64-v3 (prebuilt)"


def cpu_reference_run(steps: int, warmup: int, moves_per_step: int, policy: str):
    """Times the CPU restatement of the reference path (the oracle) on the host cores."""
    import ctypes as C
    from faunus_b200._simapi import SimLibrary, Simulation
    so, arch = build_oracle_native()
    lib = C.CDLL(so)
    lib.fo_set_parallel_ewald_init.argtypes = [C.c_int]
    lib.fo_set_openmp_threads.argtypes = [C.c_int]
    lib.fo_openmp_threads.restype = C.c_int
    lib.fo_set_openmp_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1 to its workers
    # one-off start-up passes (N·K structure factor, all-pairs energy of the start configuration) over all cores and
    # computed once for the two identical states; the timed per-move path stays as in the reference
    lib.fo_set_large_system_mode.argtypes = [C.c_int]
    lib.fo_set_large_system_mode(1)
    threads = lib.fo_openmp_threads() if policy == "openmp" else 1
    sim = Simulation(SimLibrary(lib, "fo"), workload(moves_per_step, summation_policy=policy))
    sim.trace_enable()
    for _ in range(warmup):
        sim.sweep(1)
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.sweep(1)
    dt = time.perf_counter() - t0
    trace = sim.trace()
    sim.close()
    lib.fo_set_large_system_mode(0)
    return {"moves_per_s": steps * moves_per_step / dt, "seconds": dt, "threads": threads, "arch": arch,
            "policy": policy, "moves": steps * moves_per_step, "trace": trace,
            "sweeps": warmup + steps, "moves_per_sweep": moves_per_step}


def parity_against_cpu_sample(native, device, cpu_run):
    """The trial moves the CPU baseline just timed, again on the device (same workload, same seed, runs walked on the
    device): identical accept/reject sequence, u_new / u_old within 1e-10 of the largest energy. Raises otherwise."""
    import numpy as np
    ref = cpu_run["trace"]
    sim = native.B200Simulation(workload(cpu_run["moves_per_sweep"]), device=device, run_min=1)
    sim.trace_enable()
    sim.sweep(cpu_run["sweeps"])
    got = sim.trace()
    sim.close()
    n = len(ref["du"])
    if len(got["du"]) != n or not np.array_equal(got["accepted"], ref["accepted"]):
        raise AssertionError("parity: the accept/reject sequence differs from the CPU port's")
    scale = max(np.abs(ref["u_new"]).max(), np.abs(ref["u_old"]).max())
    worst = max(np.abs(got["u_new"] - ref["u_new"]).max(), np.abs(got["u_old"] - ref["u_old"]).max()) / scale
    if not worst <= 1e-10:
        raise AssertionError(f"parity: per-move energies differ by {worst:.3e} (relative) from the CPU port's")
    return {"moves": n, "max_relative_energy_difference": float(worst), "acceptance": float(ref["accepted"].mean())}


def reference_arm(args):
    # CPU only: under torchrun rank 0 alone runs and prints, the other ranks exit without work (no rendezvous)
    if int(os.environ.get("RANK", "0")) != 0:
        return
    moves = 25  # bounded sample per step: ~4 ms/move → 0.1 s/step
    res = cpu_reference_run(args.steps, min(args.warmup, 3), moves, "openmp")
    sample = f"{res['moves']} single-ion trial moves of the N=1e5 workload ({moves}/step), -march={res['arch']}, " \
             f"summation_policy=openmp (pair sum over {res['threads']} threads; Ewald k-loops serial as in the reference); " \
             "ewaldscheme PBC, the reference's default: PBCEigen only vectorises the full rebuild of Q(k) and the energy sum " \
             "(src/energy.cpp:208-217, src/energy.h:220-235), the per-move partial update is the same serial loop"
    line = {
        "impl": "reference", "metric": "MC trial moves/s", "value": res["moves_per_s"], "unit": "moves/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * res["seconds"] / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME, "moves_per_step": moves},
        "cpu_baseline": {"value": res["moves_per_s"], "unit": "moves/s", "cores": res["threads"], "kind": "port",
                         "sample": sample},
        "e2e": {"value": res["moves_per_s"], "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def kernel_dram_traffic(kernel, workload_key="s1"):
    """dram__bytes_read + dram__bytes_write per launch of `kernel` from the newest committed `ncu --set full`
    summary of THIS workload under profiles/ (scripts/ncu_summary.py; captures of the other workloads carry their
    key in the file name: r02n_largeK_summary.csv, …), or (None, why)"""
    import csv
    import glob
    tags = {"s1-largeK": "largeK", "s2": "s2"}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_summary.csv")), reverse=True):
        name = os.path.basename(path)
        if workload_key in tags:
            if tags[workload_key] not in name:
                continue
        elif any(t in name for t in tags.values()):
            continue
        try:
            rows = list(csv.reader(open(path)))
        except OSError:
            continue
        if not rows or "dram_read_MB" not in rows[0]:
            continue
        i_r, i_w = rows[0].index("dram_read_MB"), rows[0].index("dram_write_MB")
        vals = [(float(r[i_r]) + float(r[i_w])) * 1e6 for r in rows[1:] if r and r[0].startswith(kernel)]
        if vals:
            return sum(vals) / len(vals), f"dram__bytes_read+write per launch, profiles/{os.path.basename(path)}"
    return None, "no ncu --set full summary of this kernel under profiles/"


def measure_fp64_peak(native, device):
    import ctypes as C
    lib = native.load()
    if not hasattr(lib, "fb_measure_fp64_peak"):
        return None
    lib.fb_measure_fp64_peak.restype = C.c_int
    lib.fb_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    out = C.c_double()
    return out.value if lib.fb_measure_fp64_peak(device, C.byref(out)) == 0 else None


def sharded_extras(args, native, torch, dist, rank, world, local):
    """Work that shards over GPUs, same S1 state on every rank (same seed): Widom ion-pair
    insertions (parity mode: the same ghosts everywhere, index ranges, all-gather of the insertion energies; fast
    mode: per-rank ghosts from per-rank generators, averages combined as Average::operator+), full-system energy
    split by tile rows (pair part) and k-vector slabs (reciprocal part), pair-distance histogram split by tile rows,
    parallel tempering with one replica per GPU over the device library's own NCCL communicator.
    Device-event / wall time, max over ranks; strong scaling (total work fixed) except tempering (one replica per GPU)."""
    from faunus_b200.config import primitive_model
    from faunus_b200.replica import all_reduce_pair_counts, all_reduce_sum, reduce_in_rank_order, torch_all_gather
    import math
    # Widom: the reference's Ewald term returns the TOTAL reciprocal energy for any non-empty change and never
    # sees the ghost, which makes exp(-dU) underflow; the insertion workload therefore uses the
    # cutoff scheme (Fanourgakis) on the same ion configuration. The system energy uses the Ewald S1 workload.
    cfg_widom = primitive_model(n=N_IONS, molarity=1.0, seed=5489, moves_per_sweep=10, ghost_pairs=1,
                                coulomb={"type": "fanourgakis", "epsr": 78.7, "cutoff": 28.0})
    sim = native.B200Simulation(cfg_widom, device=local)
    n_insert = 32768
    wid = sim.widom_create({"molecule": "ghost", "ninsert": n_insert})
    gather = torch_all_gather() if world > 1 else None

    def timed(fn, reps):
        fn()  # warm-up
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier(device_ids=[local])
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    sim.enable_timing(True)
    t0 = sim.device_time_ms()
    dt_widom = timed(lambda: sim.widom_sample_sharded(wid, rank, world, gather), 3)
    t1 = sim.device_time_ms()
    widom_kernel_ms = (t1["widom_ms"] - t0["widom_ms"]) / max(1, t1["widom_launches"] - t0["widom_launches"])
    res = sim.widom_result(wid)
    # fast mode: this rank's share of the insertions, its own ghosts
    share = n_insert // world
    fast = sim.widom_create({"molecule": "ghost", "ninsert": share})
    sim.seed_global(7001 + rank)
    combined = {}

    def fast_sample():
        combined["sum_exp"], combined["count"], combined["mu"] = sim.widom_sample_fast(
            fast, 1, all_reduce_sum() if world > 1 else None)

    dt_fast = timed(fast_sample, 3)
    sim.close()
    sim = native.B200Simulation(workload(moves_per_step=10, which="s1"), device=local)
    shares = {}

    def energy():
        nb, rec = sim.system_energy_shard(rank, world)
        tot = reduce_in_rank_order([nb, rec]) if world > 1 else [nb, rec]
        shares["nonbonded"], shares["reciprocal"] = float(tot[0]), float(tot[1])

    dt_energy = timed(energy, 2)
    # full rebuild of Q(k) from the positions (every volume move, tempering exchange, restart): the matrix-product kernel
    # of fb_fullq.cuh, device time of step phases + product + gather (CUDA events on the context's stream)
    lib = native.load()
    lib.fb_enable_timing(sim.ctx, 1)
    full_q_ms = []
    for _ in range(4):
        if lib.fb_ewald_update_full(sim.ctx, 0) != 0:
            raise RuntimeError("fb_ewald_update_full failed")
        full_q_ms.append(lib.fb_last_kernel_ms(sim.ctx))
    lib.fb_enable_timing(sim.ctx, 0)
    full_q_ms = min(full_q_ms[1:])
    # atomrdf (the dominant non-energy cost of examples/bulk): every Na-Cl pair of the configuration into a
    # distance histogram; tile rows dealt to the ranks, integer all-reduce of the histograms
    rdf = sim.rdf_create({"name1": "Na", "name2": "Cl", "dr": 0.1, "file": "rdf.dat"})
    rdf_total = {}

    def rdf_sample():
        sim.rdf_sample_shard(rdf, rank, world)
        counts = sim.rdf_result(rdf)[1]
        rdf_total["pairs"] = int((all_reduce_pair_counts(counts) if world > 1 else counts).sum())

    dt_rdf = timed(rdf_sample, 2)
    n_active = N_IONS
    out = {
        "scaling": "strong", "n_gpus": world,
        "widom": {"insertions_per_s": n_insert / dt_widom, "insertions_per_sample": n_insert, "ghost_atoms": 2,
                  "ms_per_sample_e2e": 1e3 * dt_widom, "kernel_ms_this_rank": widom_kernel_ms,
                  "pair_interactions_per_s": n_insert * 2 * n_active / dt_widom,
                  "mu_excess_kT": -math.log(res["sum_exp"] / res["count"]) if res["count"] and res["sum_exp"] > 0 else None,
                  "workload": "S1 ion configuration, nonbonded_coulombwca with Fanourgakis Rc=28, Na+Cl ghost pair",
                  "note": "parity mode: ghost generation (host, reference RNG order, the same on every rank) and the "
                          "all-gather of the slices are inside e2e; bit-identical to the unsharded run",
                  "fast_mode": {"insertions_per_s": share * world / dt_fast, "insertions_per_rank": share,
                                "ms_per_sample_e2e": 1e3 * dt_fast, "mu_excess_kT": combined.get("mu"),
                                "samples_combined": combined.get("count"),
                                "note": "per-rank ghosts (per-rank generators), averages combined as "
                                        "Average::operator+ (two doubles all-reduced)"}},
        "system_energy": {"ms": 1e3 * dt_energy, "nonbonded_kT": shares.get("nonbonded"),
                          "reciprocal_kT": shares.get("reciprocal"),
                          "pairs": N_IONS * (N_IONS - 1) // 2, "n_times_k": N_IONS * 57950},
        "full_q": {"ms": full_q_ms, "kernel": "ewaldStepPhaseKernel + ewaldFullGemmKernel + ewaldFullGatherKernel",
                   "algorithmic_tflops": 8.0 * N_IONS * 57950 / (full_q_ms / 1e3) / 1e12,
                   "flop_model": "8 per particle and k-vector (one complex multiply-add); not sharded: every rank rebuilds "
                                 "its own Q(k)"},
        "atom_rdf": {"ms_per_sample_e2e": 1e3 * dt_rdf, "pairs_per_sample": (N_IONS // 2) ** 2,
                     "pair_distances_per_s": (N_IONS // 2) ** 2 / dt_rdf,
                     "pairs_counted_after_3_samples": rdf_total.get("pairs"),
                     "note": "Na-Cl, dr = 0.1 A, exact integer histogram (shared-memory atomics), D2H of the histogram inside"},
    }
    sim.close()
    if world > 1:
        out["temper"] = temper_extras(native, torch, dist, rank, world, local)
    return out


def temper_extras(native, torch, dist, rank, world, local):
    """Hamiltonian parallel tempering: one replica per GPU, replicas differ in eps_r; the
    `temper` move of every sweep ships the packed mirror of the accepted state GPU to GPU over the device library's
    own NCCL communicator (ncclSend/ncclRecv on device buffers, imported on the device) and exchanges the 8-byte
    energy change the same way; nothing of the exchange passes through Python."""
    from faunus_b200.config import primitive_model
    from faunus_b200.replica import NcclReplicaSimulation
    n, moves, sweeps = 20000, 200, 24
    cfg = primitive_model(n=n, molarity=1.0, seed=5489, moves_per_sweep=moves,
                          coulomb={"type": "ewald", "epsr": 78.7 * (1.0 + 0.01 * rank), "cutoff": 14.0, "alpha": 0.22,
                                   "ncutoff": 12, "ewaldscheme": "PBC"})
    cfg["moves"].append({"temper": {"format": "xyzqi"}})
    sim = NcclReplicaSimulation(cfg, device=local)
    native.load().fbh_sim_set_window(sim.handle, 64)
    sim.sweep(8)  # communicator set-up; both odd/even partner pairings occur
    torch.cuda.synchronize()
    dist.barrier(device_ids=[local])
    x0 = sim.exchange_stats()
    t0 = time.perf_counter()
    sim.sweep(sweeps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    x1 = sim.exchange_stats()
    info = sim.info()
    temper = [m["temper"] for m in info["moves"] if "temper" in m][0]
    drift = sim.drift()
    out = {"replicas": world, "particles_per_replica": n, "sweeps_per_s": sweeps / dt,
           "moves_per_s_all_replicas": world * sweeps * moves / dt,
           "exchange_attempts_per_s_this_rank": sum(s["attempts"] for s in temper["exchange"].values()) / dt,
           "bytes_exchanged_per_sweep_this_rank": (x1["bytes"] - x0["bytes"]) / sweeps,
           "messages_per_sweep_this_rank": (x1["messages"] - x0["messages"]) / sweeps,
           "exchange_statistics_rank0": temper["exchange"], "relative_drift_rank0": drift,
           "transport": "fb_nccl_* (libnccl opened by the device library): packed mirror device to device"}
    sim.close()
    return out


def sharded_summary(extras, world):
    """The numbers of `sharded` a reader of the LAST 1500 characters of the line needs"""
    if not extras or "error" in extras:
        return {"n_gpus": world, "error": (extras or {}).get("error", "no extras")}
    w = extras["widom"]
    out = {"n_gpus": world, "widom_fast_ins_per_s": round(w["fast_mode"]["insertions_per_s"]),
           "widom_parity_ins_per_s": round(w["insertions_per_s"]), "widom_kernel_ms_rank": round(w["kernel_ms_this_rank"], 3),
           "energy_ms": round(extras["system_energy"]["ms"], 3), "rdf_ms": round(extras["atom_rdf"]["ms_per_sample_e2e"], 3)}
    if "full_q" in extras:
        out["full_q_ms"] = round(extras["full_q"]["ms"], 3)
    if "temper" in extras:
        out["temper_sweeps_per_s"] = round(extras["temper"]["sweeps_per_s"], 1)
        out["temper_moves_per_s_all"] = round(extras["temper"]["moves_per_s_all_replicas"])
    return out


def b200_arm(args):
    rank, world, local, dist = dist_setup(args.gpus, "nccl")
    import torch
    import faunus_b200.native as native
    native.require_device()
    torch.cuda.set_device(local)
    if dist is not None:
        dist.barrier(device_ids=[local])
    sampler = ClockSampler(local)  # nvidia-smi needs ~1 s to start: launched before the simulation is set up
    sampler.start()
    cfg = workload(seed=5489 + rank)
    sim = native.B200Simulation(cfg, device=local, window=args.window)
    if os.environ.get("FAUNUS_B200_RUN_FLAGS"):  # experiments: fb_configure_runs flags (1: pair sums ahead, 2: no graphs)
        sim.configure_runs(int(os.environ["FAUNUS_B200_RUN_FLAGS"]))
    n = sim.num_particles
    info0 = sim.info()
    kvectors = None
    for term in info0["energy"]:
        if "ewald" in term:
            kvectors = term["ewald"].get("wavefunctions") or kvectors
    # L2 flush between steps: a 256 MiB write (B200 L2 = 126 MB) on torch's stream, synchronised before the sweep
    flush_buffer = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")

    def flush_l2():
        flush_buffer.fill_(1)
        torch.cuda.synchronize()

    t_load_begin = time.perf_counter()
    for _ in range(args.warmup):
        flush_l2()
        sim.sweep(1)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier(device_ids=[local])
    launches0 = sim.launch_count
    wt0 = sim.window_time_ms()
    rs0 = sim.run_stats() if sim.window else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush_l2()
        sim.sweep(1)
    torch.cuda.synchronize()
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    wt1 = sim.window_time_ms()
    rs1 = sim.run_stats() if sim.window else None
    event_s = ev0.elapsed_time(ev1) / 1e3
    launches = sim.launch_count - launches0
    elapsed = max(wall, event_s)
    if dist is not None:
        t = torch.tensor([elapsed], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    moves = args.steps * MOVES_PER_STEP
    e2e = world * moves / elapsed
    # second pass over the same number of steps with CUDA-event timing of the hot kernels (events on the
    # context's own stream): device-only time, inputs resident in HBM
    sim.enable_timing(True)
    w0, s0 = sim.window_time_ms(), sim.device_time_ms()
    k0 = sim.kspace_time_ms()
    for _ in range(args.steps):
        sim.sweep(1)
    w1, s1 = sim.window_time_ms(), sim.device_time_ms()
    k1 = sim.kspace_time_ms()
    front_ms, kspace_only_ms = k1["front_ms"] - k0["front_ms"], k1["kspace_ms"] - k0["kspace_ms"]
    sim.enable_timing(False)
    clocks = sampler.stop(t_load_begin, time.perf_counter())
    windowed = sim.window > 0
    if windowed:
        pair_ms = w1["pair_ms"] - w0["pair_ms"]
        ewald_ms = w1["ewald_ms"] - w0["ewald_ms"]
        other_ms = w1["other_ms"] - w0["other_ms"]
        n_windows = w1["windows"] - w0["windows"]
        evaluated = w1["moves"] - w0["moves"]
    else:
        pair_ms = s1["pair_ms"] - s0["pair_ms"]
        ewald_ms = s1["ewald_ms"] - s0["ewald_ms"]
        other_ms = 0.0
        n_windows = s1["pair_launches"] - s0["pair_launches"]
        evaluated = n_windows
    device_s = (pair_ms + ewald_ms + other_ms) / 1e3
    if windowed:  # first to last kernel of every window in the e2e pass (pair kernel beside the k-space kernels)
        device_s = (wt1["total_ms"] - wt0["total_ms"]) / 1e3
    if dist is not None:
        t = torch.tensor([device_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        device_s = float(t.item())
    value = world * moves / device_s if device_s > 0 else e2e
    sim_info_moves = sim.info()["moves"]
    sim.close()
    extras = None
    if not args.no_extras:
        try:  # the extras must never cost the headline line
            extras = sharded_extras(args, native, torch, dist, rank, world, local)
        except Exception as e:  # noqa: BLE001
            extras = {"error": f"{type(e).__name__}: {e}"}
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    acceptance = None
    for m in sim_info_moves:
        for body in m.values():
            if isinstance(body, dict) and body.get("moves"):
                acceptance = body.get("acceptance")
    acceptance = acceptance if acceptance is not None else 0.0
    traffic, traffic_source = kernel_dram_traffic("windowKspaceKernel", ACTIVE_WORKLOAD)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    roofline = None
    if pair_ms > 0 and n_windows:
        L = cfg["geometry"]["length"]
        L = L if isinstance(L, (int, float)) else L[0]
        in_range = (4.0 / 3.0) * 3.141592653589793 * WORKLOADS[ACTIVE_WORKLOAD]["coulomb"]["cutoff"] ** 3 / L ** 3
        flop_pair = FLOP_PER_FAR_PAIR + in_range * (FLOP_PER_PAIR - FLOP_PER_FAR_PAIR)
        moves_per_launch = evaluated / n_windows
        fp64_peak = measure_fp64_peak(native, local)
        nominal = 148 * 64 * 2 * 1.965e9 / 1e12
        peak = fp64_peak or nominal
        peak_source = ("FP64 DFMA microbenchmark on this GPU in this run (fb_measure_fp64_peak); nominal "
                       f"148 SM x 64 FMA/clk x 1.965 GHz = {nominal:.1f} TFLOP/s") if fp64_peak else "nominal"
        # pair kernel of the windowed path (+ its sums / cross terms): every move of a window against all N
        pair_s = pair_ms / 1e3 / n_windows
        pair_flop = 2 * moves_per_launch * (n - 1) * flop_pair
        pair = {"kernel": "batchPairScreenKernel<COULOMB_WCA>" if windowed else "trialMoveKernel<COULOMB_WCA>",
                "bound": "fp64",
                "note": "algorithmic FP64 work of the reference's pair loop (SURVEY 8d) over the kernel's time; the "
                        "distance test is screened in FP32 (rigorous widening of the cutoff), the FP64 pipe only "
                        "evaluates the candidates, so the figure may exceed what the FP64 pipe alone could do",
                "achieved": pair_flop / pair_s / 1e12, "peak": peak, "unit": "TFLOP/s",
                "frac": pair_flop / pair_s / 1e12 / peak, "us_per_launch": pair_s * 1e6,
                "algorithmic_flop_per_launch": pair_flop, "flop_per_pair": flop_pair,
                "pairs_per_s": 2 * moves_per_launch * (n - 1) / pair_s,
                "hbm": {"algorithmic_bytes_per_launch": BYTES_PER_PARTICLE * n,
                        "achieved_gbs": BYTES_PER_PARTICLE * n / pair_s / 1e9, "peak_gbs": hbm_peak,
                        "frac": BYTES_PER_PARTICLE * n / pair_s / 1e9 / hbm_peak, "peak_source": peak_src}}
        roofline = dict(pair)
        if windowed and ewald_ms > 0 and kvectors:
            # dominant kernel of a window: the persistent k-space kernel (+ the phase-table kernel before it)
            accepted_per_window = acceptance * moves_per_launch  # acceptance measured in this run (move statistics)
            ew_flop = kvectors * (moves_per_launch * FLOP_PER_K_MOVE +
                                  moves_per_launch * (moves_per_launch - 1) / 2 * FLOP_PER_K_CROSS +
                                  accepted_per_window * 26 + 4)
            ew_s = ewald_ms / 1e3 / n_windows
            # the dominant kernel itself: the persistent k-space kernel (δ of every move, R, the Gram matrix); the
            # commit of the previous window and Σ A_k|Q_k|² are the front kernel's (26 per committed move + 4)
            ks_flop = kvectors * (moves_per_launch * FLOP_PER_K_MOVE +
                                  moves_per_launch * (moves_per_launch - 1) / 2 * FLOP_PER_K_CROSS)
            ks_s = (kspace_only_ms if kspace_only_ms > 0 else ewald_ms) / 1e3 / n_windows
            both = {"kernel": "windowFrontKernel + windowKspaceKernel (round 1's batchKspaceKernel did both)",
                    "achieved": ew_flop / ew_s / 1e12, "frac": ew_flop / ew_s / 1e12 / peak, "us_per_launch": ew_s * 1e6,
                    "algorithmic_flop_per_launch": ew_flop, "front_us_per_launch": front_ms / n_windows * 1e3,
                    "flop_model": "per k-vector: 40 per move + 4 per ordered pair of moves + 26 per committed move + 4"}
            roofline = {
                "kernel": "windowKspaceKernel",
                "bound": "fp64",
                "achieved": ks_flop / ks_s / 1e12, "peak": peak, "unit": "TFLOP/s",
                "frac": ks_flop / ks_s / 1e12 / peak, "us_per_launch": ks_s * 1e6,
                "traffic": traffic, "traffic_source": traffic_source, "acceptance": acceptance,
                "algorithmic_flop_per_launch": ks_flop,
                "algorithmic_bytes_per_launch": kvectors * 40,
                "flop_model": "per k-vector: 40 per move + 4 per ordered pair of moves",
                "with_front_kernel": both,
                # the same launch against the HBM roofline: Q(k) read + written, k-vector data read, once per window
                "hbm": {"algorithmic_bytes_per_launch": kvectors * 40,
                        "achieved_gbs": kvectors * 40 / ew_s / 1e9, "peak_gbs": hbm_peak,
                        "frac": kvectors * 40 / ew_s / 1e9 / hbm_peak, "peak_source": peak_src},
                "pair_kernel": pair,
            }
        roofline["peak_source"] = peak_source
        roofline["moves_per_launch"] = moves_per_launch
        roofline["note"] = ("FP64-pipe bound, not HBM bound (hbm.frac): the 13 MB working set is read once per window and "
                            "is L2-resident; the Gram part of the k-space kernel and the commit of the previous window run "
                            "on the FP64 tensor path (DMMA, measured peak 37.2 TFLOP/s vs 33.9 for DFMA)")
        if extras and extras.get("widom", {}).get("kernel_ms_this_rank"):
            w = extras["widom"]
            pairs = w["insertions_per_sample"] / world * w["ghost_atoms"] * n
            tf = pairs * FLOP_PER_FAR_PAIR / (w["kernel_ms_this_rank"] / 1e3) / 1e12
            roofline["widom_kernel"] = {"kernel": "widomScreenKernel<COULOMB_WCA> (FP32 screening, FP64 candidates)",
                                        "bound": "fp64", "achieved": tf,
                                        "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                                        "pairs_per_s": pairs / (w["kernel_ms_this_rank"] / 1e3),
                                        "ms_per_launch": w["kernel_ms_this_rank"]}
        if extras and extras.get("full_q", {}).get("ms"):
            f = extras["full_q"]
            roofline["full_q_kernel"] = {"kernel": "ewaldFullGemmKernel (Q = [X.Y].[Z] on the FP64 tensor path, + step phases "
                                                   "and gather)", "bound": "fp64", "achieved": f["algorithmic_tflops"],
                                         "peak": peak, "unit": "TFLOP/s", "frac": f["algorithmic_tflops"] / peak,
                                         "ms_per_launch": f["ms"], "flop_model": f["flop_model"]}
    cpu = None
    parity = None
    if not args.no_cpu_baseline:
        r = None
        try:
            r = cpu_reference_run(steps=4, warmup=1, moves_per_step=25, policy="serial")
            cpu = {"value": r["moves_per_s"], "unit": "moves/s", "cores": 1, "kind": "port",
                   "sample": f"{r['moves']} trial moves of the same N=1e5 workload, serial summation "
                             f"(reference default), -O3 -ffast-math -march={r['arch']}; start-up passes parallelised"}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": "moves/s", "cores": 1, "kind": "port", "sample": f"failed: {e}"}
        if r is not None:  # a parity failure must be loud: no bench line without it
            parity = parity_against_cpu_sample(native, local, r)
    runs = None
    if windowed:
        # windows walked on the host: the window description in (sizeof(BatchInput) = 10288 B), the result block out
        # (8 + 3S + 4S² doubles, S = 64); runs (windows walked on the device): header + 112 B per proposal in,
        # state + 24 B per decision out
        r = {k: rs1[k] - rs0[k] for k in rs1}
        host_windows = (wt1["windows"] - wt0["windows"]) - r["windows"]
        h2d_per_step = int((host_windows * 10288 + r["runs"] * 32 + r["moves"] * 112) / args.steps)
        d2h_per_step = int((host_windows * 8 * (8 + 3 * 64 + 4 * 64 * 64) + r["runs"] * 288 + r["moves"] * 24) / args.steps)
        if r["runs"]:
            runs = {"runs_per_step": r["runs"] / args.steps, "moves_per_run": r["moves"] / r["runs"],
                    "windows_per_run": r["windows"] / r["runs"], "walk_rounds_per_window": r["rounds"] / r["windows"],
                    "host_round_trips_per_step": (wt1["round_trips"] - wt0["round_trips"]) / args.steps,
                    "note": "proposals on distinct atoms are shipped a run at a time; the device walks each window "
                            "(fixed-point Metropolis walk) and feeds the next; the host replays the decisions"}
    else:
        h2d_per_step, d2h_per_step = 36 * MOVES_PER_STEP, 24 * MOVES_PER_STEP
    line = {
        "metric": "MC trial moves/s", "value": value, "unit": "moves/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[ACTIVE_WORKLOAD]["name"], "moves_per_step": MOVES_PER_STEP, "kvectors": kvectors,
                   "parallelism": f"replicas x{world}" if world > 1 else "single GPU",
                   "l2": "flushed between steps (256 MiB device write + synchronize inside the timed region); within a "
                         "step the 13 MB working set (positions, Q(k), k tables) is L2-resident by design"},
        "e2e": {"value": e2e, "unit": "moves/s", "h2d_bytes_per_step": h2d_per_step, "d2h_bytes_per_step": d2h_per_step},
        "gpu_launches": launches,
        "pair_interactions_per_s": e2e * 2 * (n - 1),
        "window": sim.window,
        "runs": runs,
        "host_split_us_per_move": {
            "evaluate_launch_and_wait": 1e3 * (wt1["host_evaluate_ms"] - wt0["host_evaluate_ms"]) / moves,
            "draw_decide_sync_spaces": 1e3 * ((wt1["host_sweep_ms"] - wt0["host_sweep_ms"]) -
                                              (wt1["host_evaluate_ms"] - wt0["host_evaluate_ms"])) / moves,
            "device_first_to_last_kernel": 1e3 * (wt1["total_ms"] - wt0["total_ms"]) / moves} if sim.window else None,
        "device_time_split_us_per_move": {"pair": 1e3 * pair_ms / moves, "kspace": 1e3 * ewald_ms / moves,
                                          "commit_phase_finish": 1e3 * other_ms / moves},
        "clocks": clocks,
    }
    if extras:
        line["sharded"] = extras
    if roofline:
        line["roofline"] = roofline
    if cpu:
        line["cpu_baseline"] = cpu
    if parity:
        line["parity_checked_moves"] = parity["moves"]
        line["parity"] = parity
    if not args.no_extras:
        line["sharded_summary"] = sharded_summary(extras, world)  # last: survives a truncated tail of the line
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sharded Widom / system-energy measurements")
    ap.add_argument("--window", type=int, default=None, help="proposals per device pass (0: one move per launch)")
    ap.add_argument("--workload", default="s1", choices=sorted(WORKLOADS),
                    help="s1: the headline line; s1-largeK / s2: extra measurements (no extras, no CPU sample)")
    args = ap.parse_args()
    global ACTIVE_WORKLOAD
    ACTIVE_WORKLOAD = args.workload
    if args.workload != "s1":
        args.no_extras = True
        args.no_cpu_baseline = True
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        pass


if __name__ == "__main__":
    main()
