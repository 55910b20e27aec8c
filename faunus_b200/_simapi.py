"""ctypes binding of the host-level simulation C ABI (``<prefix>_sim_*`` / ``<prefix>_widom_*``).

The same ABI is exported with prefix ``fbh`` by ``libfaunus_b200.so`` (B200 adaptor terms) and with
prefix ``fo`` by the oracle library (tests / CPU baseline only); this module is the generic binder.
See ``faunus_b200/csrc/host/sim_capi.hpp``.
"""
from __future__ import annotations

import ctypes as C
import json
from typing import Optional, Sequence

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_long_p = C.POINTER(C.c_long)


def _dp(a: Optional[np.ndarray]):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def _ip(a: Optional[np.ndarray]):
    return a.ctypes.data_as(c_int_p) if a is not None else None


class SimLibrary:
    """Function table of one shared library exporting the simulation ABI under ``prefix``."""

    def __init__(self, lib: C.CDLL, prefix: str):
        self.lib = lib
        self.prefix = prefix
        f = self._fn
        f("last_error", C.c_char_p, [])
        f("sim_create", C.c_void_p, [C.c_char_p])
        f("sim_destroy", None, [C.c_void_p])
        f("sim_restore", C.c_int, [C.c_void_p, C.c_char_p])
        f("sim_save_state", C.c_int, [C.c_void_p, C.c_char_p, C.c_int])
        f("sim_load_state", C.c_int, [C.c_void_p, C.c_char_p])
        f("json_to_ubjson", C.c_int, [C.c_char_p, C.c_char_p, C.c_int])
        f("ubjson_to_json", C.c_int, [C.c_char_p, C.c_int, C.c_char_p, C.c_int])
        f("sim_sweep", C.c_int, [C.c_void_p, C.c_int])
        f("sim_moves_per_sweep", C.c_int, [C.c_void_p])
        f("sim_system_energy", C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int, c_int_p])
        f("sim_energy", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  c_int_p, C.c_int, c_double_p])
        f("sim_forces", C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p])
        f("sim_trial_set", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_int_p, C.c_int,
                                     c_double_p, c_double_p, c_double_p])
        f("sim_trial_commit", C.c_int, [C.c_void_p, C.c_int])
        f("sim_seed_global", C.c_int, [C.c_void_p, C.c_uint])
        f("sim_matter_change", C.c_int, [C.c_void_p, C.c_char_p, C.c_int, c_double_p])
        f("sim_drift", C.c_double, [C.c_void_p])
        f("sim_initial_energy", C.c_double, [C.c_void_p])
        f("sim_sum_energy_changes", C.c_double, [C.c_void_p])
        f("sim_trace_enable", None, [C.c_void_p, C.c_int])
        f("sim_trace_size", C.c_long, [C.c_void_p])
        f("sim_trace_get", C.c_long, [C.c_void_p, C.c_long, C.c_long, c_double_p, c_double_p,
                                      c_double_p, c_int_p, c_int_p])
        f("sim_num_particles", C.c_int, [C.c_void_p])
        f("sim_num_groups", C.c_int, [C.c_void_p])
        f("sim_get_particles", C.c_int, [C.c_void_p, C.c_int, c_double_p, c_int_p])
        f("sim_get_groups", C.c_int, [C.c_void_p, C.c_int, c_int_p, c_double_p])
        f("sim_state_json", C.c_int, [C.c_void_p, C.c_char_p, C.c_int])
        f("sim_info_json", C.c_int, [C.c_void_p, C.c_char_p, C.c_int])
        f("widom_create", C.c_int, [C.c_void_p, C.c_char_p])
        f("widom_sample", C.c_int, [C.c_void_p, C.c_int, C.c_int])
        f("widom_result", C.c_int, [C.c_void_p, C.c_int, c_double_p, c_long_p, c_double_p, C.c_int])
        f("widom_prepare", C.c_int, [C.c_void_p, C.c_int])
        f("widom_evaluate_slice", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p])
        f("widom_collect", C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_int])
        f("virtualvolume_create", C.c_int, [C.c_void_p, C.c_char_p])
        f("virtualvolume_sample", C.c_int, [C.c_void_p, C.c_int])
        f("virtualvolume_result", C.c_int, [C.c_void_p, C.c_int, c_double_p])
        f("virtualtranslate_create", C.c_int, [C.c_void_p, C.c_char_p])
        f("virtualtranslate_sample", C.c_int, [C.c_void_p, C.c_int])
        f("virtualtranslate_result", C.c_int, [C.c_void_p, C.c_int, c_double_p])
        f("rdf_create", C.c_int, [C.c_void_p, C.c_char_p])
        f("rdf_sample", C.c_int, [C.c_void_p, C.c_int])
        f("rdf_sample_shard", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int])
        f("rdf_result", C.c_int, [C.c_void_p, C.c_int, c_double_p, C.POINTER(C.c_ulonglong), c_double_p, C.c_int])

    def _fn(self, name, restype, argtypes):
        fn = getattr(self.lib, f"{self.prefix}_{name}")
        fn.restype = restype
        fn.argtypes = argtypes
        setattr(self, name, fn)

    def error(self) -> str:
        return self.last_error().decode("utf-8", "replace")

    def to_ubjson(self, value) -> bytes:
        """Universal Binary JSON of a JSON value as the library's writer encodes it"""
        text = json.dumps(value).encode()
        n = self.json_to_ubjson(text, None, 0)
        if n < 0:
            raise RuntimeError(self.error())
        buf = C.create_string_buffer(max(n, 1))
        self.json_to_ubjson(text, buf, n)
        return buf.raw[:n]

    def from_ubjson(self, data: bytes):
        n = self.ubjson_to_json(data, len(data), None, 0)
        if n < 0:
            raise RuntimeError(self.error())
        buf = C.create_string_buffer(n)
        self.ubjson_to_json(data, len(data), buf, n)
        return json.loads(buf.value.decode())


class Simulation:
    """One Metropolis MC simulation (accepted + trial state) driven through the C ABI."""

    def __init__(self, api: SimLibrary, config: dict | str):
        self.api = api
        text = config if isinstance(config, str) else json.dumps(config)
        self.handle = api.sim_create(text.encode())
        if not self.handle:
            raise RuntimeError(f"{api.prefix}_sim_create: {api.error()}")

    def close(self):
        if getattr(self, "handle", None):
            self.api.sim_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise RuntimeError(f"{self.api.prefix}_{what}: {self.api.error()}")

    # -- state ------------------------------------------------------------------------------
    def restore(self, state: dict | str):
        text = state if isinstance(state, str) else json.dumps(state)
        self._check(self.api.sim_restore(self.handle, text.encode()), "sim_restore")

    def save_state(self, filename: str, save_random: bool = True):
        """`savestate` to a `.json` or `.ubj` file (src/analysis.cpp:640-682)"""
        self._check(self.api.sim_save_state(self.handle, filename.encode(), int(save_random)), "sim_save_state")

    def load_state(self, filename: str):
        """`--state <file>` (.json / .ubj), then restore (src/faunus.cpp:430-455)"""
        self._check(self.api.sim_load_state(self.handle, filename.encode()), "sim_load_state")

    @property
    def num_particles(self) -> int:
        return self.api.sim_num_particles(self.handle)

    def particles(self, which: int = 0):
        n = self.num_particles
        xyzq = np.zeros((n, 4))
        ids = np.zeros(n, dtype=np.int32)
        self.api.sim_get_particles(self.handle, which, _dp(xyzq), _ip(ids))
        return xyzq, ids

    def groups(self, which: int = 0):
        g = self.api.sim_num_groups(self.handle)
        rec = np.zeros((g, 4), dtype=np.int32)
        cm = np.zeros((g, 3))
        self.api.sim_get_groups(self.handle, which, _ip(rec), _dp(cm))
        return rec, cm

    def state_json(self) -> dict:
        n = self.api.sim_state_json(self.handle, None, 0)
        buf = C.create_string_buffer(n)
        self.api.sim_state_json(self.handle, buf, n)
        return json.loads(buf.value.decode())

    def info(self) -> dict:
        n = self.api.sim_info_json(self.handle, None, 0)
        buf = C.create_string_buffer(n)
        self.api.sim_info_json(self.handle, buf, n)
        return json.loads(buf.value.decode())

    # -- energies ---------------------------------------------------------------------------
    def system_energy(self):
        total = C.c_double()
        terms = np.zeros(16)
        n = C.c_int()
        self._check(self.api.sim_system_energy(self.handle, C.byref(total), _dp(terms), 16, C.byref(n)),
                    "sim_system_energy")
        return total.value, terms[: n.value].copy()

    def energy(self, which: int = 0, everything: bool = False, volume_change: bool = False,
               group: int = -1, all: bool = False, internal: bool = False,
               indices: Sequence[int] = ()) -> float:
        idx = np.asarray(indices, dtype=np.int32)
        out = C.c_double()
        self._check(self.api.sim_energy(self.handle, which, int(everything), int(volume_change), group,
                                        int(all), int(internal), _ip(idx), len(idx), C.byref(out)),
                    "sim_energy")
        return out.value

    def forces(self, term: int = -1, which: int = 0) -> np.ndarray:
        """Hamiltonian::force on a zeroed vector (term < 0) or the force of one term; [n_particles, 3], kT/Å"""
        out = np.zeros((self.num_particles, 3))
        self._check(self.api.sim_forces(self.handle, which, term, _dp(out)), "sim_forces")
        return out

    def trial_set(self, group: int, indices: Sequence[int], xyz, all: bool = False,
                  internal: bool = True):
        idx = np.asarray(indices, dtype=np.int32)
        pos = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1)
        u_new, u_old = C.c_double(), C.c_double()
        self._check(self.api.sim_trial_set(self.handle, group, int(all), int(internal), _ip(idx), len(idx),
                                           _dp(pos), C.byref(u_new), C.byref(u_old)), "sim_trial_set")
        return u_new.value, u_old.value

    def trial_commit(self, accept: bool):
        self._check(self.api.sim_trial_commit(self.handle, int(accept)), "sim_trial_commit")

    # -- Monte Carlo ------------------------------------------------------------------------
    def sweep(self, n: int = 1):
        self._check(self.api.sim_sweep(self.handle, n), "sim_sweep")

    @property
    def moves_per_sweep(self) -> int:
        return self.api.sim_moves_per_sweep(self.handle)

    def drift(self) -> float:
        return self.api.sim_drift(self.handle)

    @property
    def initial_energy(self) -> float:
        return self.api.sim_initial_energy(self.handle)

    @property
    def sum_energy_changes(self) -> float:
        return self.api.sim_sum_energy_changes(self.handle)

    def trace_enable(self, on: bool = True):
        self.api.sim_trace_enable(self.handle, int(on))

    def trace(self):
        n = self.api.sim_trace_size(self.handle)
        du, un, uo = np.zeros(n), np.zeros(n), np.zeros(n)
        acc = np.zeros(n, dtype=np.int32)
        mid = np.zeros(n, dtype=np.int32)
        self.api.sim_trace_get(self.handle, 0, n, _dp(du), _dp(un), _dp(uo), _ip(acc), _ip(mid))
        return {"du": du, "u_new": un, "u_old": uo, "accepted": acc, "move_id": mid}

    # -- Widom ------------------------------------------------------------------------------
    def widom_create(self, config: dict) -> int:
        wid = self.api.widom_create(self.handle, json.dumps(config).encode())
        if wid < 0:
            raise RuntimeError(f"{self.api.prefix}_widom_create: {self.api.error()}")
        return wid

    def widom_sample(self, wid: int, nsamples: int = 1):
        self._check(self.api.widom_sample(self.handle, wid, nsamples), "widom_sample")

    def widom_sample_sharded(self, wid: int, rank: int, size: int, all_gather=None) -> int:
        """One Widom sample event with the insertions split over `size` ranks. Every rank calls
        this with simulations in the same state (same seed): the ghosts are generated identically everywhere,
        this rank evaluates its slice, ``all_gather(local, counts) -> np.ndarray`` returns the ΔU of all ranks
        concatenated in rank order, and every rank collects all of them in insertion order — so the running
        averages are bit-identical to an unsharded run."""
        n = self.api.widom_prepare(self.handle, wid)
        if n < 0:
            raise RuntimeError(f"{self.api.prefix}_widom_prepare: {self.api.error()}")
        if n == 0:
            return 0
        bounds = [n * r // size for r in range(size + 1)]
        first, count = bounds[rank], bounds[rank + 1] - bounds[rank]
        local = np.zeros(max(count, 1))
        self._check(self.api.widom_evaluate_slice(self.handle, wid, first, count, _dp(local)), "widom_evaluate_slice")
        local = local[:count]
        everyone = local if size == 1 else all_gather(local, [bounds[r + 1] - bounds[r] for r in range(size)])
        everyone = np.ascontiguousarray(everyone, dtype=np.float64)
        if len(everyone) != n:
            raise RuntimeError("all_gather returned the wrong number of insertion energies")
        self._check(self.api.widom_collect(self.handle, wid, _dp(everyone), n), "widom_collect")
        return n

    def matter_change(self, groups: Sequence[dict], mode: int = 2) -> dict:
        """A change of the number of active particles made by the caller — ``groups`` = [{"index": group, "size": new
        size, "atoms": [relative indices of the particles that appear / disappear], "all", "internal", "dNatomic"}] —
        carried through the MC protocol (updateState, energy on both states, translational-entropy bias, sync);
        mode 0 reject, 1 accept, 2 Metropolis. What a speciation / grand-canonical move asks of the Hamiltonian
        (src/energy.h:1390-1435, src/montecarlo.cpp:271-374)."""
        out = np.zeros(4)
        self._check(self.api.sim_matter_change(self.handle, json.dumps({"groups": list(groups)}).encode(), int(mode),
                                               _dp(out)), "sim_matter_change")
        return {"u_new": out[0], "u_old": out[1], "bias": out[2], "accepted": bool(out[3])}

    def seed_global(self, seed: int):
        """Re-seed the global generator (`Faunus::random`: molecule insertion, Widom ghosts) — per-rank streams of
        :meth:`widom_sample_fast`."""
        self._check(self.api.sim_seed_global(self.handle, int(seed) & 0xFFFFFFFF), "sim_seed_global")

    def widom_sample_fast(self, wid: int, nsamples: int = 1, all_reduce=None):
        """Scalable Widom sampling over the ranks of a process group (SURVEY §8e "fast mode"): the analysis `wid` of
        every rank holds this rank's SHARE of the insertions per sample event (create it with ``ninsert = total //
        size``) and draws its own ghosts from its own generator (:meth:`seed_global` with a per-rank seed) — no
        serial ghost generation that every rank repeats, no gather of insertion energies. The per-rank averages
        combine as the reference's ``Average::operator+`` does (src/average.h:61-76: value sums and sample counts
        add): ``all_reduce(np.array([sum_exp, count])) -> np.ndarray`` sums two doubles over the ranks. Returns
        (sum_exp, count, mu_excess) of the combined average."""
        self.widom_sample(wid, nsamples)
        res = self.widom_result(wid, max_du=1)
        both = np.array([res["sum_exp"], float(res["count"])])
        if all_reduce is not None:
            both = np.asarray(all_reduce(both), dtype=np.float64)
        mu = -np.log(both[0] / both[1]) if both[1] > 0 and both[0] > 0 else float("nan")
        return float(both[0]), int(both[1]), float(mu)

    # -- virtual volume move (excess pressure) --------------------------------------------------------
    def virtualvolume_create(self, config: dict) -> int:
        """`virtualvolume` analysis (dV, scaling); returns its id"""
        vid = self.api.virtualvolume_create(self.handle, json.dumps(config).encode())
        if vid < 0:
            raise RuntimeError(f"{self.api.prefix}_virtualvolume_create: {self.api.error()}")
        return vid

    def virtualvolume_sample(self, vid: int):
        self._check(self.api.virtualvolume_sample(self.handle, vid), "virtualvolume_sample")

    def virtualvolume_result(self, vid: int) -> dict:
        out = np.zeros(4)
        self._check(self.api.virtualvolume_result(self.handle, vid, _dp(out)), "virtualvolume_result")
        return {"sum_exp": out[0], "count": int(out[1]), "last_du": out[2], "excess_pressure_kT_per_A3": out[3]}

    def virtualtranslate_create(self, config: dict) -> int:
        """`virtualtranslate` analysis (molecule, dL, dir); returns its id"""
        vid = self.api.virtualtranslate_create(self.handle, json.dumps(config).encode())
        if vid < 0:
            raise RuntimeError(f"{self.api.prefix}_virtualtranslate_create: {self.api.error()}")
        return vid

    def virtualtranslate_sample(self, vid: int):
        self._check(self.api.virtualtranslate_sample(self.handle, vid), "virtualtranslate_sample")

    def virtualtranslate_result(self, vid: int) -> dict:
        out = np.zeros(4)
        self._check(self.api.virtualtranslate_result(self.handle, vid, _dp(out)), "virtualtranslate_result")
        return {"sum_exp": out[0], "count": int(out[1]), "last_du": out[2], "mean_force_kT_per_A": out[3]}

    # -- atomic radial distribution function ----------------------------------------------------------
    def rdf_create(self, config: dict) -> int:
        """`atomrdf` analysis (name1, name2, dr, slicedir, thickness); returns its id"""
        rid = self.api.rdf_create(self.handle, json.dumps(config).encode())
        if rid < 0:
            raise RuntimeError(f"{self.api.prefix}_rdf_create: {self.api.error()}")
        return rid

    def rdf_sample(self, rid: int):
        self._check(self.api.rdf_sample(self.handle, rid), "rdf_sample")

    def rdf_sample_shard(self, rid: int, rank: int, size: int):
        """This rank's share of one sample (every rank holds the same configuration). The pair counts of the ranks
        add up to the unsharded histogram — sum them with an integer all-reduce (``rdf_result(rid)[1]``)."""
        self._check(self.api.rdf_sample_shard(self.handle, rid, rank, size), "rdf_sample_shard")

    def rdf_result(self, rid: int):
        """(r, exact pair counts per bin, g(r)) accumulated over the samples"""
        n = self.api.rdf_result(self.handle, rid, None, None, None, 0)
        if n < 0:
            raise RuntimeError(f"{self.api.prefix}_rdf_result: {self.api.error()}")
        r, g = np.zeros(n), np.zeros(n)
        pairs = np.zeros(n, dtype=np.uint64)
        self.api.rdf_result(self.handle, rid, _dp(r), pairs.ctypes.data_as(C.POINTER(C.c_ulonglong)), _dp(g), n)
        return r, pairs, g

    def widom_result(self, wid: int, max_du: int = 1 << 20):
        s = C.c_double()
        cnt = C.c_long()
        du = np.zeros(max_du)
        n = self.api.widom_result(self.handle, wid, C.byref(s), C.byref(cnt), _dp(du), max_du)
        return {"sum_exp": s.value, "count": cnt.value, "last_du": du[: min(n, max_du)].copy()}
