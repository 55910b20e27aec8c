"""ctypes binding of libfaunus_b200.so.

Two layers are exported by the library:

* ``fb_*``  — the device-level C ABI of ``include/faunus_b200.h`` (what an ``EnergyTerm`` adaptor
  calls): context, Space mirror, non-bonded ΔU, Ewald, Widom batches, replica state packing.
* ``fbh_*`` — the host-level simulation ABI (C++ MC driver on the B200 adaptor terms), bound by
  :class:`faunus_b200._simapi.SimLibrary`.

The library is required: importing this module raises if it cannot be loaded, and creating a context
without a CUDA device fails loudly (there is no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from ._simapi import SimLibrary, Simulation, c_double_p, c_int_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libfaunus_b200.so")

FB_OK = 0
POT_COULOMB_LJ, POT_COULOMB_WCA, POT_PM, POT_PMWCA, POT_FUNCTOR, POT_SPLINED = range(6)
TERM_COULOMB_SPLINED, TERM_COULOMB_PLAIN, TERM_LJ, TERM_WCA, TERM_HARDSPHERE = 1, 2, 4, 8, 16
MOL_ATOMIC, MOL_RIGID, MOL_COMPRESSIBLE = 1, 2, 4


class FbGroup(C.Structure):
    _fields_ = [("begin", C.c_int), ("size", C.c_int), ("capacity", C.c_int), ("molid", C.c_int),
                ("cm", C.c_double * 3)]


class FbGroupChange(C.Structure):
    _fields_ = [("group_index", C.c_int), ("all", C.c_int), ("internal", C.c_int), ("n_atoms", C.c_int),
                ("atoms", c_int_p)]


class FbChange(C.Structure):
    _fields_ = [("everything", C.c_int), ("volume_change", C.c_int), ("n_groups", C.c_int),
                ("groups", C.POINTER(FbGroupChange)), ("matter_change", C.c_int)]


class FbEwaldConfig(C.Structure):
    _fields_ = [("alpha", C.c_double), ("n_cutoff", C.c_double), ("kappa", C.c_double),
                ("surface_dielectric_constant", C.c_double), ("bjerrum_length", C.c_double),
                ("spherical_sum", C.c_int), ("policy", C.c_int)]


class FbBatchMove(C.Structure):
    _fields_ = [("group_index", C.c_int), ("rel_index", C.c_int), ("atom_id", C.c_int), ("xyzq", C.c_double * 4),
                ("old_atom_id", C.c_int), ("old_xyzq", C.c_double * 4)]


class FbBatchResult(C.Structure):
    _fields_ = [("n_moves", C.c_int), ("stride", C.c_int), ("u_new", c_double_p), ("u_old", c_double_p),
                ("rec_delta", c_double_p), ("cross_new", c_double_p), ("cross_old", c_double_p),
                ("cross_max", c_double_p), ("rec_cross", c_double_p), ("rec_start", C.c_double),
                ("rec_prefactor", C.c_double), ("n_atoms", C.c_int)]


class FbBatchGroupMove(C.Structure):
    _fields_ = [("group_index", C.c_int), ("n_atoms", C.c_int), ("atom_id", C.c_int * 8),
                ("xyzq", (C.c_double * 4) * 8), ("cm", C.c_double * 3), ("old_atom_id", C.c_int * 8),
                ("old_xyzq", (C.c_double * 4) * 8), ("old_cm", C.c_double * 3)]


class FbRunMove(C.Structure):
    _fields_ = [("move", FbBatchMove), ("uniform", C.c_double), ("host_new", C.c_double), ("host_old", C.c_double),
                ("flags", C.c_int), ("depends_on", C.c_int), ("alt", FbBatchMove), ("alt_host_new", C.c_double),
                ("alt_host_old", C.c_double)]


class FbRunConfig(C.Structure):
    _fields_ = [("max_energy", C.c_double), ("cancellation_limit", C.c_double)]


class FbRunResult(C.Structure):
    _fields_ = [("n_moves", C.c_int), ("n_windows", C.c_int), ("n_rounds", C.c_int), ("accepted", C.POINTER(C.c_ubyte)),
                ("u_new", c_double_p), ("u_old", c_double_p)]


class FbTrialMove(C.Structure):
    _fields_ = [("group_index", C.c_int), ("n_atoms", C.c_int), ("rel_index", C.c_int * 8),
                ("xyzq", (C.c_double * 4) * 8), ("atom_id", C.c_int * 8), ("cm", C.c_double * 3),
                ("internal", C.c_int), ("with_ewald", C.c_int)]


c_ubyte_p = C.POINTER(C.c_ubyte)
c_uint32_p = C.POINTER(C.c_uint32)


class FbConfig(C.Structure):
    _fields_ = [
        ("device", C.c_int),
        ("box", C.c_double * 3), ("periodic", C.c_int * 3),
        ("n_atom_types", C.c_int), ("n_molecule_types", C.c_int),
        ("molecule_flags", c_int_p), ("molecule_natoms", c_int_p),
        ("exclusions", C.POINTER(c_ubyte_p)), ("g2g_cutoff_squared", c_double_p),
        ("kind", C.c_int), ("pair_flags", c_uint32_p),
        ("lj_sigma2", c_double_p), ("lj_eps4", c_double_p), ("wca_sigma2", c_double_p),
        ("wca_eps4", c_double_p), ("hs_sigma2", c_double_p),
        ("coulomb_bjerrum_length", C.c_double), ("coulomb_cutoff", C.c_double), ("coulomb_kappa", C.c_double),
        ("coulomb_n_knots", C.c_int), ("coulomb_knots", c_double_p), ("coulomb_coeffs", c_double_p),
        ("plain_bjerrum_length", C.c_double),
        ("spline_offset", c_int_p), ("spline_knots", c_double_p), ("spline_coeffs", c_double_p),
        ("spline_rmin2", c_double_p), ("spline_rmax2", c_double_p), ("spline_hardsphere", c_ubyte_p),
    ]


#: every symbol include/faunus_b200.h declares
C_ABI_SYMBOLS = [
    "fb_create", "fb_destroy", "fb_last_error", "fb_device_count", "fb_upload_space", "fb_update_group",
    "fb_set_box", "fb_sync", "fb_download_space", "fb_nonbonded_energy", "fb_nonbonded_delta",
    "fb_system_energy_shard", "fb_particle_pair_energy", "fb_group_group_energy", "fb_set_force_table", "fb_nonbonded_force", "fb_ewald_force", "fb_atom_rdf", "fb_molecule_rdf", "fb_trial_energy", "fb_trial_commit", "fb_batch_trial", "fb_batch_submit", "fb_batch_submit_groups", "fb_batch_wait", "fb_run_submit", "fb_run_wait", "fb_configure_runs", "fb_get_run_stats", "fb_batch_commit", "fb_configure_cells", "fb_debug_set_cell_capacity", "fb_get_batch_timing", "fb_get_kspace_timing",
    "fb_ewald_configure", "fb_ewald_update_box", "fb_ewald_update_full", "fb_ewald_update_partial",
    "fb_ewald_energy", "fb_ewald_sync", "fb_ewald_download", "fb_debug_fullq_layout", "fb_widom_batch", "fb_state_doubles",
    "fb_export_state", "fb_import_state", "fb_export_state_host", "fb_import_state_host", "fb_upload_groups",
    "fb_nccl_unique_id", "fb_nccl_init", "fb_nccl_finalize", "fb_nccl_exchange_state", "fb_nccl_sendrecv_host",
    "fb_nccl_allgather_host", "fb_nccl_bytes_exchanged", "fb_launch_count",
    "fb_stream", "fb_enable_timing", "fb_last_kernel_ms", "fb_get_timing", "fb_measure_fp64_peak",
]

_lib: Optional[C.CDLL] = None
_simlib: Optional[SimLibrary] = None


def load() -> C.CDLL:
    """Load libfaunus_b200.so (building it first if the sources are newer). Raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("FAUNUS_B200_LIB", LIB_PATH)  # an experimental build (faunus_b200.build --variant)
    if path == LIB_PATH and not os.path.exists(LIB_PATH):
        from .build import build
        build()
    lib = C.CDLL(path)
    vp = C.c_void_p
    sig = {
        "fb_create": (C.c_int, [C.POINTER(FbConfig), C.POINTER(vp)]),
        "fb_destroy": (None, [vp]),
        "fb_last_error": (C.c_char_p, [vp]),
        "fb_device_count": (C.c_int, []),
        "fb_upload_space": (C.c_int, [vp, C.c_int, c_double_p, c_int_p, C.POINTER(FbGroup), C.c_int, C.c_int]),
        "fb_update_group": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(FbGroup), C.c_int, c_int_p, c_double_p,
                                      c_int_p]),
        "fb_set_box": (C.c_int, [vp, C.c_int, c_double_p]),
        "fb_sync": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(FbChange)]),
        "fb_download_space": (C.c_int, [vp, C.c_int, c_double_p, c_int_p, C.POINTER(FbGroup)]),
        "fb_nonbonded_energy": (C.c_int, [vp, C.c_int, C.POINTER(FbChange), c_double_p]),
        "fb_nonbonded_delta": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(FbChange), c_double_p, c_double_p]),
        "fb_trial_energy": (C.c_int, [vp, C.POINTER(FbTrialMove), c_double_p, c_double_p, c_double_p, c_double_p]),
        "fb_trial_commit": (C.c_int, [vp, C.c_int]),
        "fb_system_energy_shard": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p]),
        "fb_set_force_table": (C.c_int, [vp, C.c_int, c_double_p, c_double_p]),
        "fb_nonbonded_force": (C.c_int, [vp, C.c_int, c_double_p]),
        "fb_ewald_force": (C.c_int, [vp, C.c_int, c_double_p]),
        "fb_particle_pair_energy": (C.c_int, [vp, C.c_int, C.c_int, c_double_p, c_int_p, c_double_p, c_int_p, c_double_p]),
        "fb_group_group_energy": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, c_double_p]),
        "fb_batch_trial": (C.c_int, [vp, C.c_int, C.POINTER(FbBatchMove), C.c_int, C.POINTER(FbBatchResult)]),
        "fb_batch_submit": (C.c_int, [vp, C.c_int, C.POINTER(FbBatchMove), C.c_int]),
        "fb_batch_submit_groups": (C.c_int, [vp, C.c_int, C.POINTER(FbBatchGroupMove), C.c_int]),
        "fb_run_submit": (C.c_int, [vp, C.c_int, C.POINTER(FbRunMove), C.c_int, C.POINTER(FbRunConfig)]),
        "fb_run_wait": (C.c_int, [vp, C.POINTER(FbRunResult)]),
        "fb_get_run_stats": (C.c_int, [vp, c_double_p]),
        "fb_configure_runs": (C.c_int, [vp, C.c_int]),
        "fb_get_kspace_timing": (C.c_int, [vp, c_double_p]),
        "fb_batch_wait": (C.c_int, [vp, C.POINTER(FbBatchResult)]),
        "fb_batch_commit": (C.c_int, [vp, C.c_int, c_ubyte_p]),
        "fb_configure_cells": (C.c_int, [vp, C.c_int]),
        "fb_debug_set_cell_capacity": (C.c_int, [vp, C.c_int]),
        "fb_get_batch_timing": (C.c_int, [vp, c_double_p]),
        "fb_ewald_configure": (C.c_int, [vp, C.POINTER(FbEwaldConfig)]),
        "fb_ewald_update_box": (C.c_int, [vp, C.c_int, c_int_p]),
        "fb_ewald_update_full": (C.c_int, [vp, C.c_int]),
        "fb_ewald_update_partial": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(FbChange)]),
        "fb_ewald_energy": (C.c_int, [vp, C.c_int, C.POINTER(FbChange), c_double_p]),
        "fb_ewald_sync": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(FbChange)]),
        "fb_ewald_download": (C.c_int, [vp, C.c_int, c_double_p, c_double_p, c_double_p]),
        "fb_debug_fullq_layout": (C.c_int, [C.POINTER(FbEwaldConfig), c_double_p, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p,
                                            c_int_p, c_int_p, c_int_p, c_int_p, c_int_p]),
        "fb_widom_batch": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_int_p, c_double_p,
                                     C.c_int, c_double_p]),
        "fb_state_doubles": (C.c_size_t, [vp]),
        "fb_export_state": (C.c_int, [vp, C.c_int, vp]),
        "fb_import_state": (C.c_int, [vp, C.c_int, vp]),
        "fb_export_state_host": (C.c_int, [vp, C.c_int, c_double_p]),
        "fb_import_state_host": (C.c_int, [vp, C.c_int, c_double_p]),
        "fb_upload_groups": (C.c_int, [vp, C.c_int, C.POINTER(FbGroup), C.c_int]),
        "fb_nccl_unique_id": (C.c_int, [C.c_char_p]),
        "fb_nccl_init": (C.c_int, [vp, C.c_char_p, C.c_int, C.c_int]),
        "fb_nccl_finalize": (C.c_int, [vp]),
        "fb_nccl_exchange_state": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, c_double_p]),
        "fb_nccl_sendrecv_host": (C.c_int, [vp, c_double_p, C.c_size_t, C.c_int]),
        "fb_nccl_allgather_host": (C.c_int, [vp, C.c_double, c_double_p]),
        "fb_nccl_bytes_exchanged": (C.c_ulonglong, [vp]),
        "fb_launch_count": (C.c_ulonglong, [vp]),
        "fb_stream": (vp, [vp]),
        "fb_enable_timing": (C.c_int, [vp, C.c_int]),
        "fb_last_kernel_ms": (C.c_double, [vp]),
        "fb_get_timing": (C.c_int, [vp, c_double_p]),
        "fb_measure_fp64_peak": (C.c_int, [C.c_int, c_double_p]),
        "fbh_sim_enable_timing": (C.c_int, [vp, C.c_int]),
        "fbh_sim_get_timing": (C.c_int, [vp, c_double_p]),
        "fbh_sim_ctx": (vp, [vp]),
        "fbh_set_device": (None, [C.c_int]),
        "fbh_sim_launch_count": (C.c_ulonglong, [vp]),
        "fbh_sim_set_window": (C.c_int, [vp, C.c_int]),
        "fbh_sim_set_run": (C.c_int, [vp, C.c_int, C.c_int]),
        "fbh_system_energy_shard": (C.c_int, [vp, C.c_int, C.c_int, c_double_p]),
        "fbh_sim_get_window_timing": (C.c_int, [vp, c_double_p]),
    }
    for name, (restype, argtypes) in sig.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def sim_library() -> SimLibrary:
    global _simlib
    if _simlib is None:
        _simlib = SimLibrary(load(), "fbh")
    return _simlib


def device_count() -> int:
    return load().fb_device_count()


def require_device():
    if device_count() == 0:
        raise RuntimeError("faunus_b200: no CUDA device visible; the B200 energy path has no CPU fallback")


class B200Simulation(Simulation):
    """Metropolis MC simulation whose non-bonded/Ewald terms run on the B200 (``fbh_*`` ABI)."""

    #: proposals evaluated per device pass for runs of `transrot` moves (0: one move at a time)
    DEFAULT_WINDOW = int(os.environ.get("FAUNUS_B200_WINDOW", "64"))

    #: single-atom proposals shipped per host round trip when the device walks the windows itself (fb_run_submit;
    #: 0: every window is walked on the host). Needs the full window of 64.
    DEFAULT_RUN = int(os.environ.get("FAUNUS_B200_RUN", "512"))
    #: ... used when at least this many proposals are ready (a single window is walked on the host)
    DEFAULT_RUN_MIN = int(os.environ.get("FAUNUS_B200_RUN_MIN", "65"))

    def __init__(self, config, device: int = 0, window: Optional[int] = None, run: Optional[int] = None,
                 run_min: Optional[int] = None):
        require_device()
        load().fbh_set_device(device)
        super().__init__(sim_library(), config)
        self.run = 0
        self._run_request = self.DEFAULT_RUN if run is None else int(run)
        self._run_min = self.DEFAULT_RUN_MIN if run_min is None else int(run_min)
        self.window = self.set_window(self.DEFAULT_WINDOW if window is None else window)

    def set_window(self, capacity: int) -> int:
        """Windowed evaluation of single-atom move runs (fb_batch_trial); returns the capacity in effect
        (0 when switched off or when the Hamiltonian is not eligible)."""
        self.window = int(load().fbh_sim_set_window(self.handle, int(capacity)))
        self.set_run(self._run_request)
        return self.window

    def set_run(self, moves: int) -> int:
        """Device-decided runs of windows (fb_run_submit): up to `moves` proposals per round trip; returns the
        capacity in effect (0 unless the window is 64 and the Hamiltonian has at most 6 terms)."""
        self._run_request = int(moves)
        self.run = int(load().fbh_sim_set_run(self.handle, int(moves), self._run_min)) if self.window else 0
        return self.run

    # ---- work sharded over the ranks of a process group; see also Simulation.widom_sample_sharded
    def system_energy_shard(self, rank: int, size: int):
        """(non-bonded, reciprocal) share of this rank of the full-system energy"""
        out = np.zeros(2)
        self._check(load().fbh_system_energy_shard(self.handle, rank, size, out.ctypes.data_as(c_double_p)),
                    "system_energy_shard")
        return float(out[0]), float(out[1])

    def configure_cells(self, min_particles: int):
        """Device cell list for the pair part of windows: from `min_particles` slots on (0 always, < 0 never)."""
        self._check(load().fb_configure_cells(self.ctx, int(min_particles)), "fb_configure_cells")

    def window_time_ms(self) -> dict:
        out = np.zeros(10)
        load().fbh_sim_get_window_timing(self.handle, out.ctypes.data_as(c_double_p))
        return {"pair_ms": out[0], "ewald_ms": out[1], "other_ms": out[2], "windows": int(out[3]),
                "moves": int(out[4]), "total_ms": out[5], "host_evaluate_ms": out[6], "host_sweep_ms": out[7],
                "round_trips": int(out[8]), "runs": int(out[9])}

    def kspace_time_ms(self) -> dict:
        """timing enabled: the k-space share of window_time_ms split into the front and the persistent kernel"""
        out = np.zeros(2)
        self._check(load().fb_get_kspace_timing(self.ctx, out.ctypes.data_as(c_double_p)), "fb_get_kspace_timing")
        return {"front_ms": out[0], "kspace_ms": out[1]}

    def configure_runs(self, pair_sums_ahead: bool):
        """fb_configure_runs flags: bit 0 evaluates the pair sums of a window one window ahead (default off: measured
        6.99e5 against 9.0e5 moves/s at S1), bit 1 switches the CUDA-graph replay of a run's launches off (8.26e5)."""
        self._check(load().fb_configure_runs(self.ctx, int(pair_sums_ahead)), "fb_configure_runs")

    def run_stats(self) -> dict:
        """device-decided runs since creation: runs, windows, rounds of the fixed-point walk, moves"""
        out = np.zeros(4)
        self._check(load().fb_get_run_stats(self.ctx, out.ctypes.data_as(c_double_p)), "fb_get_run_stats")
        return {"runs": int(out[0]), "windows": int(out[1]), "rounds": int(out[2]), "moves": int(out[3])}

    @property
    def launch_count(self) -> int:
        return int(load().fbh_sim_launch_count(self.handle))

    @property
    def ctx(self):
        """raw fb_ctx* of the non-bonded/Ewald device context (for direct C-ABI calls)"""
        return load().fbh_sim_ctx(self.handle)

    def enable_timing(self, on: bool = True):
        """CUDA-event timing of every hot-kernel launch (adds ~2 event records per launch)"""
        load().fbh_sim_enable_timing(self.handle, int(on))

    def device_time_ms(self) -> dict:
        out = np.zeros(8)
        load().fbh_sim_get_timing(self.handle, out.ctypes.data_as(c_double_p))
        return {"pair_ms": out[0], "pair_launches": int(out[1]), "ewald_ms": out[2], "ewald_launches": int(out[3]),
                "full_ms": out[4], "full_launches": int(out[5]), "widom_ms": out[6], "widom_launches": int(out[7])}


def make_change(everything=False, volume_change=False, groups: Sequence[dict] = (), matter_change=False):
    """Build an :class:`FbChange` (keeps the backing arrays alive on the returned object)."""
    arr = (FbGroupChange * max(1, len(groups)))()
    keep = []
    for i, g in enumerate(groups):
        idx = np.asarray(g.get("atoms", ()), dtype=np.int32)
        keep.append(idx)
        arr[i].group_index = int(g["group"])
        arr[i].all = int(bool(g.get("all", False)))
        arr[i].internal = int(bool(g.get("internal", False)))
        arr[i].n_atoms = len(idx)
        arr[i].atoms = idx.ctypes.data_as(c_int_p)
    ch = FbChange(int(everything), int(volume_change), len(groups), arr, int(matter_change))
    ch._keep = (arr, keep)
    return ch
