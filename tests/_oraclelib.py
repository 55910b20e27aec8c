"""Loads the ORACLE library (oracle/_build/libfaunus_oracle.so) for tests. Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from faunus_b200._simapi import SimLibrary, Simulation, c_double_p, c_int_p

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libfaunus_oracle.so")


def build_oracle(force: bool = False) -> str:
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".hpp", ".cpp"))]
    host = os.path.join(ROOT, "faunus_b200", "csrc", "host")
    srcs += [os.path.join(host, f) for f in os.listdir(host)]
    stale = force or not os.path.exists(ORACLE_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)
    return ORACLE_SO


_lib = None
_api = None


def oracle_lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle())
        _lib.fo_andrea_test.restype = C.c_int
        _lib.fo_andrea_test.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, c_double_p, C.c_int,
                                        c_double_p, C.c_int, c_int_p]
        _lib.fo_andrea_test_eval.restype = C.c_double
        _lib.fo_andrea_test_eval.argtypes = [C.c_double] * 5
        _lib.fo_coulomb_table.restype = C.c_int
        _lib.fo_coulomb_table.argtypes = [C.c_char_p, C.c_double, c_double_p, c_double_p, C.c_int, c_double_p,
                                          c_double_p, c_double_p, c_double_p]
        _lib.fo_pair_energy.restype = C.c_int
        _lib.fo_pair_energy.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, c_double_p, C.c_int, c_double_p]
        _lib.fo_pair_force.restype = C.c_int
        _lib.fo_pair_force.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, c_double_p, C.c_int, c_double_p]
        _lib.fo_ewald_kat.restype = C.c_int
        _lib.fo_ewald_kat.argtypes = [C.c_char_p, C.c_double, C.c_double, c_double_p, C.c_int, c_double_p,
                                      c_double_p]
        _lib.fo_set_parallel_ewald_init.argtypes = [C.c_int]
        _lib.fo_set_large_system_mode.argtypes = [C.c_int]
        _lib.fo_openmp_threads.restype = C.c_int
    return _lib


def oracle_api() -> SimLibrary:
    global _api
    if _api is None:
        _api = SimLibrary(oracle_lib(), "fo")
    return _api


def oracle_sim(config) -> Simulation:
    return Simulation(oracle_api(), config)


def pair_force(config: dict, nonbonded_name: str, id_a: int, id_b: int, r_vectors) -> np.ndarray:
    """force on a due to b (kT/Å) for distance vectors b → a, [n, 3]"""
    import json
    r = np.ascontiguousarray(r_vectors, dtype=np.float64).reshape(-1, 3)
    f = np.zeros_like(r)
    rc = oracle_lib().fo_pair_force(json.dumps(config).encode(), nonbonded_name.encode(), id_a, id_b,
                                    r.ctypes.data_as(c_double_p), len(r), f.ctypes.data_as(c_double_p))
    if rc != 0:
        raise RuntimeError(oracle_api().error())
    return f


def pair_energy(config: dict, nonbonded_name: str, id_a: int, id_b: int, r) -> np.ndarray:
    import json
    r = np.ascontiguousarray(r, dtype=np.float64)
    u = np.zeros_like(r)
    rc = oracle_lib().fo_pair_energy(json.dumps(config).encode(), nonbonded_name.encode(), id_a, id_b,
                                     r.ctypes.data_as(c_double_p), len(r), u.ctypes.data_as(c_double_p))
    if rc != 0:
        raise RuntimeError(oracle_api().error())
    return u
