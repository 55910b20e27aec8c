"""The C-ABI shared library loads and exports every symbol include/faunus_b200.h declares (CPU only; no
compute calls without a GPU), and fails loudly when no device is present."""
import ctypes as C
import os
import re

import pytest

import faunus_b200.native as native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "faunus_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    lib = native.load()
    names = declared_functions()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/faunus_b200.h but not exported"
    assert sorted(native.C_ABI_SYMBOLS) == names


def test_host_level_symbols_exported():
    lib = native.load()
    for name in ("fbh_sim_create", "fbh_sim_sweep", "fbh_sim_energy", "fbh_widom_sample", "fbh_sim_trial_set",
                 "fbh_set_device", "fbh_sim_launch_count"):
        assert hasattr(lib, name)


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never fall back"""
    if native.device_count() > 0:
        pytest.skip("a CUDA device is present")
    cfg = native.FbConfig()
    cfg.n_atom_types = 1
    cfg.n_molecule_types = 1
    ctx = C.c_void_p()
    rc = native.load().fb_create(C.byref(cfg), C.byref(ctx))
    assert rc != native.FB_OK and not ctx.value
    assert b"no CUDA device" in native.load().fb_last_error(None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        native.B200Simulation({"temperature": 300})


def test_product_does_not_reference_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/"""
    for base, _, files in os.walk(os.path.join(ROOT, "faunus_b200")):
        if "_build" in base or "__pycache__" in base:
            continue
        for f in files:
            if f.endswith((".py", ".hpp", ".cpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f), errors="replace").read()
                path = os.path.join(base, f)
                assert not re.search(r'#include\s+["<][^">]*oracle', text), f"{path} includes oracle code"
                assert not re.search(r"^\s*(from|import)\s+\S*oracle", text, flags=re.M), f"{path} imports oracle"
                assert "_oraclelib" not in text and "libfaunus_oracle" not in text and "fo_sim_" not in text
