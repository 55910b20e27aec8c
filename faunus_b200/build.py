"""Builds libfaunus_b200.so in-tree: CUDA kernels + C ABI (nvcc, sm_100a) and the C++ host adaptor
layer (g++), linked into one shared library under faunus_b200/_build/. No GPU is needed to build."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libfaunus_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
GXX = os.environ.get("GXX", "/usr/bin/g++")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-fno-gnu-unique", "-ccbin", GXX]
GXX_FLAGS = ["-std=c++20", "-O2", "-ffp-contract=off", "-fPIC", "-fvisibility=hidden", "-fno-gnu-unique", "-Wall", "-Wextra",
             "-Wno-unused-parameter"]


def _sources():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files]
    out.append(os.path.join(os.path.dirname(HERE), "include", "faunus_b200.h"))
    return out


def _stale(target: str) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in _sources())


def _run(cmd):
    subprocess.check_call(cmd)


def build(force: bool = False, verbose: bool = False, variant: str | None = None, defines: tuple = ()) -> str:
    """``variant`` + ``defines`` (-D switches of the kernels, e.g. ("FB_KS_GROUPS=2",)) build an experimental copy under
    _build/variants/<variant>/ that ``FAUNUS_B200_LIB`` selects at load time; the product is the default build."""
    out = OUT if variant is None else os.path.join(OUT, "variants", variant)
    lib = os.path.join(out, "libfaunus_b200.so")
    os.makedirs(out, exist_ok=True)
    if not force and not _stale(lib):
        return lib
    dev_o = os.path.join(out, "fb_api.o")
    host_o = os.path.join(out, "fbh_capi.o")
    extra = ["-Xptxas", "-v"] if verbose else []
    extra += [f"-D{d}" for d in defines]
    _run([NVCC, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, "device", "fb_api.cu"), "-o", dev_o])
    _run([GXX, *GXX_FLAGS, "-c", os.path.join(CSRC, "fbh_capi.cpp"), "-o", host_o])
    _run([NVCC, "-shared", "-o", lib, dev_o, host_o, "-cudart", "static", "-ccbin", GXX,
          "-Wno-deprecated-gpu-targets"])
    return lib


if __name__ == "__main__":
    # python -m faunus_b200.build [--force] [--verbose] [--variant NAME -DSWITCH=VALUE …]
    name = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, variant=name,
                defines=tuple(a[2:] for a in sys.argv if a.startswith("-D"))))
