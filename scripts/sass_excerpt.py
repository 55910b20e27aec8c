"""cuobjdump -sass of the built library condensed into the excerpt kept under profiles/: per-kernel counts of the
instructions that prove the tensor path (DMMA), cp.async (LDGSTS) and the FP32 screening (FFMA beside DFMA), and the
Gram loop / staging lines of the window kernels.

    python scripts/sass_excerpt.py > profiles/r02_sass_excerpt.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "faunus_b200", "_build", "libfaunus_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
functions, name = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        functions[name] = []
    elif name:
        functions[name].append(line)


def pick(fragment):
    return next((body for n, body in functions.items() if fragment in n), [])


print("# cuobjdump -sass faunus_b200/_build/libfaunus_b200.so (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a), excerpts")
print("# instruction counts per kernel: DMMA (FP64 tensor path, mma.sync.m8n8k4.f64), LDGSTS (cp.async), WARPSYNC")
for kernel in ("windowKspaceKernel", "windowFrontKernel", "batchPairScreenKernelILi1", "windowTailKernelILi1",
               "widomScreenKernelILi1", "fullScreenKernelILi1", "nonbondedForceKernelILi1", "ewaldForceKernel",
               "ewaldFullCellKernel", "ewaldFullGemmKernelILi0", "ewaldFullGemmKernelILi1", "ewaldFullGemmKernelILi2",
               "ewaldStepPhaseKernel"):
    body = pick(kernel)
    count = lambda op: sum(1 for l in body if re.search(r"\b" + op, l))
    print(f"{kernel}: DMMA {count('DMMA')}  LDGSTS {count('LDGSTS')}  DFMA {count('DFMA')}  FFMA {count('FFMA')}  "
          f"WARPSYNC {count('WARPSYNC')}  BAR {count('BAR')}  MUFU {count('MUFU')}")
print()
print("# windowKspaceKernel: per-warp staging (LDGSTS), the block barrier of a unit, and the tile loop (a fragment load feeds")
print("# two tensor instructions; the re step of every tile, then the im step)")
body = pick("windowKspaceKernel")
shown = 0
for i, line in enumerate(body):
    if re.search(r"LDGSTS|BAR\.SYNC|DMMA|LDS\.128", line) and shown < 70:
        print(f"{i}:{line}")
        shown += 1
print()
print("# windowFrontKernel: staging (LDGSTS) and the complex matrix product of the commit (8 DMMA per step of four positions)")
body = pick("windowFrontKernel")
shown = 0
for i, line in enumerate(body):
    if re.search(r"LDGSTS|DMMA", line) and shown < 40:
        print(f"{i}:{line}")
        shown += 1
