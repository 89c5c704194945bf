"""In-tree build of libcarlb.so (hand-written CUDA for sm_100a behind a C ABI).

``python -m carl_b200.build`` or ``__graft_entry__.build()``. nvcc cross-compiles without a GPU.
The library is written next to its sources (``carl_b200/csrc/libcarlb.so``) so that it travels
with the tree; nothing is JIT-compiled at import time.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libcarlb.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", INCLUDE]
# classic control mirrors float64/NumPy arithmetic, which never fuses a*b+c: no FMA contraction
# there (-fmad=false); the Brax pipeline keeps FMA contraction (the reference's XLA kernels do too).
# Brax: parity first -- without FMA contraction the float32 kernel reproduces the float32 restatement
# of the reference arithmetic to ~1e-6 per env-step (with contraction the stiff joint springs amplify
# the different rounding to a few 1e-5; tests/test_brax_parity_gpu.py measures both against a float64
# yardstick). The FMA-contracted variant is built next to it (brax_fma.cu) and selected per handle
# (carlb_brax_set_arithmetic / CARLBraxEnv(arithmetic="fma")).
UNITS = [
    ("abi.cu", []),
    ("classic.cu", ["-fmad=false"]),
    ("brax.cu", ["-fmad=false"]),   # strict variant (parity mode)
    ("brax_fma.cu", []),            # the same source with FMA contraction (carlb_brax_set_arithmetic)
    ("gather.cu", []),
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libcarlb needs the CUDA toolkit to build")
    return nvcc


def sources() -> list[str]:
    out = [os.path.join(INCLUDE, "carlb.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    for src, extra in UNITS:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((cmd, subprocess.Popen(cmd)))
        objs.append(obj)
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
