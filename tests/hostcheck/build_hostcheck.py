"""TEST-ONLY: g++ build of the host-compiled kernel-source check shim (see hostcheck.cpp)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libhostcheck.so")
ROOT = os.path.dirname(os.path.dirname(HERE))


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, "hostcheck.cpp"), os.path.join(HERE, "hostcheck_brax.cpp")]
    csrc = os.path.join(ROOT, "carl_b200", "csrc")
    deps = srcs + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", LIB, *srcs], check=True)
    return LIB
