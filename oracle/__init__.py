"""TEST INFRASTRUCTURE — CPU oracle of the batched-step hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product (``carl_b200``) never does.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("classic_oracle.c", "brax_oracle.c")]
    srcs = [s for s in srcs if os.path.exists(s)]
    if (
        not force
        and os.path.exists(_LIB)
        and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in srcs)
    ):
        return _LIB
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    cmd = ["gcc", "-O2", "-fPIC", "-fopenmp", "-ffp-contract=off", "-Wall", "-shared", "-o", _LIB, *srcs, "-lm"]
    subprocess.run(cmd, check=True, cwd=_HERE)
    return _LIB


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib
