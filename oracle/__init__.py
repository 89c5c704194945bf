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
    flags = ["-O2", "-fPIC", "-fopenmp", "-ffp-contract=off", "-Wall"]
    objs = []
    for src in srcs:
        obj = os.path.join(os.path.dirname(_LIB), os.path.basename(src).replace(".c", ".o"))
        subprocess.run(["gcc", *flags, "-c", src, "-o", obj], check=True, cwd=_HERE)
        objs.append(obj)
        if os.path.basename(src) == "brax_oracle.c":  # second build in float64: the round-off yardstick
            obj64 = obj.replace(".o", "_f64.o")
            subprocess.run(["gcc", *flags, "-DORACLE_F64", "-c", src, "-o", obj64], check=True, cwd=_HERE)
            objs.append(obj64)
    subprocess.run(["gcc", "-shared", "-fopenmp", "-o", _LIB, *objs, "-lm"], check=True, cwd=_HERE)
    return _LIB


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib
