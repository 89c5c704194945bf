"""ctypes binding of libcarlb (``include/carlb.h``).

There is no CPU fallback: if the CUDA library is missing or does not load, importing this
module's users fails loudly (``NativeLibraryError``).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int64, c_uint32, c_uint64, c_void_p

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libcarlb.so")

CARLB_OK = 0
ERR_INVALID, ERR_STATE, ERR_CUDA = -1, -2, -3
F32, F64 = 0, 1
ACT_I32, ACT_I64, ACT_U8, ACT_F32 = 0, 1, 2, 3
AUTORESET_NONE, AUTORESET_SAME_STEP = 0, 1
MAX_PEERS, MAX_MIXED = 8, 8

KIND = {
    "cartpole": 0, "pendulum": 1, "acrobot": 2, "mountaincar": 3, "mountaincar_cont": 4,
    "brax_ant": 16, "brax_halfcheetah": 17, "brax_hopper": 18, "brax_walker2d": 19,
    "brax_inverted_pendulum": 20, "brax_inverted_double_pendulum": 21, "brax_reacher": 22, "brax_humanoid": 23, "brax_humanoidstandup": 24, "brax_pusher": 25,
}

EXPORTS = [
    "carlb_abi_version", "carlb_last_error", "carlb_query_env", "carlb_env_create", "carlb_env_destroy",
    "carlb_env_bind", "carlb_env_configure", "carlb_env_seed", "carlb_env_reset", "carlb_env_step",
    "carlb_env_step_host", "carlb_env_step_host_checked", "carlb_env_step_host_begin", "carlb_env_step_host_end", "carlb_stage_actions", "carlb_env_rollout", "carlb_mixed_step",
    "carlb_brax_set_system", "carlb_brax_set_arithmetic", "carlb_brax_set_reset_rng", "carlb_brax_reset_from_q", "carlb_brax_goal_step", "carlb_launch_count",
    "carlb_gather_create", "carlb_gather_export", "carlb_gather_open", "carlb_gather_attach", "carlb_gather_wait",
    "carlb_gather_destroy", "carlb_gather_bytes", "carlb_gather_create_symmetric", "carlb_gather_set_mode", "carlb_gather_resync",
]


class NativeLibraryError(RuntimeError):
    pass


class EnvInfo(Structure):
    _fields_ = [
        ("kind", c_int), ("state_words", c_int), ("obs_dim", c_int), ("act_dim", c_int),
        ("act_discrete", c_int), ("n_actions", c_int), ("n_param_rows", c_int), ("n_step_rows", c_int),
        ("default_max_steps", c_int), ("gym_reset_draws", c_int), ("act_low", c_float), ("act_high", c_float),
    ]


class Buffers(Structure):
    _fields_ = [
        ("state", c_void_p), ("ctx", c_void_p), ("elapsed", c_void_p), ("sbt", c_void_p), ("rng", c_void_p),
        ("obs", c_void_p), ("reward", c_void_p), ("terminated", c_void_p), ("truncated", c_void_p),
        ("final_obs", c_void_p), ("first_state", c_void_p), ("first_obs", c_void_p), ("act_staging", c_void_p),
    ]


class Traj(Structure):
    _fields_ = [("obs", c_void_p), ("actions", c_void_p), ("reward", c_void_p), ("done", c_void_p)]


_lib = None


def load() -> ctypes.CDLL:
    """Load libcarlb.so (built in-tree by ``carl_b200.build`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing. Build it with `python -m carl_b200.build` (needs nvcc). "
            "carl_b200 has no CPU fallback."
        )
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise NativeLibraryError(f"could not load {LIB_PATH}: {e}") from e
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise NativeLibraryError(f"{LIB_PATH} does not export {name}")
    lib.carlb_last_error.restype = c_char_p
    lib.carlb_launch_count.restype = c_int64
    lib.carlb_query_env.argtypes = [c_int, POINTER(EnvInfo)]
    lib.carlb_env_create.argtypes = [c_int, c_int, c_int, c_int, c_int64, POINTER(c_void_p)]
    lib.carlb_env_destroy.argtypes = [c_void_p]
    lib.carlb_env_bind.argtypes = [c_void_p, POINTER(Buffers)]
    lib.carlb_env_configure.argtypes = [c_void_p, c_int, c_int]
    lib.carlb_env_seed.argtypes = [c_void_p, c_uint64, c_void_p]
    lib.carlb_env_reset.argtypes = [c_void_p, c_void_p, c_void_p]
    lib.carlb_env_step.argtypes = [c_void_p, c_void_p, c_int, c_void_p]
    lib.carlb_env_step_host.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.carlb_env_step_host_checked.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                                c_void_p]
    lib.carlb_env_step_host_begin.argtypes = [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_void_p]
    lib.carlb_env_step_host_end.argtypes = [c_void_p, c_int]
    lib.carlb_stage_actions.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_int]
    lib.carlb_env_rollout.argtypes = [c_void_p, c_int, c_uint64, c_uint32, c_void_p, c_int, POINTER(Traj), c_void_p]
    lib.carlb_mixed_step.argtypes = [POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int), c_int, c_void_p]
    lib.carlb_brax_set_system.argtypes = [c_void_p, c_void_p, c_int, c_int]
    lib.carlb_brax_set_arithmetic.argtypes = [c_void_p, c_int]
    lib.carlb_brax_set_reset_rng.argtypes = [c_void_p, c_int, c_int64]
    lib.carlb_brax_reset_from_q.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.carlb_brax_goal_step.argtypes = [c_void_p, c_int, c_int, ctypes.c_double, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p]
    lib.carlb_gather_create.argtypes = [c_int, c_int, c_int, c_int64, c_int, POINTER(c_void_p)]
    lib.carlb_gather_export.argtypes = [c_void_p, c_void_p]
    lib.carlb_gather_open.argtypes = [c_void_p, c_int, c_void_p]
    lib.carlb_gather_attach.argtypes = [c_void_p, c_void_p]
    lib.carlb_gather_wait.argtypes = [c_void_p, c_int, c_void_p, POINTER(c_void_p)]
    lib.carlb_gather_destroy.argtypes = [c_void_p]
    lib.carlb_gather_bytes.argtypes = [c_int64, c_int]
    lib.carlb_gather_bytes.restype = c_int64
    lib.carlb_gather_create_symmetric.argtypes = [c_int, c_int, c_int, c_int64, c_int, POINTER(c_void_p), c_void_p, POINTER(c_void_p)]
    lib.carlb_gather_set_mode.argtypes = [c_void_p, c_int]
    lib.carlb_gather_resync.argtypes = [c_void_p, c_void_p]
    if lib.carlb_abi_version() != 1:
        raise NativeLibraryError(f"libcarlb ABI version {lib.carlb_abi_version()} != 1")
    _lib = lib
    return lib


def last_error() -> str:
    return load().carlb_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map a libcarlb return code onto the Python exception types the reference raises."""
    if rc == CARLB_OK:
        return
    msg = load().carlb_last_error().decode("utf-8", "replace")
    if rc == ERR_INVALID:
        raise ValueError(msg)
    raise RuntimeError(msg)


def query_env(kind: int) -> EnvInfo:
    info = EnvInfo()
    check(load().carlb_query_env(kind, ctypes.byref(info)))
    return info


def launch_count() -> int:
    return int(load().carlb_launch_count())
