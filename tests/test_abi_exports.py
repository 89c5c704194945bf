"""The C-ABI library loads on a CPU-only box and exports every symbol include/carlb.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "carlb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|int64_t|char\s*\*|void)\s*\*?\s*(carlb_[a-z0-9_]+)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_header_declares_expected_entry_points():
    names = declared_functions()
    for must in ("carlb_env_create", "carlb_env_bind", "carlb_env_seed", "carlb_env_reset", "carlb_env_step",
                 "carlb_env_step_host", "carlb_env_rollout", "carlb_mixed_step", "carlb_gather_create_symmetric",
                 "carlb_last_error", "carlb_query_env"):
        assert must in names


def test_library_exports_every_declared_symbol(native_lib):
    for name in declared_functions():
        assert hasattr(native_lib, name), f"libcarlb.so does not export {name}"


def test_binding_lists_every_declared_symbol():
    from carl_b200 import _native

    assert sorted(_native.EXPORTS) == declared_functions()


def test_query_env_matches_oracle_tables(native_lib):
    from carl_b200 import _native
    from oracle.classic import KINDS

    for kind, info in KINDS.items():
        q = _native.query_env(_native.KIND[kind])
        assert (q.state_words, q.obs_dim, bool(q.act_discrete), q.default_max_steps, q.gym_reset_draws) == (
            info["S"], info["D"], info["discrete"], info["max_steps"], info["gym_draws"])


def test_errors_map_to_python_exceptions(native_lib):
    from carl_b200 import _native

    with pytest.raises(ValueError):
        _native.query_env(99)
    h = ctypes.c_void_p()
    rc = native_lib.carlb_env_create(0, -5, 0, 0, 0, ctypes.byref(h))
    assert rc != 0 and b"n_envs" in native_lib.carlb_last_error() or rc == _native.ERR_CUDA or rc == _native.ERR_INVALID


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under carl_b200/ may reference it."""
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "carl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "hostcheck" in txt and f.endswith(".py"):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_env_construction_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from carl_b200.envs import CARLCartPole

    with pytest.raises(RuntimeError, match="CUDA"):
        CARLCartPole()
    with pytest.raises(ValueError, match="no CPU fallback"):
        CARLCartPole(device="cpu")


def test_new_entry_points_reject_null_handles_without_touching_the_gpu(native_lib):
    """carlb_env_step_host_checked / carlb_brax_goal_step validate their arguments before any CUDA call:
    a null handle is CARLB_ERR_INVALID / CARLB_ERR_STATE with a message, never a crash (no compute here)."""
    from carl_b200 import _native

    rc = native_lib.carlb_env_step_host_checked(None, None, _native.ACT_I32, 2, None, None, None, None, None)
    assert rc in (_native.ERR_INVALID, _native.ERR_STATE) and native_lib.carlb_last_error()
    rc = native_lib.carlb_brax_goal_step(None, 0, 1, ctypes.c_double(0.01), None, None, None, None, None, None)
    assert rc in (_native.ERR_INVALID, _native.ERR_STATE) and native_lib.carlb_last_error()
    for kind, (D, A) in {"brax_inverted_pendulum": (4, 1), "brax_inverted_double_pendulum": (8, 1), "brax_reacher": (11, 2),
                         "brax_humanoid": (244, 17), "brax_humanoidstandup": (244, 17), "brax_pusher": (23, 7)}.items():
        q = _native.query_env(_native.KIND[kind])
        assert (q.obs_dim, q.act_dim, q.default_max_steps) == (D, A, 1000)
