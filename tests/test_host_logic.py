"""Host-side logic that needs no GPU: kernel-parameter tables (batched _update_context),
sharding arithmetic, and the N>1 gather path over gloo (world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest

from carl_b200.envs import CARLAcrobot, CARLCartPole, CARLMountainCar, CARLMountainCarContinuous, CARLPendulum
from carl_b200.envs.brax import CARLBraxAnt, CARLBraxHalfcheetah, CARLBraxHopper, check_context
from carl_b200.parallel import host_gather_reference, shard_range, shard_sizes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _table(cls, **over):
    d = cls.get_context_space().get_default_context()
    d.update(over)
    names = list(d)
    return np.array([[float(d[n]) for n in names]]), names


def test_cartpole_reference_mode_keeps_stale_masses():
    t, names = _table(CARLCartPole, masscart=4.0, masspole=0.3, length=0.8)
    ref = CARLCartPole.kernel_params(t, names, "reference")[0]
    app = CARLCartPole.kernel_params(t, names, "applied")[0]
    assert ref[5] == pytest.approx(1.1) and ref[6] == pytest.approx(0.05)  # total_mass, polemass_length
    assert app[5] == pytest.approx(4.3) and app[6] == pytest.approx(0.24)
    assert ref[1] == app[1] == 0.3 and ref[2] == app[2] == 0.8  # masspole / length still act directly


def test_pendulum_gravity_feature_is_dead_in_reference_mode():
    t, names = _table(CARLPendulum, gravity=3.0)
    assert CARLPendulum.kernel_params(t, names, "reference")[0, 0] == 10.0
    assert CARLPendulum.kernel_params(t, names, "applied")[0, 0] == 3.0


def test_pendulum_applied_mode_honours_a_deliberate_gravity_of_8():
    """`gravity` = 8.0 is the feature's default; a context that names it explicitly must still win
    over `g` in applied mode (the `explicit` mask says which features each context set itself)."""
    t, names = _table(CARLPendulum, gravity=8.0)
    j = names.index("gravity")
    explicit = np.zeros((1, len(names)), dtype=bool)
    assert CARLPendulum.kernel_params(t, names, "applied", explicit=explicit)[0, 0] == 10.0   # filled-in default: g stays
    explicit[0, j] = True
    assert CARLPendulum.kernel_params(t, names, "applied", explicit=explicit)[0, 0] == 8.0    # chosen on purpose
    assert CARLPendulum.kernel_params(t, names, "reference", explicit=explicit)[0, 0] == 10.0  # dead in the reference


def test_param_row_counts_match_kernel_tables(native_lib):
    from carl_b200 import _native

    for cls in (CARLCartPole, CARLPendulum, CARLAcrobot, CARLMountainCar, CARLMountainCarContinuous):
        t, names = _table(cls)
        assert cls.kernel_params(t, names).shape[1] == _native.query_env(_native.KIND[cls.kind]).n_param_rows


def test_feature_tables_match_reference_goldens():
    import json

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "notebook_goldens.json")))
    cs = CARLBraxAnt.get_context_space()
    assert cs.context_feature_names == g["ant_feature_names"]
    assert cs.get_default_context() == g["ant_default_context"]
    assert list(cs.get_lower_and_upper_bound("friction")) == g["ant_friction_bounds"]
    assert "target_distance" not in CARLBraxAnt.get_default_context()
    assert "target_distance" in CARLBraxAnt.get_default_goal_context()
    assert list(CARLBraxHalfcheetah.get_default_context())[-1] == "mass_ffoot"
    assert CARLBraxHopper.get_default_context()["mass_foot"] == 5.3155746


def test_check_context_rejects_unknown_features():
    check_context({"gravity": 1, "mass_torso": 2}, ["gravity"])
    with pytest.raises(RuntimeError):
        check_context({"bogus": 1}, ["gravity"])


@pytest.mark.parametrize("n,w", [(65536, 8), (8192, 3), (10, 4), (7, 7)])
def test_shard_range_partitions(n, w):
    spans = [shard_range(n, r, w) for r in range(w)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = shard_sizes(n, w)
    assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(n, w, w)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from types import SimpleNamespace
from carl_b200.parallel import ObsGather, shard_range, host_gather_reference
from carl_b200.context import RoundRobinSelector
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
N, D = {n}, 3
lo, hi = shard_range(N, rank, world)
full = np.arange(N * D, dtype=np.float32).reshape(N, D)
env = SimpleNamespace(world_size=world, global_num_envs=N, device=torch.device("cpu"),
                      _info=SimpleNamespace(obs_dim=D), _obs=torch.from_numpy(full[lo:hi].copy()))
g = ObsGather(env, mode="nccl")
out = g.gather().numpy()
assert np.array_equal(out, full), (rank, out)
# context ids do not depend on sharding: every rank consumes the selector for the whole batch
sel = RoundRobinSelector({{i: {{}} for i in range(5)}})
ids = sel.select_batch(N)[lo:hi]
assert np.array_equal(ids, (np.arange(N) % 5)[lo:hi])
dist.barrier()
dist.destroy_process_group()
print("OK", rank)
"""


@pytest.mark.parametrize("n", [8, 7])
def test_gather_world_size_2_gloo(tmp_path, n):
    """N>1 host path on CPU: two gloo ranks, contiguous shards (equal and ragged), gathered obs
    equals the rank-order concatenation."""
    import subprocess

    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, n=n))
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out.decode()
    assert np.array_equal(host_gather_reference([np.ones((2, 3)), np.zeros((1, 3))]).shape, (3, 3))


def test_legacy_aliases_and_extension_features():
    """BASELINE's `torso_mass` / `joint_stiffness` (stale-docs names) map onto
    `mass_torso` and a per-env joint-constraint-stiffness scale."""
    names = list(CARLBraxAnt.get_context_space().get_default_context()) + ["joint_stiffness"]
    d = dict(CARLBraxAnt.get_context_space().get_default_context(), joint_stiffness=1.0)
    t = np.array([[float(d[n]) for n in names]])
    t[0, names.index("joint_stiffness")] = 1.7
    rows = CARLBraxAnt.kernel_params(t, names, "applied")
    assert rows.shape[1] == 5 + 9 and rows[0, 4] == 1.7
    assert CARLBraxAnt.kernel_params(t, names, "reference")[0, 4] == 1.0
    assert CARLBraxAnt.feature_aliases["torso_mass"] == "mass_torso"
    assert "joint_stiffness" not in CARLBraxAnt.get_context_features()  # the context space stays the reference's


def test_registration_is_optional():
    from carl_b200.registration import ENV_NAMES, register_envs
    import carl_b200.envs as E

    ids = register_envs()
    assert ids == [] or len(ids) == len(ENV_NAMES)
    assert all(hasattr(E, n) for n in ENV_NAMES)


def test_pinned_registry_and_check_only_staging():
    """hostmem's address registry (what lets `step` read an action array in place) and the range-check-only
    mode of carlb_stage_actions; the page-locked allocation itself needs the driver, so a plain tensor
    stands in for the block here."""
    import ctypes

    import torch

    from carl_b200 import _native, hostmem

    t = torch.zeros(4096, dtype=torch.uint8)
    hostmem._register(t)
    a = t.numpy().view(np.int32).reshape(4, 256)
    assert hostmem.is_pinned(a[1].ctypes.data, a[1].nbytes)
    assert not hostmem.is_pinned(a[3].ctypes.data, a[3].nbytes + 4)       # runs past the end of the block
    assert not hostmem.is_pinned(np.zeros(8).ctypes.data, 64)
    lib = _native.load()
    a[:] = 1
    assert lib.carlb_stage_actions(None, a[2].ctypes.data, 256, _native.ACT_I32, 2) == 0
    a[2, 17] = -3
    assert lib.carlb_stage_actions(None, a[2].ctypes.data, 256, _native.ACT_I32, 2) == _native.ERR_INVALID
    assert "invalid action" in _native.last_error()
    assert lib.carlb_stage_actions(None, a[2].ctypes.data, 256, _native.ACT_I32, 0) == 0   # validation off
    dst = np.zeros(256, dtype=np.int32)
    a[2, 17] = 0
    assert lib.carlb_stage_actions(dst.ctypes.data, a[2].ctypes.data, 256, _native.ACT_I32, 2) == 0
    assert (dst == a[2]).all()
    hostmem.release(a)
    assert not hostmem.is_pinned(a[1].ctypes.data, a[1].nbytes)


def test_brax_param_rows_and_shapes_match_native_query(native_lib):
    """Every Brax class (incl. the inverted pendulums and the reacher): host-side kernel_params rows, obs / action
    dims and the action range agree with what the C ABI reports for the kind."""
    import carl_b200.envs as E
    from carl_b200 import _native
    from carl_b200.envs import brax_system as bs

    for name in ("CARLBraxAnt", "CARLBraxHalfcheetah", "CARLBraxHopper", "CARLBraxWalker2d", "CARLBraxInvertedPendulum",
                 "CARLBraxInvertedDoublePendulum", "CARLBraxReacher", "CARLBraxHumanoid", "CARLBraxHumanoidStandup",
                 "CARLBraxPusher"):
        cls = getattr(E, name)
        info = _native.query_env(_native.KIND[cls.kind])
        sysd = bs.SYSTEMS[cls.env_name]
        d = cls.get_default_context()
        names = list(d)
        t = np.array([[float(d[n]) for n in names]])
        for mode in ("reference", "applied"):
            assert cls.kernel_params(t, names, mode).shape == (1, info.n_param_rows)
        assert (info.obs_dim, info.act_dim, info.state_words) == (sysd["obs_dim"], sysd["n_act"], sysd["state_words"])
        assert info.act_high == sysd["act_scale"] and info.act_low == -sysd["act_scale"]
        applied = cls.kernel_params(t, names, "applied")
        for j, ln in enumerate(sysd["link_names"]):
            want = d.get(f"mass_{ln}", sysd["stock_masses"][j])
            assert applied[0, 5 + j] == pytest.approx(want)
    assert E.brax.UNSUPPORTED_BODIES == ()  # every body of carl/envs/brax/__init__.py is built
    # the pusher's goal features stay in the context but never reach the family's check_context / the physics
    d = E.CARLBraxPusher.get_context_space().get_default_context()
    assert [d[k] for k in E.CARLBraxPusher.GOAL_FEATURES] == [0.45, 0.05, 0.05]
