"""Loader for real-Brax golden dumps (tools/gen_brax_golden.py). The dumps cannot be produced in the
build container (brax/jax absent): these tests skip until ``tests/golden/brax/<env>.npz`` exists, and
then compare the CUDA path with real Brax without any code change (tunables come from the dump)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "brax")
BODIES = {"ant": "CARLBraxAnt", "halfcheetah": "CARLBraxHalfcheetah", "hopper": "CARLBraxHopper",
          "walker2d": "CARLBraxWalker2d", "inverted_pendulum": "CARLBraxInvertedPendulum",
          "inverted_double_pendulum": "CARLBraxInvertedDoublePendulum", "reacher": "CARLBraxReacher",
          "humanoid": "CARLBraxHumanoid", "humanoidstandup": "CARLBraxHumanoidStandup", "pusher": "CARLBraxPusher"}


def _load(body):
    p = os.path.join(GOLD, f"{body}.npz")
    if not os.path.exists(p):
        pytest.skip(f"no real-Brax golden dump at {p} (parity unpinned; see tools/gen_brax_golden.py)")
    return np.load(p)


@pytest.mark.parametrize("body", list(BODIES))
def test_system_constants_match_dump(body):
    d = _load(body)
    from carl_b200.envs import brax_system as bs

    s = bs.SYSTEMS[body]
    np.testing.assert_allclose(s["stock_masses"], d["sys_link_mass"], rtol=1e-5)
    np.testing.assert_allclose(s["table"][bs.OFF_INIT_Q:bs.OFF_INIT_Q + s["n_q"]], d["sys_init_q"], atol=1e-6)
    assert s["table"][bs.H_DT] == pytest.approx(float(d["sys_dt"]))


@pytest.mark.gpu
@pytest.mark.parametrize("body", list(BODIES))
def test_cuda_path_matches_real_brax_trajectory(body):
    d = _load(body)
    import torch

    import carl_b200.envs as E

    tun = {k: float(np.asarray(d[f"sys_{k}"]).reshape(-1)[-1]) for k in (
        "constraint_stiffness", "constraint_vel_damping", "constraint_limit_stiffness", "constraint_ang_damping",
        "baumgarte_erp", "vel_damping", "spring_mass_scale", "spring_inertia_scale")}
    env = getattr(E, BODIES[body])(num_envs=1, brax_tunables=tun, autoreset=False)
    q, qd = d["q0"][None].astype(np.float32), d["qd0"][None].astype(np.float32)
    obs, _ = env.reset_from_q(q, qd)
    np.testing.assert_allclose(obs["obs"].cpu().numpy()[0], d["obs0"], rtol=1e-4, atol=1e-4)
    for t, a in enumerate(d["actions"]):
        if t > 0:  # teacher forcing from the dumped generalized coordinates
            env.reset_from_q(d["q"][t - 1][None].astype(np.float32), d["qd"][t - 1][None].astype(np.float32))
        obs, r, te, tr, _ = env.step(torch.from_numpy(a[None]).cuda())
        np.testing.assert_allclose(obs["obs"].cpu().numpy()[0], d["obs"][t], rtol=1e-3, atol=1e-3)
        assert bool(te.item()) == bool(d["done"][t])
