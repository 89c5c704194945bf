"""Pins the CPU oracle against every golden vector the reference path offers."""
import json
import os

import numpy as np

from oracle.classic import DEFAULTS, FEATURES, KINDS, OracleClassicEnv

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "gymnasium_known_answers.json")))


def test_seeding_recipe_reproduces_published_reset_vectors():
    for seed, key in ((0, "cartpole_reset_seed0"), (42, "cartpole_reset_seed42")):
        g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        got = g.uniform(-0.05, 0.05, 4).astype(np.float32)
        np.testing.assert_allclose(got, np.asarray(GOLD[key], dtype=np.float32), rtol=0, atol=1e-8)


def test_published_pendulum_and_mountaincar_resets_pin_the_discarded_gymnasium_draws():
    """Every CARL reset first lets gymnasium's own reset draw (and discards the result): Pendulum-v1 draws (theta,
    thetadot) in ONE uniform call with high = (pi, 1), MountainCar-v0 one U(-0.6, -0.4). The published reset
    observations pin those draw counts and ranges; the oracle's CARL state is what the SAME generator yields next."""
    for seed, key in ((0, "pendulum_reset_seed0"), (42, "pendulum_reset_seed42")):
        g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        th, thd = g.uniform(low=[-np.pi, -1.0], high=[np.pi, 1.0])
        np.testing.assert_allclose(np.array([np.cos(th), np.sin(th), thd], np.float32), np.asarray(GOLD[key], np.float32),
                                   rtol=0, atol=1e-7)
        assert KINDS["pendulum"]["gym_draws"] == 2
        env = OracleClassicEnv("pendulum", np.array([DEFAULTS["pendulum"]]))
        env.reset(seed=seed)
        want = np.array([g.uniform(high=np.pi), g.uniform(high=1.0)], dtype=np.float32)  # CARL's own two draws follow
        np.testing.assert_array_equal(env.state[0], want.astype(np.float64))
    g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(42)))
    pos = g.uniform(low=-0.6, high=-0.4)
    np.testing.assert_allclose(np.array([pos, 0.0], np.float32), np.asarray(GOLD["mountaincar_reset_seed42"], np.float32),
                               rtol=0, atol=1e-7)
    assert KINDS["mountaincar"]["gym_draws"] == 1
    env = OracleClassicEnv("mountaincar", np.array([DEFAULTS["mountaincar"]]))
    env.reset(seed=42)
    np.testing.assert_array_equal(env.state[0], [g.uniform(low=-0.6, high=-0.4), g.uniform(low=0.0, high=0.0)])


def test_oracle_cartpole_step_known_answer():
    env = OracleClassicEnv("cartpole", np.array([DEFAULTS["cartpole"]]))
    g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(0)))
    env.state[0] = g.uniform(-0.05, 0.05, 4)
    obs, r, term, trunc, _ = env.step(np.array([1]))
    np.testing.assert_allclose(obs[0], np.asarray(GOLD["cartpole_seed0_step_action1"], dtype=np.float32), rtol=0, atol=1e-8)
    assert r[0] == 1.0 and not term[0] and not trunc[0]


def test_oracle_carl_reset_draws_after_gymnasium_draws():
    """CARLCartPole.reset(seed=0): gymnasium's four draws are discarded, CARL's four follow."""
    env = OracleClassicEnv("cartpole", np.array([DEFAULTS["cartpole"]]))
    obs = env.reset(seed=0)
    np.testing.assert_array_equal(env.state[0], np.asarray(GOLD["carl_cartpole_reset_seed0_float64"]))
    np.testing.assert_array_equal(obs[0], np.asarray(GOLD["carl_cartpole_reset_seed0_float64"], dtype=np.float32))


def test_oracle_masscart_is_inert_in_reference_mode():
    """CartPoleEnv caches total_mass; setattr('masscart') changes nothing."""
    c0 = np.array([DEFAULTS["cartpole"]])
    c1 = c0.copy()
    c1[0, FEATURES["cartpole"].index("masscart")] = 5.0
    s = np.array([0.01, 0.2, 0.03, -0.1])
    outs = []
    for c, applied in ((c0, False), (c1, False), (c1, True)):
        e = OracleClassicEnv("cartpole", c, applied_mode=applied)
        e.state[0] = s
        outs.append(e.step(np.array([1]))[0][0])
    np.testing.assert_array_equal(outs[0], outs[1])
    assert not np.array_equal(outs[0], outs[2])


def test_oracle_time_limit_and_steps_beyond():
    e = OracleClassicEnv("cartpole", np.array([DEFAULTS["cartpole"]]), max_steps=5)
    e.state[0] = [2.39, 5.0, 0.0, 0.0]  # leaves the track at once
    rs, ts, trs = [], [], []
    for _ in range(6):
        _, r, t, tr, _ = e.step(np.array([1]))
        rs.append(r[0]); ts.append(bool(t[0])); trs.append(bool(tr[0]))
    assert ts == [True] * 6
    assert rs == [1.0, 0.0, 0.0, 0.0, 0.0, 0.0]  # reward 1 on the terminating step, 0 afterwards
    assert trs == [False, False, False, False, True, True]


def test_oracle_shapes_all_kinds():
    rng = np.random.default_rng(0)
    for kind, info in KINDS.items():
        e = OracleClassicEnv(kind, np.tile(np.asarray(DEFAULTS[kind], dtype=np.float64), (3, 1)))
        o = e.reset(seed=5)
        assert o.shape == (3, info["D"]) and o.dtype == np.float32
        a = rng.integers(0, 2, 3) if info["discrete"] else rng.uniform(-1, 1, 3).astype(np.float32)
        o, r, t, tr, _ = e.step(a)
        assert o.shape == (3, info["D"]) and r.shape == (3,) and np.isfinite(o).all()


def test_oracle_pendulum_properties():
    """No golden vectors exist for Pendulum (parity unpinned): check invariants instead."""
    e = OracleClassicEnv("pendulum", np.array([DEFAULTS["pendulum"]]))
    e.state[0] = [np.pi, 0.0]  # hanging... upright is 0; at pi cost is pi^2
    o, r, t, tr, _ = e.step(np.array([0.0], dtype=np.float32))
    assert abs(r[0] + np.pi**2) < 1e-12 and not t[0]
    assert abs(o[0, 0] ** 2 + o[0, 1] ** 2 - 1) < 1e-6
    e.state[0] = [0.0, 100.0]
    o, *_ = e.step(np.array([5.0], dtype=np.float32))
    assert o[0, 2] == 8.0  # speed clip


def test_oracle_mountaincar_wall_and_goal():
    e = OracleClassicEnv("mountaincar", np.array([DEFAULTS["mountaincar"]]))
    e.state[0] = [-1.2, -0.07]
    o, r, t, tr, _ = e.step(np.array([0]))
    assert o[0, 0] == np.float32(-1.2) and o[0, 1] == 0.0 and r[0] == -1.0
    e.state[0] = [0.449, 0.07]
    o, r, t, tr, _ = e.step(np.array([2]))
    assert t[0]  # CARL default goal_position 0.45 (carl_mountaincar.py:27-29)
