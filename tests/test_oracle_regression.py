"""The oracle restatements must not drift between rounds: replay the committed self-regression
fixtures (tests/golden/oracle_regression.json, made by tools/make_oracle_regression_fixtures.py).
These pin OUR restatement, not the reference (which pins nothing here)."""
import json
import os

import numpy as np
import pytest

from carl_b200.envs import brax_system as bs
from oracle.brax import OracleBraxEnv
from oracle.classic import DEFAULTS, KINDS, OracleClassicEnv

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_regression.json")))


@pytest.mark.parametrize("kind", list(KINDS))
def test_classic_oracle_replays_fixture(kind):
    fx = G["classic"][kind]
    env = OracleClassicEnv(kind, np.array([DEFAULTS[kind]], dtype=np.float64))
    np.testing.assert_array_equal(env.reset(seed=2024)[0], np.asarray(fx["obs0"], dtype=np.float32))
    for st in fx["steps"]:
        a = np.array([int(st["action"])]) if KINDS[kind]["discrete"] else np.array([st["action"]], dtype=np.float32)
        o, r, te, tr, _ = env.step(a)
        np.testing.assert_allclose(o[0], st["obs"], rtol=1e-6, atol=1e-7)
        assert r[0] == pytest.approx(st["reward"], rel=1e-9, abs=1e-12) and bool(te[0]) == st["terminated"]
    np.testing.assert_allclose(env.state[0], fx["state"], rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("body", ["ant", "halfcheetah", "hopper", "walker2d", "inverted_pendulum", "inverted_double_pendulum", "reacher"])
def test_brax_oracle_replays_fixture(body):
    fx = G["brax"][body]
    sysd = bs.SYSTEMS[body]
    ctx = np.array([[sysd["stock_gravity"], -1, -1, sysd["stock_ang_damping"], 1.0, *sysd["stock_masses"]]], dtype=np.float32)
    for f64, tol in ((True, 1e-7), (False, 2e-4)):
        env = OracleBraxEnv(sysd, ctx, autoreset=False, f64=f64)
        o0 = env.init_from_q(np.asarray([fx["q"]], np.float32), np.asarray([fx["qd"]], np.float32))
        np.testing.assert_allclose(o0[0], fx["obs0"], rtol=1e-5, atol=1e-6)
        for a, st in zip(fx["actions"], fx["steps"]):
            o, r, d, _ = env.step(np.asarray([a], np.float32))
            scale = max(1.0, np.abs(st["obs"]).max())
            assert np.abs(o[0] - np.asarray(st["obs"])).max() <= tol * scale * 10
            assert bool(d[0]) == st["done"]
