"""Regression fixtures of the ORACLE ITSELF (not reference goldens -- the reference pins nothing for
these paths): short trajectories of the float64 restatements from fixed inputs, committed under
tests/golden/oracle_regression.json so that an accidental change of the restated algorithm in a later
round is caught by the CPU suite.

    python tools/make_oracle_regression_fixtures.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from carl_b200.envs import brax_system as bs
from oracle.brax import OracleBraxEnv
from oracle.classic import DEFAULTS, KINDS, OracleClassicEnv

OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "oracle_regression.json")


def classic(kind):
    env = OracleClassicEnv(kind, np.array([DEFAULTS[kind]], dtype=np.float64))
    obs0 = env.reset(seed=2024)
    rng = np.random.default_rng(7)
    traj = []
    for t in range(6):
        a = rng.integers(0, 2 if kind == "cartpole" else 3, 1) if KINDS[kind]["discrete"] else rng.uniform(-1, 1, 1).astype(np.float32)
        o, r, te, tr, _ = env.step(a)
        traj.append(dict(action=float(a[0]), obs=[float(x) for x in o[0]], reward=float(r[0]), terminated=bool(te[0])))
    return dict(obs0=[float(x) for x in obs0[0]], steps=traj, state=[float(x) for x in env.state[0]])


def brax(body):
    sysd = bs.SYSTEMS[body]
    ctx = np.array([[sysd["stock_gravity"], -1, -1, sysd["stock_ang_damping"], 1.0, *sysd["stock_masses"]]], dtype=np.float32)
    env = OracleBraxEnv(sysd, ctx, autoreset=False, f64=True)
    rng = np.random.default_rng(11)
    init_q = sysd["table"][bs.OFF_INIT_Q:bs.OFF_INIT_Q + sysd["n_q"]]
    q = (init_q[None] + rng.uniform(-0.05, 0.05, (1, sysd["n_q"]))).astype(np.float32)
    qd = (0.05 * rng.standard_normal((1, sysd["n_qd"]))).astype(np.float32)
    obs0 = env.init_from_q(q, qd)
    acts = rng.uniform(-1, 1, (5, 1, sysd["n_act"])).astype(np.float32)
    traj = []
    for a in acts:
        o, r, d, _ = env.step(a)
        traj.append(dict(obs=[float(x) for x in o[0]], reward=float(r[0]), done=bool(d[0])))
    return dict(q=[float(x) for x in q[0]], qd=[float(x) for x in qd[0]], actions=acts[:, 0].tolist(),
                obs0=[float(x) for x in obs0[0]], steps=traj)


def main():
    g = {"note": "oracle self-regression fixtures (float64 restatements); NOT reference goldens",
         "classic": {k: classic(k) for k in KINDS}, "brax": {b: brax(b) for b in ("ant", "halfcheetah", "hopper", "walker2d", "inverted_pendulum", "inverted_double_pendulum", "reacher")}}
    with open(OUT, "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", os.path.abspath(OUT))


if __name__ == "__main__":
    main()
