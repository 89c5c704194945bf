"""Writes tests/golden/gymnasium_known_answers.json.

gymnasium is not installable here, so these are *published* known answers (gymnasium documentation / widely
reproduced outputs): CartPole-v1 reset and step, Pendulum-v1 and MountainCar-v0 reset. The reset vectors are
additionally re-derived below with NumPy alone, which validates the seeding recipe
`np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))` of gymnasium's
`utils.seeding.np_random` that the oracle and the device RNG both have to reproduce.
"""
import json
import os

import numpy as np

OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "gymnasium_known_answers.json")

PUBLISHED = {
    # CartPole-v1: env.reset(seed=0) / env.reset(seed=42) observations (float32)
    "cartpole_reset_seed0": [0.01369617, -0.02302133, -0.04590265, -0.04834724],
    "cartpole_reset_seed42": [0.0273956, -0.00611216, 0.03585979, 0.0197368],
    # env.reset(seed=0); env.step(1) observation, default parameters
    "cartpole_seed0_step_action1": [0.01323574, 0.17272775, -0.04686959, -0.3551522],
    # Pendulum-v1: env.reset(seed=0) / env.reset(seed=42) observations (cos th, sin th, thdot) as the gymnasium
    # documentation prints them: pins gymnasium's own reset -- ONE uniform call over (theta, thetadot) with
    # high = (pi, 1) -- i.e. the two draws CARLPendulum.reset discards before its own (carl_pendulum.py:47-60)
    "pendulum_reset_seed0": [0.6520163, 0.758205, -0.46042657],
    "pendulum_reset_seed42": [-0.14995256, 0.9886932, -0.12224312],
    # MountainCar-v0: env.reset(seed=42) observation: one draw U(-0.6, -0.4), velocity 0 (the draw CARLMountainCar.reset
    # discards, carl_mountaincar.py:59-80)
    "mountaincar_reset_seed42": [-0.4452088, 0.0],
}


def main():
    g = dict(PUBLISHED)
    for seed, key in ((0, "cartpole_reset_seed0"), (42, "cartpole_reset_seed42")):
        gen = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        v = gen.uniform(-0.05, 0.05, 4).astype(np.float32)
        assert np.allclose(v, np.asarray(PUBLISHED[key], dtype=np.float32), rtol=0, atol=1e-8), (seed, v)
        if seed == 0:
            # what CARLCartPole.reset(seed=0) returns on a fresh env: the NEXT four draws at U(-0.1, 0.1)
            g["carl_cartpole_reset_seed0_float64"] = [float(x) for x in gen.uniform(-0.1, 0.1, 4)]
    for seed, key in ((0, "pendulum_reset_seed0"), (42, "pendulum_reset_seed42")):
        gen = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        th, thd = gen.uniform(low=[-np.pi, -1.0], high=[np.pi, 1.0])
        v = np.array([np.cos(th), np.sin(th), thd], dtype=np.float32)
        assert np.allclose(v, np.asarray(PUBLISHED[key], dtype=np.float32), rtol=0, atol=1e-7), (seed, v)
    gen = np.random.Generator(np.random.PCG64(np.random.SeedSequence(42)))
    v = np.array([gen.uniform(low=-0.6, high=-0.4), 0.0], dtype=np.float32)
    assert np.allclose(v, np.asarray(PUBLISHED["mountaincar_reset_seed42"], dtype=np.float32), rtol=0, atol=1e-7), v
    with open(OUT, "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", os.path.abspath(OUT))


if __name__ == "__main__":
    main()
