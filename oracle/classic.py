"""TEST INFRASTRUCTURE — reference-faithful CPU oracle for the classic-control envs.

``OracleClassicEnv`` models N *independent* copies of the reference's single env
(``CARL<Env>(contexts={0: ctx_i})`` with a static context), each with its own
``np.random.Generator(PCG64(SeedSequence(seed_i)))`` exactly as gymnasium's
``Env.reset(seed=...)`` creates it (reached from carl/envs/carl_env.py:271), the
double draw of every ``CARL*.reset`` (gymnasium's discarded draws first, then CARL's:
carl_cartpole.py:50-61, carl_pendulum.py:47-60, carl_acrobot.py:77-100,
carl_mountaincar.py:59-80, carl_mountaincarcontinuous.py:56-77), gymnasium's TimeLimit
and the float64 step arithmetic in ``classic_oracle.c``.

Not part of the product; see ``oracle/__init__.py``.
"""
from __future__ import annotations

import ctypes

import numpy as np

from oracle import lib

# name, state dim, obs dim, discrete?, TimeLimit, gymnasium's own reset draws (discarded)
KINDS = {
    "cartpole": dict(S=4, D=4, discrete=True, max_steps=500, gym_draws=4),
    "pendulum": dict(S=2, D=3, discrete=False, max_steps=200, gym_draws=2),
    "acrobot": dict(S=4, D=6, discrete=True, max_steps=500, gym_draws=4),
    "mountaincar": dict(S=2, D=2, discrete=True, max_steps=200, gym_draws=1),
    "mountaincar_cont": dict(S=2, D=2, discrete=False, max_steps=999, gym_draws=1),
}

# context feature order = get_context_features() order in the reference files
FEATURES = {
    "cartpole": ["gravity", "masscart", "masspole", "length", "force_mag", "tau",
                 "initial_state_lower", "initial_state_upper"],
    "pendulum": ["gravity", "dt", "g", "m", "l", "initial_angle_max", "initial_velocity_max"],
    "acrobot": ["LINK_LENGTH_1", "LINK_LENGTH_2", "LINK_MASS_1", "LINK_MASS_2", "LINK_COM_POS_1",
                "LINK_COM_POS_2", "LINK_MOI", "MAX_VEL_1", "MAX_VEL_2", "torque_noise_max",
                "INITIAL_ANGLE_LOWER", "INITIAL_ANGLE_UPPER", "INITIAL_VELOCITY_LOWER",
                "INITIAL_VELOCITY_UPPER"],
    "mountaincar": ["min_position", "max_position", "max_speed", "goal_position", "goal_velocity",
                    "force", "gravity", "min_position_start", "max_position_start",
                    "min_velocity_start", "max_velocity_start"],
    "mountaincar_cont": ["min_position", "max_position", "max_speed", "goal_position",
                         "goal_velocity", "power", "min_position_start", "max_position_start",
                         "min_velocity_start", "max_velocity_start"],
}

DEFAULTS = {
    "cartpole": [9.8, 1.0, 0.1, 0.5, 10.0, 0.02, -0.1, 0.1],
    "pendulum": [8.0, 0.05, 10, 1, 1, np.pi, 1],
    "acrobot": [1, 1, 1, 1, 0.5, 0.5, 1, 4 * np.pi, 9 * np.pi, 0, -0.1, 0.1, -0.1, 0.1],
    "mountaincar": [-1.2, 0.6, 0.07, 0.45, 0, 0.001, 0.0025, -0.6, -0.4, 0, 0],
    "mountaincar_cont": [-1.2, 0.6, 0.07, 0.5, 0, 0.0015, -0.6, -0.4, 0, 0],
}


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


class OracleClassicEnv:
    def __init__(self, kind: str, ctx: np.ndarray, applied_mode: bool = False, max_steps: int | None = None):
        self.kind = kind
        self.info = KINDS[kind]
        self.ctx = np.ascontiguousarray(ctx, dtype=np.float64)
        assert self.ctx.ndim == 2 and self.ctx.shape[1] == len(FEATURES[kind])
        self.n = self.ctx.shape[0]
        self.applied_mode = bool(applied_mode)
        self.max_steps = self.info["max_steps"] if max_steps is None else int(max_steps)
        self.state = np.zeros((self.n, self.info["S"]), dtype=np.float64)
        self.elapsed = np.zeros(self.n, dtype=np.int32)
        self.sbt = np.zeros(self.n, dtype=np.uint8)
        self.rngs: list[np.random.Generator | None] = [None] * self.n
        self._col = {k: j for j, k in enumerate(FEATURES[kind])}

    def c(self, i, name):
        return float(self.ctx[i, self._col[name]])

    # ------------------------------------------------------------------ reset
    def _reset_one(self, i: int) -> np.ndarray:
        g = self.rngs[i]
        k = self.kind
        # gymnasium's own reset draws (discarded by CARL's override)
        g.uniform(size=self.info["gym_draws"])
        if k == "cartpole":
            st = g.uniform(low=self.c(i, "initial_state_lower"), high=self.c(i, "initial_state_upper"), size=(4,))
            self.state[i] = st
            obs = np.array(st, dtype=np.float32)
        elif k == "pendulum":
            theta = g.uniform(high=self.c(i, "initial_angle_max"))
            thetadot = g.uniform(high=self.c(i, "initial_velocity_max"))
            self.state[i] = np.array([theta, thetadot], dtype=np.float32)
            obs = np.array([np.cos(theta), np.sin(theta), thetadot], dtype=np.float32)
        elif k == "acrobot":
            ang = g.uniform(low=self.c(i, "INITIAL_ANGLE_LOWER"), high=self.c(i, "INITIAL_ANGLE_UPPER"), size=(2,))
            vel = g.uniform(low=self.c(i, "INITIAL_VELOCITY_LOWER"), high=self.c(i, "INITIAL_VELOCITY_UPPER"), size=(2,))
            st = np.concatenate([ang, vel])
            self.state[i] = st
            obs = np.array([np.cos(st[0]), np.sin(st[0]), np.cos(st[1]), np.sin(st[1]), st[2], st[3]], dtype=np.float32)
        else:
            pos = g.uniform(low=self.c(i, "min_position_start"), high=self.c(i, "max_position_start"))
            vel = g.uniform(low=self.c(i, "min_velocity_start"), high=self.c(i, "max_velocity_start"))
            self.state[i] = [pos, vel]
            obs = np.array([pos, vel], dtype=np.float32)
        self.elapsed[i] = 0
        self.sbt[i] = 0
        return obs

    def reset(self, seed: int | None = None, mask: np.ndarray | None = None) -> np.ndarray:
        """Env i is (re)seeded with ``seed + i`` (gymnasium vector-env convention) when a seed is
        given, otherwise its generator continues."""
        obs = np.zeros((self.n, self.info["D"]), dtype=np.float32)
        for i in range(self.n):
            if mask is not None and not mask[i]:
                continue
            if seed is not None:
                self.rngs[i] = np.random.Generator(np.random.PCG64(np.random.SeedSequence(int(seed) + i)))
            assert self.rngs[i] is not None, "reset(seed=...) first"
            obs[i] = self._reset_one(i)
        return obs

    # ------------------------------------------------------------------- step
    def step(self, action: np.ndarray, autoreset: bool = False):
        n, L = self.n, lib()
        D = self.info["D"]
        obs = np.zeros((n, D), dtype=np.float32)
        reward = np.zeros(n, dtype=np.float64)
        term = np.zeros(n, dtype=np.uint8)
        trunc = np.zeros(n, dtype=np.uint8)
        dbl, u8, i32, f32 = ctypes.c_double, ctypes.c_ubyte, ctypes.c_int, ctypes.c_float
        common = (_p(self.elapsed, i32),)
        tail = (_p(obs, f32), _p(reward, dbl), _p(term, u8), _p(trunc, u8))
        k = self.kind
        if self.info["discrete"]:
            act = np.ascontiguousarray(action, dtype=np.int32).reshape(n)
        else:
            act = np.ascontiguousarray(action, dtype=np.float32).reshape(n)
        if k == "cartpole":
            L.oracle_cartpole_step(n, _p(self.state, dbl), _p(self.ctx, dbl), _p(act, i32), *common, _p(self.sbt, u8),
                                   self.max_steps, int(self.applied_mode), *tail)
        elif k == "pendulum":
            L.oracle_pendulum_step(n, _p(self.state, dbl), _p(self.ctx, dbl), _p(act, f32), *common, self.max_steps, *tail)
        elif k == "acrobot":
            noise = np.zeros(n, dtype=np.float64)
            for i in range(n):
                tnm = self.c(i, "torque_noise_max")
                if tnm > 0:
                    noise[i] = self.rngs[i].uniform(-tnm, tnm)
            L.oracle_acrobot_step(n, _p(self.state, dbl), _p(self.ctx, dbl), _p(act, i32), _p(noise, dbl), *common,
                                  self.max_steps, *tail)
        elif k == "mountaincar":
            L.oracle_mountaincar_step(n, _p(self.state, dbl), _p(self.ctx, dbl), _p(act, i32), *common, self.max_steps, *tail)
        else:
            L.oracle_mountaincar_cont_step(n, _p(self.state, dbl), _p(self.ctx, dbl), _p(act, f32), *common,
                                           self.max_steps, *tail)
        final_obs = None
        if autoreset:
            done = (term | trunc).astype(bool)
            if done.any():
                final_obs = obs.copy()
                for i in np.nonzero(done)[0]:
                    obs[i] = self._reset_one(int(i))
        return obs, reward, term.astype(bool), trunc.astype(bool), final_obs


def scalar_python_cartpole_steps_per_s(n_steps: int = 100_000, seed: int = 0) -> float:
    """The reference's actual cost shape: one env, scalar Python float64 math, TimeLimit
    bookkeeping and a fresh ``{"obs","context"}`` dict per step (carl_env.py:295-305,339-342).
    Timed for context next to the compiled oracle; gymnasium itself is absent."""
    import math
    import time

    ctx = dict(zip(FEATURES["cartpole"], DEFAULTS["cartpole"]))
    obs_features = list(ctx)
    rng = np.random.default_rng(seed)
    acts = rng.integers(0, 2, size=n_steps).tolist()
    state = list(rng.uniform(-0.1, 0.1, 4))
    total_mass, polemass_length = 1.1, 0.05
    thr = 12 * 2 * math.pi / 360
    elapsed = 0
    t0 = time.perf_counter()
    for a in acts:
        x, x_dot, theta, theta_dot = state
        force = ctx["force_mag"] if a == 1 else -ctx["force_mag"]
        costheta, sintheta = math.cos(theta), math.sin(theta)
        temp = (force + polemass_length * theta_dot**2 * sintheta) / total_mass
        thetaacc = (ctx["gravity"] * sintheta - costheta * temp) / (
            ctx["length"] * (4.0 / 3.0 - ctx["masspole"] * costheta**2 / total_mass))
        xacc = temp - polemass_length * thetaacc * costheta / total_mass
        x = x + ctx["tau"] * x_dot
        x_dot = x_dot + ctx["tau"] * xacc
        theta = theta + ctx["tau"] * theta_dot
        theta_dot = theta_dot + ctx["tau"] * thetaacc
        state = [x, x_dot, theta, theta_dot]
        terminated = bool(x < -2.4 or x > 2.4 or theta < -thr or theta > thr)
        elapsed += 1
        truncated = elapsed >= 500
        obs = {"obs": np.array(state, dtype=np.float32),
               "context": {k: v for k, v in ctx.items() if k in obs_features}}
        info = {"context_id": 0}
        if terminated or truncated:
            state = list(rng.uniform(-0.1, 0.1, 4))
            elapsed = 0
    dt = time.perf_counter() - t0
    del obs, info
    return n_steps / dt
