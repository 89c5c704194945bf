"""Brax test helpers (test infrastructure)."""
from __future__ import annotations

import ctypes

import numpy as np

from carl_b200.envs import brax_system as bs


def random_ctx(sysd, n, rng, applied=True):
    """Per-env kernel context rows: gravity, friction, elasticity, ang_damping, link masses."""
    L = sysd["n_links"]
    ctx = np.zeros((n, 5 + L), dtype=np.float32)
    if applied:
        ctx[:, 0] = rng.uniform(-15, -5, n)
        ctx[:, 1] = rng.uniform(0.5, 1.5, n)
        ctx[:, 2] = rng.uniform(0.0, 0.3, n)
        ctx[:, 3] = rng.uniform(-0.1, 0.0, n)
        ctx[:, 4] = rng.uniform(0.5, 1.5, n)  # joint-stiffness scale
        ctx[:, 5:] = np.asarray(sysd["stock_masses"])[None] * rng.uniform(0.5, 2.0, (n, L))
    else:
        ctx[:, 0] = sysd["stock_gravity"]
        ctx[:, 1] = -1.0
        ctx[:, 2] = -1.0
        ctx[:, 3] = sysd["stock_ang_damping"]
        ctx[:, 4] = 1.0
        ctx[:, 5:] = np.asarray(sysd["stock_masses"])[None]
    return ctx


def random_q(sysd, n, rng, scale=1.0):
    nq, nqd = sysd["n_q"], sysd["n_qd"]
    init_q = sysd["table"][bs.OFF_INIT_Q:bs.OFF_INIT_Q + nq].astype(np.float32)
    q = init_q[None] + rng.uniform(-0.1 * scale, 0.1 * scale, (n, nq)).astype(np.float32)
    qd = (0.1 * scale * rng.standard_normal((n, nqd))).astype(np.float32)
    return q.astype(np.float32), qd


def pusher_contact_states(sysd, n, rng):
    """Pusher states in which the gripper is down at the table next to the ball (the arm lowered by the shoulder-lift
    joint, the ball placed within 0.12 of the wrist-roll link's centre of mass): roughly a quarter of them have an
    active capsule-vs-ball contact pair, which generic random states around the initial pose never reach."""
    from oracle.brax import OracleBraxEnv

    nq = sysd["n_q"]
    q = np.zeros((n, nq), np.float32)
    qd = np.zeros((n, sysd["n_qd"]), np.float32)
    q[:, 0] = rng.uniform(-0.5, 0.8, n)
    q[:, 1] = rng.uniform(0.30, 0.42, n)
    q[:, 2:7] = rng.uniform(-0.1, 0.1, (n, 5))
    q[:, 3] -= 0.15
    q[:, 5] -= 0.1
    qd[:, :7] = rng.uniform(-1, 1, (n, 7))
    ora = OracleBraxEnv(sysd, np.ones((n, 5 + sysd["n_links"]), np.float32), autoreset=False)
    ora.init_from_q(q, qd)
    wrist = ora.state[:, :13 * sysd["n_links"]].reshape(n, sysd["n_links"], 13)[:, 6, :3]
    ang, dist = rng.uniform(0, 2 * np.pi, n), rng.uniform(0.0, 0.12, n)
    q[:, 8] = wrist[:, 0] + dist * np.cos(ang) - 0.45   # slide along x (second coordinate of the object)
    q[:, 7] = wrist[:, 1] + dist * np.sin(ang) + 0.05   # slide along y
    return q, qd


class BraxHostCheck:
    def __init__(self):
        from tests.hostcheck.build_hostcheck import build

        self.lib = ctypes.CDLL(build())

    @staticmethod
    def _p(a):
        return None if a is None else a.ctypes.data_as(ctypes.c_void_p)

    def init(self, sysd, q, qd, ctx=None):
        n = q.shape[0]
        if ctx is None:  # stock masses (they enter the humanoid's observation)
            ctx = np.zeros((n, 5 + sysd["n_links"]), dtype=np.float32)
            ctx[:, 5:] = np.asarray(sysd["stock_masses"], dtype=np.float32)[None]
        ctx = np.ascontiguousarray(ctx, dtype=np.float32)
        state = np.zeros((n, sysd["state_words"]), dtype=np.float32)
        obs = np.zeros((n, sysd["obs_dim"]), dtype=np.float32)
        t = np.ascontiguousarray(sysd["table"], dtype=np.float32)
        self.lib.hc_brax_init(self._p(t), n, self._p(np.ascontiguousarray(q)), self._p(np.ascontiguousarray(qd)),
                              self._p(state), sysd["state_words"], self._p(obs), sysd["obs_dim"], self._p(ctx), ctx.shape[1])
        return state, obs

    def step(self, sysd, state, ctx, actions, elapsed, max_steps, autoreset, first_state, first_obs, stock_contact=0, fast=False):
        """fast=True: the reformulated arithmetic of the FMA build (world-frame hinge, unit-inertia shortcut)."""
        n = state.shape[0]
        self.lib.hc_brax_set_fast(1 if fast else 0)
        obs = np.zeros((n, sysd["obs_dim"]), dtype=np.float32)
        reward = np.zeros(n, dtype=np.float32)
        done = np.zeros(n, dtype=np.uint8)
        t = np.ascontiguousarray(sysd["table"], dtype=np.float32)
        self.lib.hc_brax_step(self._p(t), n, self._p(state), sysd["state_words"], self._p(ctx), ctx.shape[1],
                              self._p(np.ascontiguousarray(actions, dtype=np.float32)), self._p(elapsed), int(max_steps),
                              int(autoreset), self._p(first_state), self._p(first_obs), self._p(obs), sysd["obs_dim"],
                              self._p(reward), self._p(done), int(stock_contact))
        return obs, reward, done.astype(bool)


def assert_close_scaled(got, want, rel=1e-5, steps=1, what="obs"):
    """North-star tolerance for the float32 Brax path: 1e-5 relative *to the magnitude of the
    env's own vector* (joint velocities reach ~10 while angles are ~0.1: an element-wise rtol on
    near-zero entries would test float32 cancellation noise, not the physics), scaled by the number
    of env-steps compared without re-synchronisation."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = np.maximum(1.0, np.abs(want).reshape(want.shape[0], -1).max(axis=1))
    tol = rel * steps * scale.reshape((-1,) + (1,) * (want.ndim - 1))
    err = np.abs(got - want)
    bad = err > tol
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.size} entries beyond {rel:g} x {steps} x scale; "
                           f"worst err {err.max():.3e} (allowed {tol.reshape(-1)[np.argmax((err / tol).reshape(err.shape[0], -1).max(axis=1))]:.3e})")
