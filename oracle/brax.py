"""TEST INFRASTRUCTURE — CPU oracle wrapper for the Brax-locomotion path (see brax_oracle.c).

PARITY UNPINNED: brax/jax are not installable here and the reference pins no numbers for this path;
``tools/gen_brax_golden.py`` dumps trajectories wherever real Brax exists so that
``tests/test_brax_golden.py`` can close the gap without code changes.
"""
from __future__ import annotations

import ctypes

import numpy as np

from oracle import lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class OracleBraxEnv:
    """N independent envs of one body; ``sysd`` is an entry of carl_b200.envs.brax_system.SYSTEMS
    (the packed table is a shared *data format*; the stepping code is oracle/brax_oracle.c)."""

    def __init__(self, sysd: dict, ctx: np.ndarray, max_steps: int = 1000, autoreset: bool = True, f64: bool = False):
        """f64=True runs the same restatement in float64: the round-off yardstick used to tell
        float32 noise (amplified by the stiff joint springs) from algorithmic differences."""
        self.sysd = sysd
        self.f64 = bool(f64)
        self.real = np.float64 if f64 else np.float32
        self.table = np.ascontiguousarray(sysd["table"], dtype=np.float32)
        self.ctx = np.ascontiguousarray(ctx, dtype=np.float32)
        self.n = self.ctx.shape[0]
        self.words = sysd["state_words"]
        self.D = sysd["obs_dim"]
        self.max_steps, self.autoreset = int(max_steps), bool(autoreset)
        self.state = np.zeros((self.n, self.words), dtype=self.real)
        self.first_state = np.zeros_like(self.state)
        self.first_obs = np.zeros((self.n, self.D), dtype=np.float32)
        self.elapsed = np.zeros(self.n, dtype=np.int32)

    def init_from_q(self, q: np.ndarray, qd: np.ndarray) -> np.ndarray:
        q = np.ascontiguousarray(q, dtype=np.float32)
        qd = np.ascontiguousarray(qd, dtype=np.float32)
        obs = np.zeros((self.n, self.D), dtype=np.float32)
        getattr(lib(), 'brax_oracle_init64' if self.f64 else 'brax_oracle_init')(_p(self.table), self.n, _p(q), _p(qd), _p(self.state), self.words, _p(obs), self.D,
            _p(self.ctx), self.ctx.shape[1])
        self.first_state[:] = self.state
        self.first_obs[:] = obs
        self.elapsed[:] = 0
        return obs

    def step(self, actions: np.ndarray):
        a = np.ascontiguousarray(actions, dtype=np.float32)
        obs = np.zeros((self.n, self.D), dtype=np.float32)
        final = np.zeros((self.n, self.D), dtype=np.float32)
        reward = np.zeros(self.n, dtype=np.float32)
        done = np.zeros(self.n, dtype=np.uint8)
        getattr(lib(), 'brax_oracle_step64' if self.f64 else 'brax_oracle_step')(_p(self.table), self.n, _p(self.state), self.words, _p(self.ctx), self.ctx.shape[1], _p(a),
                               _p(self.elapsed), self.max_steps, int(self.autoreset), _p(self.first_state),
                               _p(self.first_obs), _p(obs), self.D, _p(reward), _p(done), _p(final))
        return obs, reward, done.astype(bool), final
