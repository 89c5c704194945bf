"""Shared helpers of the test-suite (test infrastructure; may use the oracle)."""
from __future__ import annotations

import ctypes

import numpy as np

from oracle.classic import DEFAULTS, FEATURES, KINDS

KIND_ID = {"cartpole": 0, "pendulum": 1, "acrobot": 2, "mountaincar": 3, "mountaincar_cont": 4}
ENV_CLASS = {
    "cartpole": "CARLCartPole", "pendulum": "CARLPendulum", "acrobot": "CARLAcrobot",
    "mountaincar": "CARLMountainCar", "mountaincar_cont": "CARLMountainCarContinuous",
}


def env_class(kind):
    import carl_b200.envs as E

    return getattr(E, ENV_CLASS[kind])


def sample_context_table(kind: str, n: int, rng: np.random.Generator, f32: bool = True) -> np.ndarray:
    """Random contexts inside the reference's feature bounds (moderate ranges)."""
    names = FEATURES[kind]
    t = np.tile(np.asarray(DEFAULTS[kind], dtype=np.float64), (n, 1))

    def setcol(name, lo, hi):
        t[:, names.index(name)] = rng.uniform(lo, hi, n)

    if kind == "cartpole":
        setcol("gravity", 5, 15); setcol("masscart", 0.5, 2.0); setcol("masspole", 0.05, 0.5)
        setcol("length", 0.25, 1.0); setcol("force_mag", 5, 20); setcol("tau", 0.01, 0.04)
        setcol("initial_state_lower", -0.2, -0.05); setcol("initial_state_upper", 0.05, 0.2)
    elif kind == "pendulum":
        setcol("g", 5, 15); setcol("m", 0.5, 2.0); setcol("l", 0.5, 2.0); setcol("dt", 0.02, 0.08)
        setcol("initial_angle_max", 1.0, np.pi); setcol("initial_velocity_max", 0.5, 2.0)
    elif kind == "acrobot":
        setcol("LINK_LENGTH_1", 0.5, 2.0); setcol("LINK_MASS_1", 0.5, 2.0); setcol("LINK_MASS_2", 0.5, 2.0)
        setcol("LINK_COM_POS_1", 0.3, 0.7); setcol("LINK_COM_POS_2", 0.3, 0.7); setcol("LINK_MOI", 0.5, 2.0)
        setcol("INITIAL_ANGLE_LOWER", -0.3, -0.05); setcol("INITIAL_ANGLE_UPPER", 0.05, 0.3)
    elif kind == "mountaincar":
        setcol("force", 0.0005, 0.002); setcol("gravity", 0.0015, 0.0035); setcol("goal_position", 0.3, 0.55)
        setcol("max_speed", 0.05, 0.09); setcol("min_velocity_start", -0.01, 0.0); setcol("max_velocity_start", 0.0, 0.01)
    else:
        setcol("power", 0.001, 0.002); setcol("goal_position", 0.3, 0.55); setcol("max_speed", 0.05, 0.09)
    if f32:
        t = t.astype(np.float32).astype(np.float64)
    return t


def sample_states(kind: str, n: int, rng: np.random.Generator, f32: bool = True) -> np.ndarray:
    """States from each env's reachable box (some beyond the termination thresholds)."""
    if kind == "cartpole":
        s = np.stack([rng.uniform(-2.6, 2.6, n), rng.uniform(-3, 3, n), rng.uniform(-0.25, 0.25, n), rng.uniform(-3, 3, n)], 1)
    elif kind == "pendulum":
        s = np.stack([rng.uniform(-10, 10, n), rng.uniform(-8, 8, n)], 1)
    elif kind == "acrobot":
        s = np.stack([rng.uniform(-np.pi, np.pi, n), rng.uniform(-np.pi, np.pi, n), rng.uniform(-12, 12, n), rng.uniform(-28, 28, n)], 1)
    else:
        s = np.stack([rng.uniform(-1.2, 0.6, n), rng.uniform(-0.07, 0.07, n)], 1)
    if f32:
        s = s.astype(np.float32).astype(np.float64)
    return s


def sample_actions(kind: str, n: int, rng: np.random.Generator) -> np.ndarray:
    info = KINDS[kind]
    if info["discrete"]:
        return rng.integers(0, 2 if kind == "cartpole" else 3, size=n).astype(np.int32)
    half = 2.5 if kind == "pendulum" else 1.2  # beyond the clip range on purpose
    return rng.uniform(-half, half, size=n).astype(np.float32)


def kernel_rows(kind: str, table: np.ndarray, mode: str = "reference", dtype=np.float64) -> np.ndarray:
    """Per-env kernel-parameter table [P][n] as the host layer uploads it."""
    cls = env_class(kind)
    rows = cls.kernel_params(table, FEATURES[kind], mode)
    return np.ascontiguousarray(rows.T.astype(dtype))


def done_margin(kind: str, state_after: np.ndarray, table: np.ndarray) -> np.ndarray:
    """Distance of each env from its nearest termination threshold (for mismatch reports)."""
    if kind == "cartpole":
        thr = 12 * 2 * np.pi / 360
        return np.minimum(np.abs(np.abs(state_after[:, 0]) - 2.4), np.abs(np.abs(state_after[:, 2]) - thr))
    if kind == "acrobot":
        return np.abs(-np.cos(state_after[:, 0]) - np.cos(state_after[:, 1] + state_after[:, 0]) - 1.0)
    if kind in ("mountaincar", "mountaincar_cont"):
        gp = table[:, FEATURES[kind].index("goal_position")]
        return np.abs(state_after[:, 0] - gp)
    return np.full(len(state_after), np.inf)


class HostCheck:
    """ctypes wrapper of tests/hostcheck (the product's __host__ __device__ physics compiled by g++)."""

    def __init__(self):
        from tests.hostcheck.build_hostcheck import build

        self.lib = ctypes.CDLL(build())

    @staticmethod
    def _p(a):
        return None if a is None else a.ctypes.data_as(ctypes.c_void_p)

    def seed_rng(self, n: int, seed: int, offset: int = 0) -> np.ndarray:
        rng = np.zeros((4, n), dtype=np.uint64)
        out = (ctypes.c_uint64 * 4)()
        for i in range(n):
            self.lib.hc_pcg64_seed(ctypes.c_uint64(seed + offset + i), out)
            rng[:, i] = list(out)
        return rng

    def reset(self, kind, f64, state, rows, rng):
        n = state.shape[0]
        obs = np.zeros((n, KINDS[kind]["D"]), dtype=np.float32)
        rc = self.lib.hc_reset(KIND_ID[kind], int(f64), n, self._p(state), self._p(rows), self._p(rng), self._p(obs))
        assert rc == 0
        return obs

    def step(self, kind, f64, state, rows, actions, rng, sbt, elapsed, max_steps, autoreset):
        n = state.shape[0]
        D = KINDS[kind]["D"]
        obs = np.zeros((n, D), dtype=np.float32)
        final = np.zeros((n, D), dtype=np.float32)
        reward = np.zeros(n, dtype=np.float32)
        term = np.zeros(n, dtype=np.uint8)
        trunc = np.zeros(n, dtype=np.uint8)
        ai = af = None
        if KINDS[kind]["discrete"]:
            ai = np.ascontiguousarray(actions, dtype=np.int32)
        else:
            af = np.ascontiguousarray(actions, dtype=np.float32)
        rc = self.lib.hc_step(KIND_ID[kind], int(f64), n, self._p(state), self._p(rows), self._p(ai), self._p(af),
                              self._p(rng), self._p(sbt), self._p(elapsed), int(max_steps), int(autoreset),
                              self._p(obs), self._p(reward), self._p(term), self._p(trunc), self._p(final))
        assert rc == 0
        return obs, reward, term.astype(bool), trunc.astype(bool), final
