"""Independent anchors for the UNPINNED restatements (VERDICT r01 "next" #1).

gymnasium 0.29.1 / brax 0.12.1 cannot be installed here or on the GPU box (profiles/r02a_pip_install_attempt.txt),
so Pendulum / Acrobot / MountainCar have no reference-run golden vectors. What can be checked without the
reference is whether the restated arithmetic obeys the physics it claims to integrate -- facts that are NOT
encoded in the code and that a transcription error (a wrong sign, a dropped Coriolis term, a wrong inertia
factor) breaks:

* Pendulum: small-oscillation period of a uniform rod about its hanging position, T = 2 pi sqrt(2 l / (3 g));
  energy E = thdot^2 / 2 + (3 g / 2 l) cos(th) of the torque-free pendulum drifts O(dt) (semi-implicit Euler).
* Acrobot ("book" dynamics): one env step equals one RK4 step of the textbook two-link manipulator equations
  (full mass matrix, Coriolis / centrifugal and gravity vectors written from the Lagrangian, torque on joint 2)
  to 1e-10; the torque-free step conserves the two-link Lagrangian energy to O(dt^5).
* MountainCar: the step is symplectic Euler on H = v^2/2 + gravity sin(3p)/3; near the valley bottom
  (p* = -pi/6) the discrete period is 2 pi / acos(1 - 3 gravity / 2) steps; H stays bounded.
* MountainCarContinuous: same hill with the hard-coded 0.0025.

Each anchor is evaluated on the float64 oracle AND on the product's kernel source compiled by g++
(tests/hostcheck), so both sides of the parity tests are tied to something outside themselves.
"""
import numpy as np
import pytest

from oracle.classic import DEFAULTS, FEATURES, KINDS, OracleClassicEnv
from tests.util import HostCheck, kernel_rows


@pytest.fixture(scope="module")
def hc():
    return HostCheck()


class _Stepper:
    """Uniform step(state, action) -> state over the oracle or the kernel source (float64 mode)."""

    def __init__(self, kind, table, impl, hc):
        self.kind, self.impl, self.hc = kind, impl, hc
        self.n = table.shape[0]
        self.table = table
        if impl == "oracle":
            self.ora = OracleClassicEnv(kind, table, max_steps=10**9)
        else:
            self.rows = kernel_rows(kind, table, "reference", np.float64)
            self.rngs = np.zeros((4, self.n), dtype=np.uint64)
            self.sbt = np.zeros(self.n, dtype=np.uint8)
            self.el = np.zeros(self.n, dtype=np.int32)

    def step(self, state, action):
        if self.impl == "oracle":
            self.ora.state[:] = state
            self.ora.elapsed[:] = 0
            self.ora.step(action)
            return self.ora.state.copy()
        st = np.ascontiguousarray(state, dtype=np.float64).copy()
        self.el[:] = 0
        self.hc.step(self.kind, True, st, self.rows, action, self.rngs, self.sbt, self.el, 10**9, 0)
        return st


def _table(kind, n=1, **over):
    t = np.tile(np.asarray(DEFAULTS[kind], dtype=np.float64), (n, 1))
    for k, v in over.items():
        t[:, FEATURES[kind].index(k)] = v
    return t


def _period_from_crossings(x, dt):
    """Mean period from the upward zero crossings of a sampled oscillation (linear interpolation)."""
    idx = np.nonzero((x[:-1] < 0) & (x[1:] >= 0))[0]
    t = (idx + (-x[idx]) / (x[idx + 1] - x[idx])) * dt
    assert len(t) >= 3
    return float(np.mean(np.diff(t)))


IMPLS = ["oracle", "kernel_source"]


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("g,l", [(10.0, 1.0), (5.0, 2.0), (15.0, 0.5)])
def test_pendulum_small_oscillation_period_of_a_uniform_rod(hc, impl, g, l):
    dt = 1e-3
    st = _Stepper("pendulum", _table("pendulum", g=g, l=l, m=1.3, dt=dt), impl, hc)
    s = np.array([[np.pi + 0.01, 0.0]])  # theta = pi is hanging down (theta = 0 upright)
    zero = np.zeros(1, dtype=np.float32)
    xs = []
    for _ in range(int(3.5 * 2 * np.pi * np.sqrt(2 * l / (3 * g)) / dt)):
        s = st.step(s, zero)
        xs.append(s[0, 0] - np.pi)
    T = _period_from_crossings(np.array(xs), dt)
    assert T == pytest.approx(2 * np.pi * np.sqrt(2 * l / (3 * g)), rel=2e-3)


@pytest.mark.parametrize("impl", IMPLS)
def test_pendulum_energy_drift_is_first_order_in_dt(hc, impl):
    """Torque-free: E = thdot^2/2 + (3g/2l) cos(th). Semi-implicit Euler keeps it within O(dt) of its initial value."""
    g, l = 9.0, 1.2
    errs = []
    for dt in (2e-3, 1e-3):
        st = _Stepper("pendulum", _table("pendulum", g=g, l=l, dt=dt), impl, hc)
        s = np.array([[2.0, 0.5]])
        e0 = 0.5 * s[0, 1] ** 2 + 1.5 * g / l * np.cos(s[0, 0])
        worst = 0.0
        for _ in range(int(2.0 / dt)):
            s = st.step(s, np.zeros(1, dtype=np.float32))
            worst = max(worst, abs(0.5 * s[0, 1] ** 2 + 1.5 * g / l * np.cos(s[0, 0]) - e0))
        errs.append(worst)
    assert errs[0] < 0.05 and errs[1] < 0.6 * errs[0]  # halving dt halves the drift


def _pendulum_torque_response(hc, impl):
    """One step from rest at the bottom with torque u: thdot = 3 u dt / (m l^2) (rod inertia m l^2 / 3)."""
    m, l, dt, u = 1.7, 0.8, 0.01, 1.5
    st = _Stepper("pendulum", _table("pendulum", g=0.0, m=m, l=l, dt=dt), impl, hc)
    s = st.step(np.array([[np.pi, 0.0]]), np.array([u], dtype=np.float32))
    return s[0, 1], 3 * u * dt / (m * l * l)


@pytest.mark.parametrize("impl", IMPLS)
def test_pendulum_torque_accelerates_a_rod_of_inertia_ml2_over_3(hc, impl):
    got, want = _pendulum_torque_response(hc, impl)
    assert got == pytest.approx(want, rel=1e-12)


# ---------------------------------------------------------------------------------- Acrobot
def _acrobot_energy(s, m1, m2, l1, lc1, lc2, moi):
    th1, th2, w1, w2 = s
    g = 9.8
    d11 = m1 * lc1**2 + m2 * (l1**2 + lc2**2 + 2 * l1 * lc2 * np.cos(th2)) + 2 * moi
    d12 = m2 * (lc2**2 + l1 * lc2 * np.cos(th2)) + moi
    d22 = m2 * lc2**2 + moi
    kin = 0.5 * d11 * w1**2 + d12 * w1 * w2 + 0.5 * d22 * w2**2
    pot = -(m1 * lc1 + m2 * l1) * g * np.cos(th1) - m2 * lc2 * g * np.cos(th1 + th2)
    return kin + pot


@pytest.mark.parametrize("impl", IMPLS)
def test_acrobot_rk4_step_conserves_the_two_link_lagrangian_energy(hc, impl):
    """Zero torque, moderate speeds: one classic RK4 step of 0.2 s has local error O(dt^5), so the textbook
    energy of the two-link pendulum (mass matrix d11/d12/d22, gravity potential) moves by < 1e-3 relative; a
    dropped Coriolis term or a wrong sign in phi1 / phi2 moves it by O(1)."""
    rng = np.random.default_rng(0)
    n = 256
    p = dict(m1=rng.uniform(0.5, 2, n), m2=rng.uniform(0.5, 2, n), l1=rng.uniform(0.5, 2, n), lc1=rng.uniform(0.3, 0.7, n),
             lc2=rng.uniform(0.3, 0.7, n), moi=rng.uniform(0.5, 2, n))
    t = _table("acrobot", n, LINK_MASS_1=p["m1"], LINK_MASS_2=p["m2"], LINK_LENGTH_1=p["l1"], LINK_COM_POS_1=p["lc1"],
               LINK_COM_POS_2=p["lc2"], LINK_MOI=p["moi"], MAX_VEL_1=1e6, MAX_VEL_2=1e6)
    st = _Stepper("acrobot", t, impl, hc)
    s = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(-0.5, 0.5, n), rng.uniform(-0.5, 0.5, n)], 1)
    e0 = _acrobot_energy(s.T, **p)
    s1 = st.step(s, np.ones(n, dtype=np.int32))  # action 1 = zero torque
    e1 = _acrobot_energy(s1.T, **p)
    scale = np.abs(e0) + 1.0
    assert np.max(np.abs(e1 - e0) / scale) < 1e-3


def _manipulator_rhs(y, torque, m1, m2, l1, lc1, lc2, moi):
    """Textbook two-link manipulator equations M(q) qdd + C(q, qd) + G(q) = (0, torque), angles from the hanging
    position (Spong 1995; Sutton & Barto's acrobot) -- written from the Lagrangian, NOT from `_dsdt`'s eliminated form."""
    th1, th2, w1, w2 = y
    g = 9.8
    d11 = m1 * lc1**2 + m2 * (l1**2 + lc2**2 + 2 * l1 * lc2 * np.cos(th2)) + 2 * moi
    d12 = m2 * (lc2**2 + l1 * lc2 * np.cos(th2)) + moi
    d22 = m2 * lc2**2 + moi
    h = m2 * l1 * lc2 * np.sin(th2)
    c1 = -h * (2 * w1 * w2 + w2**2)
    c2 = h * w1**2
    g1 = (m1 * lc1 + m2 * l1) * g * np.sin(th1) + m2 * lc2 * g * np.sin(th1 + th2)
    g2 = m2 * lc2 * g * np.sin(th1 + th2)
    b1, b2 = -c1 - g1, torque - c2 - g2
    det = d11 * d22 - d12 * d12
    return np.array([w1, w2, (d22 * b1 - d12 * b2) / det, (d11 * b2 - d12 * b1) / det])


@pytest.mark.parametrize("impl", IMPLS)
def test_acrobot_step_is_rk4_of_the_textbook_manipulator_equations(hc, impl):
    """One env step == one classic RK4 step (dt = 0.2) of the manipulator equations solved with the full 2x2 mass
    matrix, for random link parameters and all three torques -- to 1e-10. The restated `_dsdt` reaches the same
    accelerations through the book's eliminated form (d1, d2, phi1, phi2); agreeing with the Lagrangian form
    pins its signs, its Coriolis / centrifugal terms, the gravity terms and which joint the motor drives."""
    rng = np.random.default_rng(1)
    n = 512
    p = dict(m1=rng.uniform(0.5, 2, n), m2=rng.uniform(0.5, 2, n), l1=rng.uniform(0.5, 2, n), lc1=rng.uniform(0.3, 0.7, n),
             lc2=rng.uniform(0.3, 0.7, n), moi=rng.uniform(0.5, 2, n))
    t = _table("acrobot", n, LINK_MASS_1=p["m1"], LINK_MASS_2=p["m2"], LINK_LENGTH_1=p["l1"], LINK_COM_POS_1=p["lc1"],
               LINK_COM_POS_2=p["lc2"], LINK_MOI=p["moi"], MAX_VEL_1=1e6, MAX_VEL_2=1e6)
    st = _Stepper("acrobot", t, impl, hc)
    s = np.stack([rng.uniform(-1.5, 1.5, n), rng.uniform(-1.5, 1.5, n), rng.uniform(-2, 2, n), rng.uniform(-3, 3, n)], 1)
    act = rng.integers(0, 3, size=n).astype(np.int32)
    torque = act.astype(np.float64) - 1.0
    got = st.step(s, act)
    y0, dt = s.T.copy(), 0.2
    f = lambda y: _manipulator_rhs(y, torque, **p)
    k1 = f(y0); k2 = f(y0 + dt / 2 * k1); k3 = f(y0 + dt / 2 * k2); k4 = f(y0 + dt * k3)
    want = (y0 + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)).T
    inside = (np.abs(want[:, 0]) < np.pi) & (np.abs(want[:, 1]) < np.pi)  # the env wraps angles beyond +-pi
    assert inside.sum() > n // 2
    np.testing.assert_allclose(got[inside], want[inside], rtol=1e-10, atol=1e-10)


# ------------------------------------------------------------------------------- MountainCar
@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("kind,grav", [("mountaincar", 0.0025), ("mountaincar", 0.004), ("mountaincar_cont", 0.0025)])
def test_mountaincar_valley_oscillation_period(hc, impl, kind, grav):
    """No engine (action 1 / force 0): v += -gravity cos(3p); p += v is symplectic Euler on the hill
    U(p) = gravity sin(3p)/3. Around the valley bottom p* = -pi/6, U'' = 3 gravity, so the discrete map has the
    period 2 pi / acos(1 - 3 gravity / 2) steps (72.5 for the stock gravity)."""
    over = {} if kind == "mountaincar_cont" else {"gravity": grav}
    st = _Stepper(kind, _table(kind, **over), impl, hc)
    s = np.array([[-np.pi / 6 + 0.01, 0.0]])
    act = np.ones(1, dtype=np.int32) if kind == "mountaincar" else np.zeros(1, dtype=np.float32)
    xs = []
    for _ in range(400):
        s = st.step(s, act)
        xs.append(s[0, 0] + np.pi / 6)
    T = _period_from_crossings(np.array(xs), 1.0)
    assert T == pytest.approx(2 * np.pi / np.arccos(1 - 1.5 * grav), rel=2e-3)


@pytest.mark.parametrize("impl", IMPLS)
def test_mountaincar_engine_force_and_hill_energy(hc, impl):
    """With the engine pushing right (action 2) the modified energy H = v^2/2 + g sin(3p)/3 - force p of the
    symplectic map stays within O(step) of its start as long as no clip is hit."""
    force, grav = 0.001, 0.0025
    st = _Stepper("mountaincar", _table("mountaincar", force=force, gravity=grav), impl, hc)
    s = np.array([[-0.5, 0.0]])
    H = lambda q: 0.5 * q[0, 1] ** 2 + grav * np.sin(3 * q[0, 0]) / 3 - force * q[0, 0]
    h0, worst = H(s), 0.0
    for _ in range(60):
        s = st.step(s, np.full(1, 2, dtype=np.int32))
        assert -1.2 < s[0, 0] < 0.6 and abs(s[0, 1]) < 0.07
        worst = max(worst, abs(H(s) - h0))
    assert worst < 2e-5  # ~ |v| * |dv| per step; a sign error in either term gives > 1e-3
