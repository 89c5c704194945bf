"""No-GPU check of the Brax kernel source logic (physics_brax.h through tests/hostcheck) against
the CPU oracle (oracle/brax_oracle.c): forward kinematics, n_frames spring substeps, env layer,
EpisodeWrapper truncation and AutoReset. float32 on both sides; 1e-5 relative per step (north
star), done masks identical."""
import numpy as np
import pytest

from carl_b200.envs import brax_system as bs
from oracle.brax import OracleBraxEnv
from tests.brax_util import BraxHostCheck, assert_close_scaled, random_ctx, random_q

BODIES = ["ant", "halfcheetah", "hopper", "walker2d", "inverted_pendulum", "inverted_double_pendulum", "reacher"]


@pytest.fixture(scope="module")
def hc():
    return BraxHostCheck()


@pytest.mark.parametrize("body", BODIES)
def test_pipeline_init_matches(hc, body):
    sysd = bs.SYSTEMS[body]
    rng = np.random.default_rng(0)
    q, qd = random_q(sysd, 64, rng, scale=3.0)
    ora = OracleBraxEnv(sysd, random_ctx(sysd, 64, rng))
    o_ref = ora.init_from_q(q, qd)
    st, o = hc.init(sysd, q, qd)
    np.testing.assert_allclose(st, ora.state, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(o, o_ref, rtol=1e-5, atol=2e-6)
    # forward then inverse kinematics is the identity on (q[ex:], qd)
    ex = int(sysd["table"][bs.H_EXCLUDE_POS])
    want = np.concatenate([q[:, ex:], qd], axis=1)
    if body == "inverted_double_pendulum":  # _get_obs: q[:1], sin(q[1:]), cos(q[1:]), clip(qd)
        want = np.concatenate([q[:, :1], np.sin(q[:, 1:]), np.cos(q[:, 1:]), np.clip(qd, -10, 10)], axis=1)
    if body == "reacher":  # cos(theta), sin(theta), target, arm qd, fingertip - target (planar 2-link arm)
        th1, th12 = q[:, 0].astype(np.float64), (q[:, 0] + q[:, 1]).astype(np.float64)
        tip = np.stack([0.1 * np.cos(th1) + 0.11 * np.cos(th12), 0.1 * np.sin(th1) + 0.11 * np.sin(th12), 0.01 + 0 * th1], axis=1)
        tgt = np.concatenate([q[:, 2:4], np.full((q.shape[0], 1), 0.01)], axis=1)
        want = np.concatenate([np.cos(q[:, :2]), np.sin(q[:, :2]), q[:, 2:4], qd[:, :2], tip - tgt], axis=1)
    if body == "ant":  # the free root's quaternion is normalised by forward()
        want[:, 1:5] /= np.linalg.norm(want[:, 1:5], axis=1, keepdims=True)
        want[:, 13 + 3:13 + 6] = o_ref[:, 13 + 3:13 + 6]  # angular velocity is reported in the local frame
    clip = sysd["table"][bs.H_QD_CLIP]
    if clip > 0:
        want[:, -sysd["n_qd"]:] = np.clip(want[:, -sysd["n_qd"]:], -clip, clip)
    np.testing.assert_allclose(o_ref, want, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("body", BODIES)
@pytest.mark.parametrize("applied", [True, False])
def test_single_env_step_matches(hc, body, applied):
    """P1/P4: one env-step (n_frames substeps) from random states, contexts and actions."""
    sysd = bs.SYSTEMS[body]
    n = 256
    rng = np.random.default_rng(1)
    ctx = random_ctx(sysd, n, rng, applied=applied)
    q, qd = random_q(sysd, n, rng, scale=2.0)
    ora = OracleBraxEnv(sysd, ctx, autoreset=False)
    ora.init_from_q(q, qd)
    st, _ = hc.init(sysd, q, qd)
    a = rng.uniform(-1.2, 1.2, (n, sysd["n_act"])).astype(np.float32)
    o_ref, r_ref, d_ref, _ = ora.step(a)
    el = np.zeros(n, dtype=np.int32)
    o, r, d = hc.step(sysd, st, ctx, a, el, 1000, 0, st.copy(), o_ref.copy(), stock_contact=0 if applied else 1)
    assert_close_scaled(o, o_ref, rel=1e-5)
    np.testing.assert_allclose(r, r_ref, rtol=1e-4, atol=1e-4)
    assert (d == d_ref).all()
    assert_close_scaled(st, ora.state, rel=1e-5, what="state")


@pytest.mark.parametrize("body", BODIES)
def test_rollout_with_autoreset_matches(hc, body):
    """P2: 60 env-steps with shared random actions, short episodes (truncation -> done ->
    AutoReset to the stored first state); done masks identical, drift bounded."""
    sysd = bs.SYSTEMS[body]
    n, T, max_steps = 32, 60, 25
    rng = np.random.default_rng(2)
    ctx = random_ctx(sysd, n, rng)
    q, qd = random_q(sysd, n, rng)
    ora = OracleBraxEnv(sysd, ctx, max_steps=max_steps, autoreset=True)
    o0 = ora.init_from_q(q, qd)
    st, o = hc.init(sysd, q, qd)
    first_state, first_obs = st.copy(), o.copy()
    el = np.zeros(n, dtype=np.int32)
    n_done = 0
    for t in range(T):
        a = rng.uniform(-1, 1, (n, sysd["n_act"])).astype(np.float32)
        o_ref, r_ref, d_ref, _ = ora.step(a)
        o, r, d = hc.step(sysd, st, ctx, a, el, max_steps, 1, first_state, first_obs)
        assert (d == d_ref).all(), f"done mismatch at step {t}"
        k = (t % max_steps) + 1
        assert_close_scaled(o, o_ref, rel=2e-5, steps=k)
        n_done += int(d.sum())
        assert (el == ora.elapsed).all()
    assert n_done >= n  # every env was truncated at least twice


@pytest.mark.parametrize("body", BODIES)
def test_physical_sanity(body):
    """The restated pipeline is physically sane: from the initial pose under zero action the body
    settles (no blow-up), contacts hold it above the ground, energy does not grow."""
    sysd = bs.SYSTEMS[body]
    ctx = random_ctx(sysd, 1, np.random.default_rng(0), applied=False)
    env = OracleBraxEnv(sysd, ctx, autoreset=False)
    q = sysd["table"][bs.OFF_INIT_Q:bs.OFF_INIT_Q + sysd["n_q"]][None].astype(np.float32)
    env.init_from_q(q, np.zeros((1, sysd["n_qd"]), np.float32))
    for t in range(100):
        obs, r, d, _ = env.step(np.zeros((1, sysd["n_act"]), np.float32))
    if body == "reacher":
        assert np.isfinite(obs).all() and np.abs(obs[0, 6:8]).max() < 0.5
        return
    if body.startswith("inverted"):  # an unstable equilibrium: the pole may fall, the cart stays on its rail
        assert np.isfinite(obs).all() and abs(obs[0, 0]) < 1.1
        return
    assert np.isfinite(obs).all() and np.abs(obs[0, -sysd["n_qd"]:]).max() < 0.5  # came to rest
    rows = env.state[0, :13 * sysd["n_links"]].reshape(-1, 13)
    if sysd["n_points"] > 0:
        assert rows[:, 2].min() > 0.0  # every link COM above the plane
    if body == "ant":
        assert 0.4 < obs[0, 0] < 0.6  # torso height ~0.55 (the reference notebook shows z = 0.559)


def test_cart_pendulum_small_oscillation_period_is_analytic(hc):
    """Known answer for the restated joint physics: without joint limits the pole of the inverted-pendulum
    body, released 0.05 rad from hanging DOWN on its freely sliding cart, must swing with the textbook period
    of a compound pendulum on a cart, omega^2 = m g d (M + m) / (I_end (M + m) - m^2 d^2), where the inertia
    about the COM is the spring backend's effective one (spring_inertia_scale = 1: unity).
    Checked for the oracle and for the kernel source."""
    sysd = bs.build_system(bs.MODELS["inverted_pendulum"](), {"constraint_limit_stiffness": 0.0})
    assert sysd["tunables"]["spring_inertia_scale"] == 1.0 and sysd["tunables"]["spring_mass_scale"] == 0.0
    ctx = random_ctx(sysd, 1, np.random.default_rng(0), applied=False)
    ctx[:, 3] = 0.0  # no angular damping
    t = sysd["table"]
    o = bs.OFF_LINKS + bs.LINK_STRIDE
    M, m = float(t[bs.OFF_LINKS + bs.L_MASS]), float(t[o + bs.L_MASS])
    d = float(np.linalg.norm(t[o + bs.L_COM:o + bs.L_COM + 3]))
    i_end = 1.0 + m * d * d  # effective COM inertia I^(1 - 1) = 1
    omega = np.sqrt(m * 9.81 * d * (M + m) / (i_end * (M + m) - m * m * d * d))
    q = np.array([[0.0, np.pi - 0.05]], np.float32)
    qd = np.zeros((1, 2), np.float32)
    n_steps, dt = 300, sysd["dt"]

    def fitted_omega(angles):
        a = np.unwrap(np.asarray(angles, np.float64)) - np.pi
        tt = dt * np.arange(1, n_steps + 1)
        grid = omega * np.linspace(0.7, 1.3, 601)
        resid = [np.sum((a + 0.05 * np.cos(w * tt)) ** 2) for w in grid]
        return grid[int(np.argmin(resid))]

    ora = OracleBraxEnv(sysd, ctx, autoreset=False, max_steps=0)
    ora.init_from_q(q, qd)
    st, ob = hc.init(sysd, q, qd)
    el = np.zeros(1, dtype=np.int32)
    ang_o, ang_k = [], []
    zero = np.zeros((1, 1), np.float32)
    for _ in range(n_steps):
        ang_o.append(ora.step(zero)[0][0, 1])
        ang_k.append(hc.step(sysd, st, ctx, zero, el, 0, 0, st.copy(), ob.copy())[0][0, 1])
    assert abs(fitted_omega(ang_o) / omega - 1.0) < 0.02
    assert abs(fitted_omega(ang_k) / omega - 1.0) < 0.02


def _crooked_hopper():
    """A table no shipped body has: rotated link transforms (L_TROT != identity) and joints away from the link
    origin (L_JPOS != 0) -- the general branches of joint_resolve / forward_link that the table-derived
    JointFlags fast path skips for every real model."""
    m = bs.hopper_model()
    m = dict(m, name="crooked_hopper")
    links = [dict(l) for l in m["links"]]
    links[1]["quat"] = bs.quat_axis_angle((0, 1, 0), 0.3)
    links[1]["joint_pos"] = np.array([0.02, 0.0, -0.05])
    links[2]["quat"] = bs.quat_axis_angle((1, 2, 0.5), -0.2)
    links[2]["joint_pos"] = np.array([0.0, 0.01, 0.03])
    links[3]["joint_pos"] = np.array([0.05, 0.0, 0.0])
    m["links"] = links
    return bs.build_system(m)


def test_general_joint_frames_match(hc):
    """hostcheck (kernel source) vs oracle on a body with rotated link transforms and offset joints: pipeline_init,
    one env-step, and the JointFlags really are off for these links (so the general path is what ran)."""
    sysd = _crooked_hopper()
    t = sysd["table"]
    for l in (1, 2):
        o = bs.OFF_LINKS + bs.LINK_STRIDE * l
        assert not (t[o + bs.L_TROT] == 1.0) and np.abs(t[o + bs.L_JPOS:o + bs.L_JPOS + 3]).max() > 0
    n = 128
    rng = np.random.default_rng(5)
    ctx = random_ctx(sysd, n, rng)
    q, qd = random_q(sysd, n, rng, scale=2.0)
    ora = OracleBraxEnv(sysd, ctx, autoreset=False)
    o_ref = ora.init_from_q(q, qd)
    st, o = hc.init(sysd, q, qd)
    np.testing.assert_allclose(st, ora.state, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(o, o_ref, rtol=1e-5, atol=2e-6)
    # forward then inverse kinematics is the identity on (q[1:], qd) for offset joints too
    want = np.concatenate([q[:, 1:], np.clip(qd, -10, 10)], axis=1)
    np.testing.assert_allclose(o_ref, want, rtol=2e-5, atol=2e-5)
    for _ in range(3):
        a = rng.uniform(-1, 1, (n, sysd["n_act"])).astype(np.float32)
        o_ref, r_ref, d_ref, _ = ora.step(a)
        el = np.zeros(n, dtype=np.int32)
        o, r, d = hc.step(sysd, st, ctx, a, el, 1000, 0, st.copy(), o_ref.copy())
        assert_close_scaled(o, o_ref, rel=2e-5)
        assert (d == d_ref).all()
        assert_close_scaled(st, ora.state, rel=2e-5, what="state")
        st[:] = ora.state  # teacher forcing


@pytest.mark.parametrize("body", ["ant", "reacher"])
def test_free_flight_conserves_momentum(hc, body):
    """Size-independent physical invariant of the restated spring pipeline: joint spring / limit / actuator
    wrenches are internal (equal and opposite on child and parent), so for a body out of contact one env-step
    changes the total linear momentum by exactly sum(m) g t -- whatever the actions -- and, with the joint DAMPERS
    off (a damper force acts at two anchors that a stretched spring holds apart, so it alone carries a small net
    torque), leaves the angular momentum about the vertical axis untouched. Checked with the effective masses of
    the backend (Ant / Reacher: unit masses and isotropic unit inertias, so the spin part is simply the angular
    velocity) on the float64 oracle (2e-5, see below), the float32 oracle and the kernel source (float32 round-off).
    A wrong lever arm or a missing reaction term in joint_resolve breaks this at once."""
    base = bs.MODELS[body]()
    if body == "reacher":  # cut the arm loose: a world-anchored hinge exchanges momentum with the world
        links = [dict(l) for l in base["links"]]
        links[0]["type"], links[0]["axis"] = bs.TYPE_FREE, None
        base = dict(base, links=links[:2], init_q=np.array([0, 0, 0, 1, 0, 0, 0, 0.0]), site=None,
                    actuator_links=["body1"], env=bs.ENV_ANT)
        base.pop("obs_dim")
    sysd = bs.build_system(base, {"constraint_vel_damping": 0.0, "constraint_ang_damping": 0.0})
    assert sysd["tunables"]["spring_mass_scale"] == 1.0 and sysd["tunables"]["spring_inertia_scale"] == 1.0
    n, L = 16, sysd["n_links"]
    rng = np.random.default_rng(9)
    ctx = random_ctx(sysd, n, rng)
    ctx[:, 3] = 0.0  # no global angular damping either: it is an external (dissipative) torque
    q, qd = random_q(sysd, n, rng, scale=3.0)
    q[:, 2] += 50.0  # far above the ground for the whole step
    qd += rng.normal(0, 1.0, qd.shape).astype(np.float32)
    a = rng.uniform(-1, 1, (n, sysd["n_act"])).astype(np.float32)
    dt, nf = float(sysd["table"][bs.H_DT]), int(sysd["table"][bs.H_N_FRAMES])
    g = ctx[:, 0].astype(np.float64)
    want_dp = np.stack([0 * g, 0 * g, L * g * np.float32(dt) * nf], axis=1)

    def momenta(state):
        rows = np.asarray(state, np.float64)[:, :13 * L].reshape(n, L, 13)
        pos, vel, ang = rows[..., 0:3], rows[..., 7:10], rows[..., 10:13]
        return vel.sum(1), (np.cross(pos, vel) + ang).sum(1)  # unit effective masses / inertias

    # float64 arithmetic on the float32 TABLE: its frame quaternions are unit only to 1e-7, so a torque pair reaches
    # child and parent scaled by (1 +- 1e-7) -- the residual is ~1e-6 with 150 N m actuators, 1e-8 without actions
    for f64, tol in ((True, 2e-5), (False, 3e-3)):
        ora = OracleBraxEnv(sysd, ctx, autoreset=False, f64=f64)
        ora.init_from_q(q, qd)
        p0, l0 = momenta(ora.state)
        ora.step(a)
        p1, l1 = momenta(ora.state)
        np.testing.assert_allclose(p1 - p0, want_dp, rtol=0, atol=tol)
        np.testing.assert_allclose((l1 - l0)[:, 2], 0.0, atol=tol)
    st, ob = hc.init(sysd, q, qd)
    p0, l0 = momenta(st)
    hc.step(sysd, st, ctx, a, np.zeros(n, np.int32), 0, 0, st.copy(), ob.copy())
    p1, l1 = momenta(st)
    np.testing.assert_allclose(p1 - p0, want_dp, rtol=0, atol=3e-3)
    np.testing.assert_allclose((l1 - l0)[:, 2], 0.0, atol=3e-3)


@pytest.mark.parametrize("body", ["ant", "halfcheetah", "hopper", "walker2d"])
def test_free_fall_follows_the_semi_implicit_euler_closed_form(hc, body):
    """Independent anchor for the integrator, the env's substep schedule and the gravity context: a body released at
    rest in its rest pose, far above the ground, with zero actions, falls rigidly -- after one env-step of n_frames
    substeps of length dt every link has dropped by g dt^2 n (n + 1) / 2 and moves at g dt n (semi-implicit Euler:
    v += g dt; x += v dt), with dt and n_frames the values brax documents for the spring backend (Ant 0.005 x 10,
    Halfcheetah 0.003125 x 16, Hopper / Walker2d 0.002 x 4). Oracle (float64 and float32) and kernel source."""
    sysd = bs.SYSTEMS[body]
    dt_n = {"ant": (0.005, 10), "halfcheetah": (0.003125, 16), "hopper": (0.002, 4), "walker2d": (0.002, 4)}[body]
    n = 4
    rng = np.random.default_rng(3)
    ctx = random_ctx(sysd, n, rng)          # per-env gravity in [-15, -5], random masses
    ctx[:, 3] = 0.0                          # no angular damping
    nq = sysd["n_q"]
    q = np.tile(sysd["table"][bs.OFF_INIT_Q:bs.OFF_INIT_Q + nq].astype(np.float32), (n, 1))
    zi = 2 if body == "ant" else 1           # free root: q = (x, y, z, quat...); planar root: (x, z, pitch)
    q[:, zi] += 40.0
    qd = np.zeros((n, sysd["n_qd"]), np.float32)
    a = np.zeros((n, sysd["n_act"]), np.float32)
    g = ctx[:, 0].astype(np.float64)
    dt, nf = dt_n
    drop = g * dt * dt * nf * (nf + 1) / 2
    speed = g * dt * nf
    L = sysd["n_links"]

    def link_z(state):
        rows = np.array(state, dtype=np.float64)[:, :13 * L].reshape(n, L, 13)  # a copy: the oracle steps in place
        return rows[..., 2], rows[..., 9]   # COM z, COM vz of every link

    for f64, tol in ((True, 1e-9), (False, 2e-5)):
        ora = OracleBraxEnv(sysd, ctx, autoreset=False, f64=f64)
        ora.init_from_q(q, qd)
        z0, _ = link_z(ora.state)
        ora.step(a)
        z1, vz1 = link_z(ora.state)
        np.testing.assert_allclose(z1 - z0, np.repeat(drop[:, None], L, 1), rtol=0, atol=max(tol, 1e-7 * 40))
        np.testing.assert_allclose(vz1, np.repeat(speed[:, None], L, 1), rtol=0, atol=max(tol, 1e-7))
    st, ob = hc.init(sysd, q, qd)
    z0, _ = link_z(st)
    hc.step(sysd, st, ctx, a, np.zeros(n, np.int32), 0, 0, st.copy(), ob.copy())
    z1, vz1 = link_z(st)
    np.testing.assert_allclose(z1 - z0, np.repeat(drop[:, None], L, 1), rtol=0, atol=2e-5)
    np.testing.assert_allclose(vz1, np.repeat(speed[:, None], L, 1), rtol=0, atol=2e-5)
