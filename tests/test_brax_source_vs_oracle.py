"""No-GPU check of the Brax kernel source logic (physics_brax.h through tests/hostcheck) against
the CPU oracle (oracle/brax_oracle.c): forward kinematics, n_frames spring substeps, env layer,
EpisodeWrapper truncation and AutoReset. float32 on both sides; 1e-5 relative per step (north
star), done masks identical."""
import numpy as np
import pytest

from carl_b200.envs import brax_system as bs
from oracle.brax import OracleBraxEnv
from tests.brax_util import BraxHostCheck, assert_close_scaled, pusher_contact_states, random_ctx, random_q

BODIES = ["ant", "halfcheetah", "hopper", "walker2d", "inverted_pendulum", "inverted_double_pendulum", "reacher",
          "humanoid", "humanoidstandup", "pusher"]
HUMANOIDS = ("humanoid", "humanoidstandup")


def _stacked_rate_projection(sysd, q, qd):
    """What kinematics.inverse reports for the rates of a 3-dof stacked hinge: the projections of the relative
    angular velocity on the Euler axes e_x, Rx e_y, R e_z -- the first and the third are not orthogonal
    (their dot product is sin(theta)), so the projections mix qd0 and qd2; dof 1 and 2-dof joints are exact."""
    t = sysd["table"]
    out = qd.astype(np.float64).copy()
    for l in range(sysd["n_links"]):
        o = bs.OFF_LINKS + bs.LINK_STRIDE * l
        if int(t[o + bs.L_TYPE]) != bs.TYPE_HINGE3:
            continue
        qi, di = int(t[o + bs.L_QIDX]), int(t[o + bs.L_QDIDX])
        sg = t[bs.OFF_DOF + bs.DOF_STRIDE * l + bs.D_SIGN0:bs.OFF_DOF + bs.DOF_STRIDE * l + bs.D_SIGN0 + 3].astype(np.float64)
        sth = np.sin(sg[1] * q[:, qi + 1].astype(np.float64))
        w0, w2 = sg[0] * qd[:, di], sg[2] * qd[:, di + 2]
        out[:, di], out[:, di + 2] = sg[0] * (w0 + sth * w2), sg[2] * (w2 + sth * w0)
    return out


@pytest.fixture(scope="module")
def hc():
    return BraxHostCheck()


@pytest.mark.parametrize("body", BODIES)
def test_pipeline_init_matches(hc, body):
    sysd = bs.SYSTEMS[body]
    rng = np.random.default_rng(0)
    q, qd = random_q(sysd, 64, rng, scale=3.0)
    ctx = random_ctx(sysd, 64, rng)
    ora = OracleBraxEnv(sysd, ctx)
    o_ref = ora.init_from_q(q, qd)
    st, o = hc.init(sysd, q, qd, ctx)
    np.testing.assert_allclose(st, ora.state, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(o, o_ref, rtol=1e-5, atol=2e-6)
    if body in HUMANOIDS:
        # forward then inverse kinematics: identity on q[2:] (Euler angles of the stacked hinges included) and on qd up
        # to the documented projection of the 3-dof rates; then the extra blocks of brax.envs.humanoid._get_obs
        nq, nqd, L = sysd["n_q"], sysd["n_qd"], sysd["n_links"]
        want_q = q[:, 2:].astype(np.float64)
        want_q[:, 1:5] /= np.linalg.norm(want_q[:, 1:5], axis=1, keepdims=True)
        np.testing.assert_allclose(o_ref[:, :nq - 2], want_q, rtol=2e-5, atol=2e-5)
        want_qd = _stacked_rate_projection(sysd, q, qd)
        want_qd[:, 3:6] = o_ref[:, nq - 2 + 3:nq - 2 + 6]  # root angular velocity is reported in the local frame
        np.testing.assert_allclose(o_ref[:, nq - 2:nq - 2 + nqd], want_qd, rtol=2e-5, atol=2e-5)
        k = nq - 2 + nqd
        cin = o_ref[:, k:k + 10 * L].reshape(-1, L, 10).astype(np.float64)
        inertia = cin[..., :9].reshape(-1, L, 3, 3)
        np.testing.assert_allclose(inertia, inertia.transpose(0, 1, 3, 2), atol=1e-5)          # symmetric
        assert (np.linalg.eigvalsh(inertia) > 0).all()                                         # positive definite
        np.testing.assert_allclose(cin[..., 9], ctx[:, 5:], rtol=1e-6)                         # mass_scale 0: the context masses
        rows = ora.state[:, :13 * L].reshape(-1, L, 13).astype(np.float64)
        m = ctx[:, 5:].astype(np.float64)
        com = (m[..., None] * rows[..., :3]).sum(1) / m.sum(1)[:, None]
        p = rows[..., :3] - com[:, None]
        # trace of R I R^T + m (|p|^2 E - p p^T) with unit effective inertia: 3 + 2 m |p|^2
        np.testing.assert_allclose(np.trace(inertia, axis1=2, axis2=3), 3.0 + 2.0 * m * (p * p).sum(-1), rtol=2e-5)
        cvel = o_ref[:, k + 10 * L:k + 16 * L].reshape(-1, L, 6)
        np.testing.assert_allclose(cvel[..., :3].sum(1), (m[..., None] * rows[..., 7:10]).sum(1) / m.sum(1)[:, None],
                                   rtol=1e-4, atol=1e-5)                                       # sums to the COM velocity
        np.testing.assert_allclose(o_ref[:, k + 16 * L:], 0.0)                                 # no action at reset
        return
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
    if body == "pusher":  # q[:7], qd[:7], centres of mass of the wrist-flex link, the object and the goal
        L = sysd["n_links"]
        rows = ora.state[:, :13 * L].reshape(-1, L, 13)
        com = rows[:, [5, 7, 8], :3].copy()
        com[..., 2] -= 0.325  # heights are reported in the MJCF's world (table at z = -0.325)
        want = np.concatenate([q[:, :7], qd[:, :7], com.reshape(-1, 9)], axis=1)
        # the object and the goal ride on their slides: x = 0.45 + second coordinate, y = -0.05 + first coordinate
        np.testing.assert_allclose(com[:, 1, :2], np.stack([0.45 + q[:, 8], -0.05 + q[:, 7]], axis=1), atol=2e-6)
        np.testing.assert_allclose(com[:, 2, :2], np.stack([0.45 + q[:, 10], -0.05 + q[:, 9]], axis=1), atol=2e-6)
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
    st, _ = hc.init(sysd, q, qd, ctx)
    a = rng.uniform(-1.2, 1.2, (n, sysd["n_act"])).astype(np.float32) * sysd["act_scale"]
    o_ref, r_ref, d_ref, _ = ora.step(a)
    el = np.zeros(n, dtype=np.int32)
    o, r, d = hc.step(sysd, st, ctx, a, el, 1000, 0, st.copy(), o_ref.copy(), stock_contact=0 if applied else 1)
    assert_close_scaled(o, o_ref, rel=1e-5)
    np.testing.assert_allclose(r, r_ref, rtol=1e-4, atol=1e-4)
    assert (d == d_ref).all()
    assert_close_scaled(st, ora.state, rel=1e-5, what="state")


@pytest.mark.parametrize("body", BODIES)
def test_fast_arithmetic_is_the_same_mathematics(hc, body):
    """The FAST reformulations of the FMA build (`arithmetic="fma"`: hinge wrench evaluated in the world frame,
    unit-inertia shortcut, contact centres taken from the centre of mass -- physics_brax.h) are the same mathematics
    rounded differently: against the float64 yardstick one env-step stays where the reference-order float32 arithmetic
    itself sits (observations 2e-5, link state 6e-5 of the env's vector magnitude; the float32 restatement is 0.3-1.7e-5 /
    0.4-3.1e-5 away from float64 on these samples), done masks identical."""
    sysd = bs.SYSTEMS[body]
    n = 512
    rng = np.random.default_rng(1)
    ctx = random_ctx(sysd, n, rng)
    if body == "pusher":
        # the MJCF's zero gravity: under gravity the ball, which rests exactly tangent to the table, chatters between
        # contact and no contact -- in float32 and float64 differently, whatever the implementation
        ctx[:, 0] = 0.0
    q, qd = random_q(sysd, n, rng, scale=2.0)
    ora64 = OracleBraxEnv(sysd, ctx, autoreset=False, f64=True)
    ora64.init_from_q(q, qd)
    a = (rng.uniform(-1.2, 1.2, (n, sysd["n_act"])) * sysd["act_scale"]).astype(np.float32)
    o64, r64, d64, _ = ora64.step(a)
    st, o0 = hc.init(sysd, q, qd, ctx)
    el = np.zeros(n, dtype=np.int32)
    o, r, d = hc.step(sysd, st, ctx, a, el, 1000, 0, st.copy(), o0.copy(), fast=True)
    # (the pusher integrates 50 substeps per env-step, five times the Ant's: its round-off floor is accordingly higher)
    assert_close_scaled(o, o64, rel=1e-4 if body == "pusher" else 2e-5)
    assert_close_scaled(st, ora64.state, rel=1e-4 if body == "pusher" else 6e-5, what="state")
    assert (d == d64).all()


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
    st, o = hc.init(sysd, q, qd, ctx)
    first_state, first_obs = st.copy(), o.copy()
    el = np.zeros(n, dtype=np.int32)
    n_done = 0
    for t in range(T):
        a = rng.uniform(-1, 1, (n, sysd["n_act"])).astype(np.float32) * sysd["act_scale"]
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
    if body in HUMANOIDS:
        # 1.5 s without control: the humanoid sags / topples, the lying one settles on the ground -- no blow-up,
        # bounded rates, every link COM stays above the plane, joints stay inside their (soft) limits + slack
        rows = env.state[0, :13 * sysd["n_links"]].reshape(-1, 13)
        assert np.isfinite(obs).all() and np.abs(obs[0, 22:45]).max() < 30.0
        assert rows[:, 2].min() > 0.0 and rows[:, 2].max() < 1.8
        assert np.abs(obs[0, 5:22]).max() < np.pi
        if body == "humanoidstandup":
            assert obs[0, 0] < 0.3  # still lying
        return
    if body == "reacher":
        assert np.isfinite(obs).all() and np.abs(obs[0, 6:8]).max() < 0.5
        return
    if body == "pusher":  # zero gravity, no control: the arm stays where it is, ball and goal do not move
        assert np.isfinite(obs).all() and np.abs(obs[0, :14]).max() < 1e-3
        np.testing.assert_allclose(obs[0, 14:], [0.821, -0.6, 0.0, 0.45, -0.05, -0.275, 0.45, -0.05, -0.323], atol=1e-4)
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


def _hanging_rod(stacked_type):
    """A synthetic body: one capsule hanging from the world on a stacked hinge (2-dof: axes x, y; 3-dof: axes
    x, z, y -- the humanoid hip's left-handed MJCF order, whose third coordinate runs against the joint frame's z)."""
    axes = [(1, 0, 0), (0, 1, 0)] if stacked_type == bs.TYPE_HINGE2 else [(1, 0, 0), (0, 0, 1), (0, 1, 0)]
    nd = len(axes)
    link = bs._link("rod", -1, stacked_type, (0, 0, 2.0), [bs.capsule((0, 0, 0), (0, 0, -0.6), 0.05)], axis=axes,
                    limit=[(-bs.UNLIMITED, bs.UNLIMITED)] * nd, gear=[10.0] * nd)
    return dict(
        name="hanging_rod", env=bs.ENV_HALFCHEETAH, links=[link], density=1000.0, total_mass=None, friction=1.0,
        init_q=np.zeros(nd), dt=0.002, n_frames=4, contacts=False,
        tunables=dict(constraint_stiffness=20000.0, constraint_vel_damping=50.0, constraint_limit_stiffness=0.0,
                      constraint_ang_damping=0.0, baumgarte_erp=0.1, vel_damping=0.0, spring_mass_scale=0.0,
                      spring_inertia_scale=1.0),
        env_params=dict(reset_noise=0.0, ctrl_cost=0.0, healthy_reward=0.0, z_min=-1e9, z_max=1e9, forward_weight=0.0,
                        angle_min=0.0, angle_max=0.0, exclude_pos=0, qd_clip=0.0, terminate=0.0),
        stock_gravity=-9.81, stock_ang_damping=0.0, stock_elasticity=0.0,
        actuator_links=[f"rod:{k}" for k in range(nd)],
    )


@pytest.mark.parametrize("stacked_type", [bs.TYPE_HINGE2, bs.TYPE_HINGE3])
def test_stacked_hinge_pendulum_period_and_torque_response_are_analytic(hc, stacked_type):
    """Known answers for the stacked (universal / spherical) joints that nothing in the code encodes: a rod hanging
    from the world on such a joint is a compound pendulum about BOTH horizontal axes, omega^2 = m g d / (I + m d^2)
    with the backend's unit effective inertia; released with small angles about x and y it must swing in both
    coordinates at that frequency (the y coordinate is dof 1 of the 2-dof joint, and dof 2 -- with the reversed
    sign of the left-handed x, z, y stack -- of the 3-dof joint). And a constant actuator torque tau on one dof
    deflects that coordinate statically by asin(tau / (m g d)), in the positive direction of its MJCF axis."""
    sysd = bs.build_system(_hanging_rod(stacked_type))
    nd = sysd["n_q"]
    ydof = 1 if stacked_type == bs.TYPE_HINGE2 else 2
    t = sysd["table"]
    m = float(t[bs.OFF_LINKS + bs.L_MASS])
    d = float(np.linalg.norm(t[bs.OFF_LINKS + bs.L_COM:bs.OFF_LINKS + bs.L_COM + 3]))
    omega = np.sqrt(m * 9.81 * d / (1.0 + m * d * d))
    ctx = random_ctx(sysd, 1, np.random.default_rng(0), applied=False)
    ctx[:, 3] = 0.0
    q = np.zeros((1, nd), np.float32)
    q[0, 0], q[0, ydof] = 0.05, 0.03
    qd = np.zeros((1, nd), np.float32)
    n_steps, dt = 400, sysd["dt"]
    tt = dt * np.arange(1, n_steps + 1)

    def fitted(series, amp):
        grid = omega * np.linspace(0.7, 1.3, 601)
        resid = [np.sum((np.asarray(series, np.float64) - amp * np.cos(w * tt)) ** 2) for w in grid]
        return grid[int(np.argmin(resid))]

    zero = np.zeros((1, nd), np.float32)
    ora = OracleBraxEnv(sysd, ctx, autoreset=False, max_steps=0)
    ora.init_from_q(q, qd)
    st, ob = hc.init(sysd, q, qd, ctx)
    el = np.zeros(1, dtype=np.int32)
    tr_o, tr_k = [], []
    for _ in range(n_steps):
        tr_o.append(ora.step(zero)[0][0, :nd].copy())
        tr_k.append(hc.step(sysd, st, ctx, zero, el, 0, 0, st.copy(), ob.copy())[0][0, :nd].copy())
    for tr in (np.array(tr_o), np.array(tr_k)):
        assert abs(fitted(tr[:, 0], 0.05) / omega - 1.0) < 0.02
        assert abs(fitted(tr[:, ydof], 0.03) / omega - 1.0) < 0.02
        if nd == 3:
            assert np.abs(tr[:, 1]).max() < 5e-3  # no spin about the rod's own axis appears
    # static deflection under a constant torque about the y-like dof (critically damped by the global ang_damping)
    ctx2 = ctx.copy()
    ctx2[:, 3] = -8.0
    tau = 0.5 * m * 9.81 * d * np.sin(0.3)
    act = np.zeros((1, nd), np.float32)
    act[0, ydof] = tau / 10.0  # gear 10, inside the ctrl range +-1
    assert tau / 10.0 < 1.0
    ora = OracleBraxEnv(sysd, ctx2, autoreset=False, max_steps=0)
    ora.init_from_q(np.zeros((1, nd), np.float32), qd)
    st, ob = hc.init(sysd, np.zeros((1, nd), np.float32), qd, ctx2)
    for _ in range(1500):
        o_o = ora.step(act)[0]
        o_k = hc.step(sysd, st, ctx2, act, el, 0, 0, st.copy(), ob.copy())[0]
    want = np.arcsin(tau / (m * 9.81 * d))
    for o in (o_o, o_k):
        assert o[0, ydof] == pytest.approx(want, rel=0.01)
        assert abs(o[0, 0]) < 1e-3


def test_pusher_gripper_ball_contacts_match(hc):
    """The body-vs-body pairs of the pusher (gripper capsules against the pushed ball) on states built to have them:
    kernel source vs oracle, teacher-forced over several env-steps; the ball must really be struck in a good share of
    the envs (it has no actuator and no gravity: only a contact can set it in motion)."""
    sysd = bs.SYSTEMS["pusher"]
    n = 512
    rng = np.random.default_rng(4)
    ctx = random_ctx(sysd, n, rng)
    ctx[:, 0] = 0.0  # the MJCF's zero gravity
    q, qd = pusher_contact_states(sysd, n, rng)
    ora = OracleBraxEnv(sysd, ctx, autoreset=False, max_steps=0)
    o0 = ora.init_from_q(q, qd)
    st, o = hc.init(sysd, q, qd, ctx)
    np.testing.assert_allclose(st, ora.state, rtol=1e-6, atol=1e-6)
    el = np.zeros(n, dtype=np.int32)
    struck = np.zeros(n, bool)
    for _ in range(4):
        a = rng.uniform(-2, 2, (n, 7)).astype(np.float32)
        o_ref, r_ref, d_ref, _ = ora.step(a)
        o, r, d = hc.step(sysd, st, ctx, a, el, 0, 0, st.copy(), o0.copy())
        assert_close_scaled(o, o_ref, rel=1e-5)
        assert_close_scaled(st, ora.state, rel=1e-5, what="state")
        np.testing.assert_allclose(r, r_ref, rtol=1e-5, atol=1e-5)
        assert not d.any() and not d_ref.any()
        struck |= np.abs(ora.state[:, 13 * 7 + 7:13 * 7 + 9]).max(axis=1) > 1e-3
        st[:] = ora.state
    assert struck.mean() > 0.15
    # reward of brax.envs.pusher.step: positions BEFORE the step: -|object - goal| - 0.1 |a|^2 - 0.5 |object - wrist-flex link|
    rows = ora.state[:, :13 * 9].reshape(n, 9, 13).copy()
    a = rng.uniform(-2, 2, (n, 7)).astype(np.float32)
    _, r_ref, _, _ = ora.step(a)
    want = (-np.linalg.norm(rows[:, 7, :3] - rows[:, 8, :3], axis=1) - 0.1 * (a.astype(np.float64) ** 2).sum(1)
            - 0.5 * np.linalg.norm(rows[:, 7, :3] - rows[:, 5, :3], axis=1))
    np.testing.assert_allclose(r_ref, want, rtol=1e-5, atol=1e-5)


def test_pusher_matches_the_reference_notebooks_printed_step(hc):
    """The one Brax step whose output the reference tree holds (examples/brax_with_goals.ipynb, cell 5: CARLBraxPusher,
    spring backend, default context, `reset()` then ONE `step` with an unrecorded random action; extracted into
    tests/golden/notebook_goldens.json by tools/extract_notebook_goldens.py). The action is unknown, so the joint part
    cannot be replayed, but everything else the printout fixes must hold for the restatement:
    * the layout f32[23] = q[:7], qd[:7], three positions; the goal at (0.45, -0.05, -0.323) and the object's height
      -0.275 in the MJCF's world (the engine's shifted plane is subtracted again), exactly;
    * the observed "tip" is the centre of mass of the wrist-flex link of the seven-link chain: (0.821, -0.6, 0) at the
      initial pose -- the printout, one 0.05 s step later, is within 3e-4 of it;
    * the object sits where brax.envs.pusher.reset can put it with OUR reading of the two slide coordinates (first
      coordinate along y from U(-0.3, 0), second along x from U(-0.2, 0.2), at least 0.17 from the goal);
    * the printed reward minus the two distance terms (taken before the step) leaves -0.1 |a|^2 with |a|^2 = 2.28,
      inside what seven actions of the env's action space can give."""
    import json
    import os

    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "notebook_goldens.json")))
    ref = np.array(gold["pusher_obs_after_one_random_step"], np.float64)
    reward = gold["pusher_reward_after_one_random_step"]
    sysd = bs.SYSTEMS["pusher"]
    assert ref.shape == (sysd["obs_dim"],)
    dy, dx = ref[18] + 0.05, ref[17] - 0.45
    assert -0.3 <= dy <= 0.0 and -0.2 <= dx <= 0.2 and np.hypot(dx, dy) >= 0.17
    q = np.zeros((1, 11), np.float32)
    q[0, 7], q[0, 8] = dy, dx
    qd = np.zeros((1, 11), np.float32)
    ctx = random_ctx(sysd, 1, np.random.default_rng(0), applied=False)
    ora = OracleBraxEnv(sysd, ctx, autoreset=False, max_steps=0)
    o_ora = ora.init_from_q(q, qd)[0]
    _, o_hc = hc.init(sysd, q, qd, ctx)
    for o in (o_ora, o_hc[0]):
        np.testing.assert_allclose(o[17:23], ref[17:23], atol=2e-7)          # object and goal: exact
        np.testing.assert_allclose(o[14:17], ref[14:17], atol=3e-4)          # tip: the printout is one step later
        np.testing.assert_allclose(o[14:17], [0.821, -0.6, 0.0], atol=2e-6)
    # the reward of brax.envs.pusher.step reads the positions before the step
    dist = np.linalg.norm(ref[17:20] - ref[20:23])
    near = np.linalg.norm(ref[17:20] - np.array([0.821, -0.6, 0.0]))
    a_sq = (-reward - dist - 0.5 * near) / 0.1
    assert a_sq == pytest.approx(2.278, abs=0.01) and 0.0 < a_sq < 7 * sysd["act_scale"] ** 2
    # the same through our own step: zero action from this state gives exactly the two distance terms
    _, r0, d0, _ = ora.step(np.zeros((1, 7), np.float32))
    assert r0[0] == pytest.approx(-(dist + 0.5 * near), abs=2e-6) and not d0[0]


def test_ant_initial_pose_matches_the_reference_notebooks_printed_observation():
    """examples/brax_with_goals.ipynb, cell 3: CARLBraxAnt's observation f32[27] after reset() and one random step.
    The reset noise is +-0.1 on q, so the printout pins the initial pose's PATTERN: torso height about 0.55, a
    near-identity root quaternion first, then the eight joint angles around (0, 1, 0, -1, 0, -1, 0, 1) in this order."""
    import json
    import os

    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "notebook_goldens.json")))
    ref = np.array(gold["ant_obs_after_one_random_step"], np.float64)
    sysd = bs.SYSTEMS["ant"]
    assert ref.shape == (sysd["obs_dim"],)
    init_q = sysd["table"][bs.OFF_INIT_Q:bs.OFF_INIT_Q + 15].astype(np.float64)
    assert abs(ref[0] - init_q[2]) < 0.12                                    # z (x, y are excluded from the observation)
    assert abs(np.linalg.norm(ref[1:5]) - 1.0) < 1e-5 and ref[1] > 0.98      # unit quaternion, w first
    np.testing.assert_allclose(ref[5:13], init_q[7:15], atol=0.2)            # +-0.1 reset noise + one step of motion


def _bar_and_ball():
    """A synthetic two-body system no shipped env has: a free-floating capsule and a free-floating ball (real masses,
    unit effective inertias), paired for body-vs-body contact, far above the ground, no gravity."""
    bar = bs._link("bar", -1, bs.TYPE_FREE, (0, 0, 5.0), [bs.capsule((-0.3, 0, 0), (0.3, 0, 0), 0.05)])
    ball = bs._link("ball", -1, bs.TYPE_FREE, (0, 0.5, 5.0), [bs.sphere((0, 0, 0), 0.08)])
    init_q = np.array([0, 0, 5, 1, 0, 0, 0, 0, 0.5, 5, 1, 0, 0, 0], dtype=np.float64)
    return dict(
        name="bar_and_ball", env=bs.ENV_HALFCHEETAH, links=[bar, ball], density=1000.0, total_mass=None, friction=0.0,
        init_q=init_q, dt=0.001, n_frames=1, pairs=[("bar", 0, "ball", 0)],
        tunables=dict(constraint_stiffness=1000.0, constraint_vel_damping=10.0, constraint_limit_stiffness=0.0,
                      constraint_ang_damping=0.0, baumgarte_erp=0.1, vel_damping=0.0, spring_mass_scale=0.0,
                      spring_inertia_scale=1.0),
        env_params=dict(reset_noise=0.0, ctrl_cost=0.0, healthy_reward=0.0, z_min=-1e9, z_max=1e9, forward_weight=0.0,
                        angle_min=0.0, angle_max=0.0, exclude_pos=0, qd_clip=0.0, terminate=0.0),
        stock_gravity=0.0, stock_ang_damping=0.0, stock_elasticity=0.0, actuator_links=[],
    )


def test_two_body_collision_conserves_momentum_and_restitutes(hc):
    """Independent anchor of the body-vs-body impulse (nothing in the code encodes it): a ball thrown at a free bar.
    The impulse is equal and opposite and acts at ONE point, so the total linear momentum and the total angular
    momentum about the world origin (orbital + spin, unit effective inertias) are conserved through the collision;
    and along the contact normal the separation speed after one substep is e times the approach speed plus the
    Baumgarte term erp * penetration / dt. Oracle (float64, float32) and kernel source."""
    model = _bar_and_ball()
    model["env_params"] = dict(model["env_params"])
    sysd = bs.build_system(model)
    assert sysd["n_act"] == 0 and sysd["n_q"] == 14
    n = 64
    rng = np.random.default_rng(11)
    ctx = random_ctx(sysd, n, rng)
    ctx[:, 0] = 0.0       # no gravity
    ctx[:, 1] = 0.0       # frictionless: the impulse is purely normal
    ctx[:, 3] = 0.0       # no angular damping
    el_ = ctx[:, 2].astype(np.float64)
    m = ctx[:, 5:7].astype(np.float64)
    q = np.tile(np.asarray(model["init_q"], np.float32), (n, 1))
    # ball centre 0.12 from the bar's axis (radii 0.05 + 0.08: penetration 0.01), somewhere along the bar
    q[:, 7] = rng.uniform(-0.25, 0.25, n)
    q[:, 8] = 0.12
    qd = np.zeros((n, 12), np.float32)
    qd[:, 7] = -rng.uniform(0.5, 2.0, n)      # ball flies towards the bar (-y)
    qd[:, 6] = rng.uniform(-1, 1, n)          # and sideways
    qd[:, 0:3] = rng.uniform(-0.2, 0.2, (n, 3))
    qd[:, 5] = rng.uniform(-1, 1, n)          # the bar spins about z
    a = np.zeros((n, 0), np.float32)

    def momenta(state):
        rows = np.array(state, dtype=np.float64)[:, :26].reshape(n, 2, 13)  # a copy: the oracle steps in place
        pos, vel, ang = rows[..., 0:3], rows[..., 7:10], rows[..., 10:13]
        mv = m[..., None] * vel
        return mv.sum(1), (np.cross(pos, mv) + ang).sum(1), rows

    def normal_speed(rows):  # contact normal = +y here (ball above the bar's axis in y), approach speed of the contact points
        contact = np.stack([rows[:, 1, 0], 0.5 * (0.05 + (0.12 - 0.08)) + 0 * rows[:, 1, 0], rows[:, 1, 2]], axis=1)
        va = rows[:, 0, 7:10] + np.cross(rows[:, 0, 10:13], contact - rows[:, 0, 0:3])
        vb = rows[:, 1, 7:10] + np.cross(rows[:, 1, 10:13], contact - rows[:, 1, 0:3])
        return (vb - va)[:, 1]

    for f64, tol in ((True, 1e-6), (False, 2e-5)):
        ora = OracleBraxEnv(sysd, ctx, autoreset=False, f64=f64, max_steps=0)
        ora.init_from_q(q, qd)
        # the spinning, translating bar: make the geometry exact at the instant of contact (bar along x at y = 0)
        p0, l0, r0 = momenta(ora.state)
        v0 = normal_speed(r0)
        assert (v0 < 0).all()
        ora.step(a)
        p1, l1, r1 = momenta(ora.state)
        np.testing.assert_allclose(p1, p0, atol=tol)
        np.testing.assert_allclose(l1, l0, atol=tol * 10)
        # the velocities changed (a collision happened) ...
        assert (np.abs(r1[:, 1, 8] - r0[:, 1, 8]) > 1e-3).all()
        # ... and the normal separation speed is -e v0 + erp * penetration / dt
        rows_after = r1.copy()
        rows_after[..., 0:7] = r0[..., 0:7]  # evaluate at the contact geometry of the impulse
        v1 = normal_speed(rows_after)
        np.testing.assert_allclose(v1, -el_ * v0 + 0.1 * 0.01 / 0.001, rtol=2e-4 if f64 else 2e-3, atol=1e-4)
    st, ob = hc.init(sysd, q, qd, ctx)
    p0, l0, r0 = momenta(st)
    hc.step(sysd, st, ctx, a, np.zeros(n, np.int32), 0, 0, st.copy(), ob.copy())
    p1, l1, r1 = momenta(st)
    np.testing.assert_allclose(p1, p0, atol=2e-5)
    np.testing.assert_allclose(l1, l0, atol=2e-4)
    np.testing.assert_allclose(st, ora.state, rtol=1e-5, atol=1e-5)


def _lone_ball(dt, n_frames):
    """A synthetic one-link body: a free ball of radius 0.1 above the ground plane."""
    ball = bs._link("ball", -1, bs.TYPE_FREE, (0, 0, 0.5), [bs.sphere((0, 0, 0), 0.1)])
    return dict(
        name="lone_ball", env=bs.ENV_HALFCHEETAH, links=[ball], density=1000.0, total_mass=None, friction=1.0,
        init_q=np.array([0, 0, 0.5, 1, 0, 0, 0], dtype=np.float64), dt=dt, n_frames=n_frames,
        tunables=dict(constraint_stiffness=1000.0, constraint_vel_damping=10.0, constraint_limit_stiffness=0.0,
                      constraint_ang_damping=0.0, baumgarte_erp=0.1, vel_damping=0.0, spring_mass_scale=0.0,
                      spring_inertia_scale=1.0),
        env_params=dict(reset_noise=0.0, ctrl_cost=0.0, healthy_reward=0.0, z_min=-1e9, z_max=1e9, forward_weight=0.0,
                        angle_min=0.0, angle_max=0.0, exclude_pos=0, qd_clip=0.0, terminate=0.0),
        stock_gravity=-9.81, stock_ang_damping=0.0, stock_elasticity=0.0, actuator_links=[],
    )


@pytest.mark.parametrize("dt", [0.005, 0.002])
def test_ball_in_the_ground_is_pushed_out_at_the_baumgarte_rate(hc, dt):
    """Independent anchor of the ground-contact impulse (brax.spring.collisions with Baumgarte stabilisation): a ball
    at rest 5 mm inside the plane, zero restitution. In the first substep the semi-implicit update makes the normal
    velocity -g dt, the contact impulse replaces it by erp * penetration / dt WHATEVER the mass and the gravity, and the
    pose moves by that times dt: penetration (1 - erp) * 5 mm, velocity erp * 5 mm / dt. Afterwards the ball only
    feels the contact while it approaches the plane: replaying that scalar rule (free flight under gravity while moving
    out, the Baumgarte velocity when moving in) reproduces height and velocity over 40 substeps. Oracle (float64,
    float32) and kernel source, two step sizes, random per-env gravity and mass."""
    sysd = bs.build_system(_lone_ball(dt, 1))
    n = 16
    rng = np.random.default_rng(5)
    ctx = random_ctx(sysd, n, rng)
    ctx[:, 1], ctx[:, 2], ctx[:, 3] = 1.0, 0.0, 0.0   # friction 1, no restitution, no angular damping
    g = -ctx[:, 0].astype(np.float64)
    q = np.tile(np.asarray(sysd["table"][bs.OFF_INIT_Q:bs.OFF_INIT_Q + 7], np.float32), (n, 1))
    q[:, 2] = 0.095
    qd = np.zeros((n, 6), np.float32)
    a = np.zeros((n, 0), np.float32)
    pen0 = 0.1 - np.float64(np.float32(0.095))
    erp, dtf = 0.1, np.float64(np.float32(dt))

    def replay(steps):  # the scalar rule of the docstring
        pen, v, out = np.full(n, pen0), np.zeros(n), []
        for _ in range(steps):
            v = v - g * dtf
            hit = (pen > 0) & (v < 0)
            v = np.where(hit, erp * pen / dtf, v)
            pen = pen - v * dtf
            out.append((pen.copy(), v.copy()))
        return out

    ref = replay(40)
    for f64, tol in ((True, 1e-7), (False, 2e-6)):
        ora = OracleBraxEnv(sysd, ctx, autoreset=False, f64=f64, max_steps=0)
        ora.init_from_q(q, qd)
        for k in range(40):
            ora.step(a)
            row = np.array(ora.state, dtype=np.float64)
            if k == 0:  # the closed form of the first substep
                np.testing.assert_allclose(0.1 - row[:, 2], (1 - erp) * pen0, rtol=1e-5)
                np.testing.assert_allclose(row[:, 9], erp * pen0 / dtf, rtol=1e-5)
            np.testing.assert_allclose(0.1 - row[:, 2], ref[k][0], atol=tol * (k + 1))
            np.testing.assert_allclose(row[:, 9], ref[k][1], atol=tol * (k + 1) / dt)
        assert (np.abs(0.1 - row[:, 2]) < pen0).all()   # closer to the surface than it started
    st, ob = hc.init(sysd, q, qd, ctx)
    el = np.zeros(n, np.int32)
    for k in range(40):
        hc.step(sysd, st, ctx, a, el, 0, 0, st.copy(), ob.copy())
        np.testing.assert_allclose(0.1 - st[:, 2].astype(np.float64), ref[k][0], atol=2e-6 * (k + 1))


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


@pytest.mark.parametrize("body", ["ant", "reacher", "humanoid"])
def test_free_flight_conserves_momentum(hc, body):
    """Size-independent physical invariant of the restated spring pipeline: joint spring / limit / actuator
    wrenches are internal (equal and opposite on child and parent), so for a body out of contact one env-step
    changes the total linear momentum by exactly sum(m) g t -- whatever the actions -- and, with the joint DAMPERS
    off (a damper force acts at two anchors that a stretched spring holds apart, so it alone carries a small net
    torque), leaves the angular momentum about the vertical axis untouched. Checked with the effective masses of
    the backend (Ant / Reacher: unit masses and isotropic unit inertias, so the spin part is simply the angular
    velocity) on the float64 oracle (2e-5, see below), the float32 oracle and the kernel source (float32 round-off).
    The humanoid adds the stacked (2- / 3-dof) hinges with offset anchors and real link masses (spring_mass_scale 0):
    its momenta are mass weighted and its tolerances scale with its 350 N m actuators and ~40 kg.
    A wrong lever arm or a missing reaction term in joint_resolve breaks this at once."""
    base = bs.MODELS[body]()
    if body == "reacher":  # cut the arm loose: a world-anchored hinge exchanges momentum with the world
        links = [dict(l) for l in base["links"]]
        links[0]["type"], links[0]["axis"] = bs.TYPE_FREE, None
        base = dict(base, links=links[:2], init_q=np.array([0, 0, 0, 1, 0, 0, 0, 0.0]), site=None,
                    actuator_links=["body1"], env=bs.ENV_ANT)
        base.pop("obs_dim")
    sysd = bs.build_system(base, {"constraint_vel_damping": 0.0, "constraint_ang_damping": 0.0})
    assert sysd["tunables"]["spring_inertia_scale"] == 1.0
    assert sysd["tunables"]["spring_mass_scale"] == (0.0 if body == "humanoid" else 1.0)
    n, L = 16, sysd["n_links"]
    rng = np.random.default_rng(9)
    ctx = random_ctx(sysd, n, rng)
    ctx[:, 3] = 0.0  # no global angular damping either: it is an external (dissipative) torque
    q, qd = random_q(sysd, n, rng, scale=3.0)
    q[:, 2] += 50.0  # far above the ground for the whole step
    qd += rng.normal(0, 1.0, qd.shape).astype(np.float32)
    a = rng.uniform(-1, 1, (n, sysd["n_act"])).astype(np.float32) * sysd["act_scale"]
    dt, nf = float(sysd["table"][bs.H_DT]), int(sysd["table"][bs.H_N_FRAMES])
    g = ctx[:, 0].astype(np.float64)
    # effective masses of the backend: mass^(1 - spring_mass_scale)
    meff = ctx[:, 5:].astype(np.float64) if body == "humanoid" else np.ones((n, L))
    want_dp = np.stack([0 * g, 0 * g, meff.sum(1) * g * np.float32(dt) * nf], axis=1)
    scale = 40.0 if body == "humanoid" else 1.0

    def momenta(state):
        rows = np.asarray(state, np.float64)[:, :13 * L].reshape(n, L, 13)
        pos, vel, ang = rows[..., 0:3], rows[..., 7:10], rows[..., 10:13]
        mv = meff[..., None] * vel
        return mv.sum(1), (np.cross(pos, mv) + ang).sum(1)  # unit effective inertias

    # float64 arithmetic on the float32 TABLE: its frame quaternions are unit only to 1e-7, so a torque pair reaches
    # child and parent scaled by (1 +- 1e-7) -- the residual is ~1e-6 with 150 N m actuators, 1e-8 without actions
    for f64, tol in ((True, 2e-5 * scale), (False, 3e-3 * scale)):
        ora = OracleBraxEnv(sysd, ctx, autoreset=False, f64=f64)
        ora.init_from_q(q, qd)
        p0, l0 = momenta(ora.state)
        ora.step(a)
        p1, l1 = momenta(ora.state)
        np.testing.assert_allclose(p1 - p0, want_dp, rtol=0, atol=tol)
        np.testing.assert_allclose((l1 - l0)[:, 2], 0.0, atol=tol)
    st, ob = hc.init(sysd, q, qd, ctx)
    p0, l0 = momenta(st)
    hc.step(sysd, st, ctx, a, np.zeros(n, np.int32), 0, 0, st.copy(), ob.copy())
    p1, l1 = momenta(st)
    np.testing.assert_allclose(p1 - p0, want_dp, rtol=0, atol=3e-3 * scale)
    np.testing.assert_allclose((l1 - l0)[:, 2], 0.0, atol=3e-3 * scale)


@pytest.mark.parametrize("body", ["ant", "halfcheetah", "hopper", "walker2d", "humanoid"])
def test_free_fall_follows_the_semi_implicit_euler_closed_form(hc, body):
    """Independent anchor for the integrator, the env's substep schedule and the gravity context: a body released at
    rest in its rest pose, far above the ground, with zero actions, falls rigidly -- after one env-step of n_frames
    substeps of length dt every link has dropped by g dt^2 n (n + 1) / 2 and moves at g dt n (semi-implicit Euler:
    v += g dt; x += v dt), with dt and n_frames the values brax documents for the spring backend (Ant 0.005 x 10,
    Halfcheetah 0.003125 x 16, Hopper / Walker2d 0.002 x 4). Oracle (float64 and float32) and kernel source."""
    sysd = bs.SYSTEMS[body]
    dt_n = {"ant": (0.005, 10), "halfcheetah": (0.003125, 16), "hopper": (0.002, 4), "walker2d": (0.002, 4),
            "humanoid": (0.0015, 10)}[body]
    n = 4
    rng = np.random.default_rng(3)
    ctx = random_ctx(sysd, n, rng)          # per-env gravity in [-15, -5], random masses
    ctx[:, 3] = 0.0                          # no angular damping
    nq = sysd["n_q"]
    q = np.tile(sysd["table"][bs.OFF_INIT_Q:bs.OFF_INIT_Q + nq].astype(np.float32), (n, 1))
    zi = 2 if body in ("ant", "humanoid") else 1           # free root: q = (x, y, z, quat...); planar root: (x, z, pitch)
    # (float32 positions 40 m up round to 4e-6 m, which the humanoid's 27 000 N/m joint springs turn into visible
    # forces: it is dropped from 4 m and held to 1e-4)
    q[:, zi] += 4.0 if body == "humanoid" else 40.0
    if body == "humanoid":  # the MJCF's zero pose has the knees outside their range (-160, -2) deg: bend them a little
        for name in ("right_shin", "left_shin"):
            o = bs.OFF_LINKS + bs.LINK_STRIDE * sysd["link_names"].index(name)
            q[:, int(sysd["table"][o + bs.L_QIDX])] = -0.1
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

    # humanoid, float64: its float32 table places the anchors of a 27 000 N/m joint spring ~1e-8 m apart
    tol32 = 1e-4 if body == "humanoid" else 2e-5
    for f64, tol in ((True, 1e-6 if body == "humanoid" else 1e-9), (False, tol32)):
        ora = OracleBraxEnv(sysd, ctx, autoreset=False, f64=f64)
        ora.init_from_q(q, qd)
        z0, _ = link_z(ora.state)
        ora.step(a)
        z1, vz1 = link_z(ora.state)
        np.testing.assert_allclose(z1 - z0, np.repeat(drop[:, None], L, 1), rtol=0, atol=max(tol, 1e-7 * 40))
        np.testing.assert_allclose(vz1, np.repeat(speed[:, None], L, 1), rtol=0, atol=max(tol, 1e-7))
    st, ob = hc.init(sysd, q, qd, ctx)
    z0, _ = link_z(st)
    hc.step(sysd, st, ctx, a, np.zeros(n, np.int32), 0, 0, st.copy(), ob.copy())
    z1, vz1 = link_z(st)
    np.testing.assert_allclose(z1 - z0, np.repeat(drop[:, None], L, 1), rtol=0, atol=tol32)
    np.testing.assert_allclose(vz1, np.repeat(speed[:, None], L, 1), rtol=0, atol=tol32)
