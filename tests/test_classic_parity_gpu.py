"""Parity tests proper: the CUDA path, called through the C ABI (via the CARLEnv host layer),
against the CPU oracle on the same seeded inputs. Tolerance: 1e-5 relative fp32 (north star),
done masks identical; integer streams (resets) bit-exact."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.classic import DEFAULTS, FEATURES, KINDS, OracleClassicEnv
from tests.util import done_margin, env_class, sample_actions, sample_context_table, sample_states

pytestmark = pytest.mark.gpu
KIND_LIST = list(KINDS)
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "gymnasium_known_answers.json")))


def make_env(kind, table, **kw):
    from carl_b200.envs import ContextTable

    return env_class(kind)(contexts=ContextTable(FEATURES[kind], table), device="cuda:0", **kw)


def to_dev(kind, a):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t


def sync_oracle_from_env(ora, env):
    ora.state[:] = env.state.cpu().numpy().astype(np.float64)
    ora.elapsed[:] = env._elapsed.cpu().numpy()
    ora.sbt[:] = env._sbt.cpu().numpy()


def test_known_answers_through_cuda_path():
    """The reference's only physics known answers (CartPole-v1), produced by the CUDA kernels."""
    for dtype in ("float32", "float64"):
        env = make_env("cartpole", np.array([DEFAULTS["cartpole"]]), dtype=dtype)
        obs, info = env.reset(seed=0)
        want = np.asarray(GOLD["carl_cartpole_reset_seed0_float64"])
        np.testing.assert_array_equal(obs["obs"].cpu().numpy()[0], want.astype(np.float32))
        assert info["context_id"] == 0
        g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(0)))
        s0 = g.uniform(-0.05, 0.05, 4)
        env.state.copy_(torch.from_numpy(s0[None]).to(env.state.dtype))
        obs, r, te, tr, _ = env.step(torch.tensor([1], dtype=torch.int32, device="cuda"))
        np.testing.assert_allclose(obs["obs"].cpu().numpy()[0], np.asarray(GOLD["cartpole_seed0_step_action1"], np.float32),
                                   rtol=1e-6, atol=1e-7)
        assert r.item() == 1.0 and not te.item() and not tr.item()


@pytest.mark.parametrize("kind", KIND_LIST)
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_single_step_parity(kind, dtype):
    n = 8192
    f64 = dtype == "float64"
    rng = np.random.default_rng(10)
    table = sample_context_table(kind, n, rng, f32=not f64)
    states = sample_states(kind, n, rng, f32=not f64)
    actions = sample_actions(kind, n, rng)
    env = make_env(kind, table, dtype=dtype)
    env.reset(seed=0)
    env.state.copy_(torch.from_numpy(states).to(env.state.dtype))
    ora = OracleClassicEnv(kind, table)
    ora.state[:] = states
    o_ref, r_ref, t_ref, tr_ref, _ = ora.step(actions)
    obs, r, te, tr, info = env.step(to_dev(kind, actions))
    o = obs["obs"].cpu().numpy()
    rtol, atol = (2e-7, 1e-9) if f64 else (1e-5, 2e-6)
    np.testing.assert_allclose(o, o_ref, rtol=rtol, atol=atol)
    np.testing.assert_allclose(r.cpu().numpy(), r_ref, rtol=max(rtol, 1e-6), atol=max(atol, 1e-7))
    mism = te.cpu().numpy() != t_ref
    if mism.any():
        assert not f64
        assert done_margin(kind, ora.state, table)[mism].max() < 1e-6 and mism.sum() <= 2
    assert (tr.cpu().numpy() == tr_ref).all()
    if f64:
        np.testing.assert_allclose(env.state.cpu().numpy(), ora.state, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("kind", KIND_LIST)
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_reset_bit_exact(kind, dtype):
    n = 257  # ragged: not a multiple of the block size
    f64 = dtype == "float64"
    rng = np.random.default_rng(11)
    table = sample_context_table(kind, n, rng, f32=not f64)
    env = make_env(kind, table, dtype=dtype)
    ora = OracleClassicEnv(kind, table)
    obs, _ = env.reset(seed=321)
    o_ref = ora.reset(seed=321)
    np.testing.assert_array_equal(obs["obs"].cpu().numpy(), o_ref)
    want = ora.state if f64 else ora.state.astype(np.float32)
    np.testing.assert_array_equal(env.state.cpu().numpy(), want)
    # unseeded second reset continues every env's own stream; masked reset touches only the mask
    mask = rng.random(n) < 0.5
    obs2, _ = env.reset(mask=mask)
    o_ref2 = ora.reset(mask=mask)
    got = obs2["obs"].cpu().numpy()
    np.testing.assert_array_equal(got[mask], o_ref2[mask])
    np.testing.assert_array_equal(got[~mask], o_ref[~mask])


@pytest.mark.parametrize("kind", KIND_LIST)
def test_long_rollout_fp64_autoreset(kind):
    """Reference-precision mode: 400 steps with terminations, TimeLimit truncations, on-device
    auto-resets from each env's PCG64 stream; done masks identical at every step."""
    n, T = 64, 400
    rng = np.random.default_rng(12)
    table = sample_context_table(kind, n, rng, f32=False)
    if kind == "acrobot":
        table[: n // 2, 9] = 0.2
    max_steps = min(KINDS[kind]["max_steps"], 150)
    env = make_env(kind, table, dtype="float64", autoreset=True, max_episode_steps=max_steps)
    ora = OracleClassicEnv(kind, table, max_steps=max_steps)
    obs, _ = env.reset(seed=5)
    np.testing.assert_array_equal(obs["obs"].cpu().numpy(), ora.reset(seed=5))
    n_done = 0
    for t in range(T):
        a = sample_actions(kind, n, rng)
        o_ref, r_ref, t_ref, tr_ref, fin_ref = ora.step(a, autoreset=True)
        obs, r, te, tr, info = env.step(to_dev(kind, a))
        te, tr = te.cpu().numpy(), tr.cpu().numpy()
        assert (te == t_ref).all() and (tr == tr_ref).all(), f"done mismatch at step {t}"
        np.testing.assert_allclose(obs["obs"].cpu().numpy(), o_ref, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(r.cpu().numpy(), r_ref, rtol=1e-6, atol=1e-7)
        st = env.state.cpu().numpy()
        np.testing.assert_allclose(st, ora.state, rtol=1e-9, atol=1e-10)
        done = te | tr
        if done.any():
            np.testing.assert_array_equal(st[done], ora.state[done])
            np.testing.assert_allclose(info["final_observation"].cpu().numpy()[done], fin_ref[done], rtol=1e-6, atol=1e-7)
            n_done += int(done.sum())
    assert n_done > 0


@pytest.mark.parametrize("kind", KIND_LIST)
def test_short_rollout_fp32(kind):
    n, T = 1024, 32
    rng = np.random.default_rng(13)
    table = sample_context_table(kind, n, rng, f32=True)
    env = make_env(kind, table)
    ora = OracleClassicEnv(kind, table)
    env.reset(seed=4)
    ora.reset(seed=4)
    ora.state[:] = ora.state.astype(np.float32)
    alive = np.ones(n, dtype=bool)
    for t in range(T):
        a = sample_actions(kind, n, rng)
        o_ref, r_ref, t_ref, tr_ref, _ = ora.step(a)
        obs, r, te, tr, _ = env.step(to_dev(kind, a))
        tol = 1e-5 * (t + 1) * (30 if kind == "acrobot" else 4)
        np.testing.assert_allclose(obs["obs"].cpu().numpy()[alive], o_ref[alive], rtol=tol, atol=tol)
        alive &= (te.cpu().numpy() == t_ref) & ~t_ref
    assert alive.sum() > 0


@pytest.mark.parametrize("kind", KIND_LIST)
def test_fused_rollout_equals_stepwise(kind):
    """The K-step fused kernel (state in registers, in-kernel Philox policy) must reproduce the
    single-step kernel driven with the actions it recorded -- bit for bit, resets included."""
    n, K = 777, 96
    rng = np.random.default_rng(14)
    table = sample_context_table(kind, n, rng)
    max_steps = min(KINDS[kind]["max_steps"], 40)
    a_env = make_env(kind, table, autoreset=True, max_episode_steps=max_steps)
    b_env = make_env(kind, table, autoreset=True, max_episode_steps=max_steps)
    a_env.reset(seed=7)
    b_env.reset(seed=7)
    traj = a_env.rollout(K, policy_seed=99, record=True)
    torch.cuda.synchronize()
    for t in range(K):
        obs, r, te, tr, _ = b_env.step(traj["actions"][t])
        assert torch.equal(obs["obs"], traj["obs"][t]), f"obs differ at step {t}"
        assert torch.equal(r, traj["reward"][t])
        done = te.to(torch.uint8) | (tr.to(torch.uint8) << 1)
        assert torch.equal(done, traj["done"][t])
    assert torch.equal(a_env.state, b_env.state) and torch.equal(a_env._rng, b_env._rng)
    assert torch.equal(a_env._elapsed, b_env._elapsed)
    assert int((traj["done"] != 0).sum()) > 0
    # given actions instead of the in-kernel policy: same trajectory again
    c_env = make_env(kind, table, autoreset=True, max_episode_steps=max_steps)
    c_env.reset(seed=7)
    traj2 = c_env.rollout(K, actions=traj["actions"], record=True)
    assert torch.equal(traj2["obs"], traj["obs"]) and torch.equal(traj2["done"], traj["done"])


def test_policy_actions_are_sharding_invariant():
    """Philox policy keys on the GLOBAL env id: two half-shards reproduce the full batch."""
    kind, n, K = "cartpole", 512, 40
    rng = np.random.default_rng(15)
    table = sample_context_table(kind, n, rng)
    from carl_b200.envs import CARLCartPole, ContextTable

    full = make_env(kind, table, autoreset=True)
    full.reset(seed=3)
    tf = full.rollout(K, policy_seed=5, record=True)
    parts = []
    for r in range(2):
        e = CARLCartPole(contexts=ContextTable(FEATURES[kind], table), device="cuda:0", autoreset=True, shard=(r, 2))
        e.reset(seed=3)
        parts.append(e.rollout(K, policy_seed=5, record=True))
    assert torch.equal(torch.cat([parts[0]["obs"], parts[1]["obs"]], dim=1), tf["obs"])
    assert torch.equal(torch.cat([parts[0]["actions"], parts[1]["actions"]], dim=1), tf["actions"])


def test_mixed_batch_equals_separate_launches():
    """Config 3 shape: Pendulum + Acrobot shards stepped by ONE launch."""
    from carl_b200.envs.mixed import MixedBatch

    rng = np.random.default_rng(16)
    n = 3000
    envs_a = [make_env("pendulum", sample_context_table("pendulum", n, rng)),
              make_env("acrobot", sample_context_table("acrobot", n + 17, rng))]
    envs_b = [make_env("pendulum", envs_a[0].context_table.values), make_env("acrobot", envs_a[1].context_table.values)]
    for e in envs_a + envs_b:
        e.reset(seed=8)
    mixed = MixedBatch(envs_a)
    for t in range(5):
        acts = [to_dev("pendulum", sample_actions("pendulum", n, rng)), to_dev("acrobot", sample_actions("acrobot", n + 17, rng))]
        outs = mixed.step(acts)
        for e, a, out in zip(envs_b, acts, outs):
            obs, r, te, tr, _ = e.step(a)
            assert torch.equal(out[0]["obs"], obs["obs"]) and torch.equal(out[1], r)
            assert torch.equal(out[2], te) and torch.equal(out[3], tr)


def test_host_buffer_step_matches_device_step():
    """numpy in -> numpy out (carlb_env_step_host, pinned staging) equals the device path."""
    kind, n = "cartpole", 1000
    rng = np.random.default_rng(17)
    table = sample_context_table(kind, n, rng)
    a_env, b_env = make_env(kind, table), make_env(kind, table)
    a_env.reset(seed=1); b_env.reset(seed=1)
    for dt in (np.int64, np.int32, np.uint8):
        a = sample_actions(kind, n, rng).astype(dt)
        obs_h, r_h, te_h, tr_h, _ = a_env.step(a)
        obs_d, r_d, te_d, tr_d, _ = b_env.step(torch.from_numpy(a.astype(np.int32)).cuda())
        assert isinstance(obs_h["obs"], np.ndarray)
        np.testing.assert_array_equal(obs_h["obs"], obs_d["obs"].cpu().numpy())
        np.testing.assert_array_equal(r_h, r_d.cpu().numpy())
        np.testing.assert_array_equal(te_h, te_d.cpu().numpy())


def test_full_size_properties_config2():
    """BASELINE config 2 size (65 536 contexts): determinism, fused==stepwise checksum, done-mask
    consistency, no NaNs -- size-independent properties instead of a 65k-env oracle loop."""
    from carl_b200.context import ContextSampler, UniformFloatContextFeature
    from carl_b200.envs import CARLCartPole, ContextTable

    n = 65536
    names = FEATURES["cartpole"]
    def table():
        s = ContextSampler([UniformFloatContextFeature("gravity", 5, 15), UniformFloatContextFeature("length", 0.25, 1.0),
                            UniformFloatContextFeature("masscart", 0.5, 2.0)], CARLCartPole.get_context_space(), seed=0)
        return s.sample_context_table(n, names)
    e1 = CARLCartPole(contexts=ContextTable(names, table()), device="cuda:0", autoreset=True)
    e2 = CARLCartPole(contexts=ContextTable(names, table()), device="cuda:0", autoreset=True)
    e1.reset(seed=0); e2.reset(seed=0)
    t1 = e1.rollout(200, policy_seed=1, record=True)
    t2 = e2.rollout(200, policy_seed=1, record=True)
    assert torch.equal(t1["obs"], t2["obs"]) and torch.equal(t1["done"], t2["done"])  # deterministic
    assert torch.isfinite(t1["obs"]).all()
    term = (t1["done"] & 1).bool()
    # every terminated step's *pre-reset* state left the box; rewards are 1 on every step under autoreset
    assert torch.equal(t1["reward"], torch.ones_like(t1["reward"]))
    assert 0 < term.float().mean().item() < 0.2
    # oracle spot-check on a 512-env slice of the same batch (teacher-forced from the recorded actions)
    sl = slice(1000, 1512)
    ora = OracleClassicEnv("cartpole", e1.context_table.values[sl].astype(np.float32).astype(np.float64))
    e3 = CARLCartPole(contexts=ContextTable(names, table()[sl]), device="cuda:0", autoreset=True, num_envs=512)
    e3.reset(seed=1000)  # env i of the slice had global seed 0 + (1000 + i)
    ora.reset(seed=1000)
    np.testing.assert_array_equal(e3.state.cpu().numpy(), ora.state.astype(np.float32))


def test_small_angle_sincos_is_bit_identical_to_sincosf():
    """`m_sincos`'s short branch (|x| < 0.78: sincosf's polynomials without its argument reduction)
    against CUDA's sincosf for EVERY float in [-1, 1] -- run on the device by tests/devcheck."""
    import subprocess

    from tests.devcheck import build_devcheck

    p = subprocess.run([build_devcheck.build()], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.startswith("mismatches 0 checked 2130706434"), p.stdout


def test_trajectory_offsets_beyond_32_bits():
    """Maximum sizes: 4 Mi envs x 300 fused steps puts the last observation rows at element offsets
    > 2^32 of the trajectory buffer (19 GiB). The row written for step t must be what a second
    run, cut into two launches at an unaligned step, writes for the same global step."""
    from carl_b200.envs import CARLCartPole

    n, K, cut = 1 << 22, 300, 173
    a = CARLCartPole(num_envs=n, autoreset=True)
    a.reset(seed=7)
    ta = a.rollout(K, policy_seed=9, record=True)
    assert ta["obs"].numel() > (1 << 32)
    last = ta["obs"][-1].clone()
    mid = ta["obs"][cut - 1].clone()
    acts_tail = ta["actions"][-1].clone()
    done_sum = int((ta["done"] != 0).sum().item())
    assert torch.equal(last, a._obs) and torch.isfinite(last).all()
    del ta
    torch.cuda.empty_cache()
    b = CARLCartPole(num_envs=n, autoreset=True)
    b.reset(seed=7)
    t1 = b.rollout(cut, policy_seed=9, step_base=0, record=True)
    assert torch.equal(t1["obs"][-1], mid)
    d1 = int((t1["done"] != 0).sum().item())
    del t1
    torch.cuda.empty_cache()
    t2 = b.rollout(K - cut, policy_seed=9, step_base=cut, record=True)
    assert torch.equal(t2["obs"][-1], last) and torch.equal(t2["actions"][-1], acts_tail)
    assert d1 + int((t2["done"] != 0).sum().item()) == done_sum
    assert 0.02 < done_sum / (n * K) < 0.08   # random policy: one episode end every ~22 steps
