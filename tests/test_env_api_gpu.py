"""API-shape tests ported from the reference's test/test_CARLEnv.py, test_context_selector.py,
test_gymnasium_envs.py and the notebook behaviour (round-robin ids, context_id setter), run
against the batched CUDA-backed classes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pendulum_contexts():
    context = {"dt": 0.03, "gravity": 10.0, "m": 1.0, "l": 1.8}
    return {k: dict(context) for k in "abc"}


def test_observation_dict_and_context_lengths():
    from carl_b200.envs import CARLPendulum

    env = CARLPendulum()
    context = CARLPendulum.get_default_context()
    obs, info = env.reset()
    assert type(obs) is dict and "obs" in obs and "context" in obs
    assert len(obs["context"]) == len(context)
    assert obs["obs"].shape == (1, 3)
    env = CARLPendulum(obs_context_features=[])
    state, info = env.reset()
    assert len(state["context"]) == 0
    keys = list(context.keys())[:3]
    env = CARLPendulum(obs_context_features=keys)
    state, info = env.reset()
    assert len(state["context"]) == 3
    env = CARLPendulum(obs_context_as_dict=False)
    state, _ = env.reset()
    assert state["context"].shape == (1, len(context))


def test_selector_wiring():
    from carl_b200.context import RandomSelector, RoundRobinSelector
    from carl_b200.envs import CARLPendulum

    contexts = _pendulum_contexts()
    env = CARLPendulum(contexts=contexts, num_envs=1)
    env.reset()
    assert type(env.context_selector) is RoundRobinSelector and env.context_selector.n_calls == 1
    env.reset()
    assert env.context_selector.n_calls == 2
    assert type(CARLPendulum(contexts=contexts, context_selector=RoundRobinSelector(contexts=contexts)).context_selector) is RoundRobinSelector
    assert type(CARLPendulum(contexts=contexts, context_selector=RandomSelector(contexts=contexts)).context_selector) is RandomSelector
    assert type(CARLPendulum(contexts=contexts, context_selector=RandomSelector).context_selector) is RandomSelector
    with pytest.raises(ValueError):
        CARLPendulum(contexts=contexts, context_selector="bork")


def test_round_robin_ids_and_context_id_setter():
    """sample_contexts_with_brax.ipynb cells 7-11: first reset -> id 0, second -> 1, setter -> 4."""
    from carl_b200.envs import CARLCartPole

    contexts = {i: {"gravity": 9.8 + i} for i in range(5)}
    env = CARLCartPole(contexts=contexts, num_envs=1)
    _, info = env.reset()
    assert env.context_id == 0 and info["context_id"] == 0 and env.context["gravity"] == 9.8
    env.reset()
    assert env.context_id == 1 and env.context["gravity"] == 10.8
    env.context_id = 4
    assert env.context_id == 4 and env.context["gravity"] == 13.8
    assert env._ctx[0, 0].item() == pytest.approx(13.8)  # re-injected at once
    with pytest.raises(AssertionError):
        env.context_id = 17
    # contexts setter fills defaults (carl_env.py:135-137)
    assert env.contexts[3]["masscart"] == 1.0 and len(env.contexts[3]) == 8


def test_batch_binds_env_i_to_context_i_and_defaults():
    from carl_b200.envs import CARLCartPole

    contexts = {i: {"gravity": 5.0 + i} for i in range(6)}
    env = CARLCartPole(contexts=contexts)
    assert env.num_envs == 6
    obs, info = env.reset(seed=0)
    np.testing.assert_array_equal(info["context_id"], np.arange(6))
    np.testing.assert_allclose(obs["context"]["gravity"].cpu().numpy(), 5.0 + np.arange(6))
    np.testing.assert_allclose(env._ctx[0].cpu().numpy(), 5.0 + np.arange(6))
    # more envs than contexts: round robin wraps
    env = CARLCartPole(contexts=contexts, num_envs=8)
    _, info = env.reset(seed=0)
    np.testing.assert_array_equal(info["context_id"], np.arange(8) % 6)


@pytest.mark.parametrize("name", ["CARLCartPole", "CARLPendulum", "CARLAcrobot", "CARLMountainCar", "CARLMountainCarContinuous"])
def test_all_classic_envs_construct_reset_step(name):
    """test/test_gymnasium_envs.py + test_all_envs.py: features, construct, progress, update, reset."""
    import carl_b200.envs as E

    cls = getattr(E, name)
    cls.get_context_features()
    env = cls()
    env._progress_instance()
    env._update_context()
    obs, info = env.reset()
    assert obs["obs"].shape == (1, env._info.obs_dim)
    a = env.single_action_space.sample()
    out = env.step(np.asarray([a]))
    assert len(out) == 5 and out[0]["obs"].shape == (1, env._info.obs_dim)
    assert "context_id" in out[4]
    assert env.observation_space["obs"].shape == (env._info.obs_dim,)
    env.close()


def test_invalid_arguments_raise():
    from carl_b200.envs import CARLCartPole

    with pytest.raises(ValueError):
        CARLCartPole(contexts={0: {"not_a_feature": 1.0}})
    with pytest.raises(ValueError):
        CARLCartPole(dtype="float16")
    env = CARLCartPole(num_envs=4)
    env.reset(seed=0)
    with pytest.raises(AssertionError):
        env.step(torch.zeros(5, dtype=torch.int32, device="cuda"))
    with pytest.raises(ValueError):
        env._lib  # noqa: B018
        from carl_b200 import _native
        _native.check(env._lib.carlb_env_step(env._handle, torch.zeros(4, device="cuda").data_ptr(), _native.ACT_F32, None))


def test_order_enforcing_and_action_validation():
    from carl_b200.envs import CARLAcrobot

    env = CARLAcrobot(num_envs=3)
    with pytest.raises(RuntimeError, match="before calling env.reset"):
        env.step(np.zeros(3, dtype=np.int64))
    env.reset(seed=0)
    env.step(np.array([0, 1, 2]))
    with pytest.raises(AssertionError, match="invalid action"):
        env.step(np.array([0, 3, 1]))
    with pytest.raises(AssertionError, match="invalid action"):
        env.step(np.array([0, -1, 1]))
    CARLAcrobot(num_envs=3, validate_actions=False)


def test_checkpoint_roundtrip():
    from carl_b200.envs import CARLAcrobot

    env = CARLAcrobot(num_envs=64, autoreset=True)
    env.reset(seed=2)
    env.rollout(50, policy_seed=1)
    sd = env.state_dict()
    t1 = env.rollout(30, policy_seed=2, record=True)
    env2 = CARLAcrobot(num_envs=64, autoreset=True)
    env2.reset(seed=99)
    env2.load_state_dict(sd)
    t2 = env2.rollout(30, policy_seed=2, record=True)
    assert torch.equal(t1["obs"], t2["obs"]) and torch.equal(t1["done"], t2["done"])


def test_checkpoint_resume_with_fewer_contexts_than_envs_and_selection_state():
    """round-robin selection state (per-env reset counters, the selector's context_id / n_calls) is
    part of the checkpoint -- with M contexts != N envs an uninterrupted run and a resumed one must keep selecting
    the same contexts at the next resets."""
    from carl_b200.envs import CARLCartPole

    ctxs = {k: {"gravity": 5.0 + k, "length": 0.3 + 0.1 * k} for k in range(5)}   # M = 5 contexts
    env = CARLCartPole(contexts=ctxs, num_envs=12, autoreset=False)               # N = 12 envs
    env.reset(seed=4)
    env.reset()
    env.step(np.ones(12, dtype=np.int64))
    sd = env.state_dict()
    env.reset()
    ids_a, n_calls_a = env.context_id.copy(), env.context_selector.n_calls
    st_a = env.state.clone()
    env2 = CARLCartPole(contexts=ctxs, num_envs=12, autoreset=False)
    env2.load_state_dict(sd)
    env2.reset()
    np.testing.assert_array_equal(env2.context_id, ids_a)
    assert env2.context_selector.n_calls == n_calls_a
    assert torch.equal(env2.state, st_a)
    with pytest.raises(AssertionError, match="different env"):
        CARLCartPole(contexts=ctxs, num_envs=13).load_state_dict(sd)


def test_checkpoint_resume_brax_reset_stream_and_goal_state():
    """A resumed Brax env draws the same reset noise as the uninterrupted one (seed + per-env reset counters are
    restored) and keeps the goal wrapper's dead-reckoned positions."""
    from carl_b200.envs import CARLBraxAnt

    ctxs = {k: {"target_direction": d, "target_distance": 5.0 + k} for k, d in enumerate([1, 12, 3, 34])}
    env = CARLBraxAnt(contexts=ctxs, num_envs=4)
    assert env._goal_active
    env.reset(seed=3)
    a = torch.rand(4, 8, device="cuda") * 2 - 1
    for _ in range(3):
        env.step(a)
    sd = env.state_dict()
    env.step(a)
    pos_a = env._goal_state["position"].clone()
    env.reset()
    st_a = env.state.clone()
    env2 = CARLBraxAnt(contexts=ctxs, num_envs=4)
    env2.load_state_dict(sd)
    env2.step(a)
    assert torch.equal(env2._goal_state["position"], pos_a)
    env2.reset()
    assert torch.equal(env2.state, st_a)


def test_masked_goal_reset_keeps_the_other_positions():
    """reset(mask=...) re-zeroes the dead-reckoned position only of the envs being reset."""
    from carl_b200.envs import CARLBraxAnt

    ctxs = {k: {"target_direction": d, "target_distance": 5.0 + k} for k, d in enumerate([1, 12, 3, 34])}
    env = CARLBraxAnt(contexts=ctxs, num_envs=4)
    env.reset(seed=0)
    a = torch.rand(4, 8, device="cuda") * 2 - 1
    for _ in range(4):
        env.step(a)
    pos = env._goal_state["position"].clone()
    assert (pos.abs().sum(dim=1) > 0).all()
    mask = np.array([True, False, False, True])
    env.reset(mask=mask)
    now = env._goal_state["position"]
    assert torch.equal(now[1:3], pos[1:3]) and (now[0] == 0).all() and (now[3] == 0).all()


def test_discrete_env_accepts_float_cuda_actions_like_the_host_path():
    from carl_b200.envs import CARLCartPole

    a_env, b_env = CARLCartPole(num_envs=64), CARLCartPole(num_envs=64)
    a_env.reset(seed=1); b_env.reset(seed=1)
    acts = torch.randint(0, 2, (64,), device="cuda")
    oa, *_ = a_env.step(acts.to(torch.float32))
    ob, *_ = b_env.step(acts.to(torch.int32))
    assert torch.equal(oa["obs"], ob["obs"])


def test_page_locked_actions_are_read_in_place_and_match_staged_path():
    """`step(numpy)` with an array from carl_b200.hostmem.pinned_empty (no staging copy, the kernel reads
    it over PCIe) must give exactly what the same actions in an ordinary array give; the range check
    still fires for them."""
    from carl_b200 import hostmem
    from carl_b200.envs import CARLCartPole

    n = 3000
    a_env, b_env = CARLCartPole(num_envs=n), CARLCartPole(num_envs=n)
    a_env.reset(seed=11)
    b_env.reset(seed=11)
    rng = np.random.default_rng(0)
    pinned = hostmem.pinned_empty((5, n), np.int32)
    pinned[...] = rng.integers(0, 2, size=(5, n))
    for k in range(5):
        oa, ra, ta, tra, _ = a_env.step(pinned[k])
        ob, rb, tb, trb, _ = b_env.step(np.array(pinned[k]))  # pageable copy -> staged
        np.testing.assert_array_equal(oa["obs"], ob["obs"])
        np.testing.assert_array_equal(ra, rb)
        np.testing.assert_array_equal(ta, tb)
        np.testing.assert_array_equal(tra, trb)
    np.testing.assert_array_equal(a_env.state.cpu().numpy(), b_env.state.cpu().numpy())
    pinned[2, 7] = 2
    with pytest.raises(AssertionError, match="invalid action"):
        a_env.step(pinned[2])
    u8 = hostmem.pinned_empty((n,), np.uint8)
    u8[...] = 1
    oa, *_ = a_env.step(u8)
    ob, *_ = b_env.step(np.ones(n, dtype=np.int64))
    np.testing.assert_array_equal(oa["obs"], ob["obs"])
    hostmem.release(pinned)
    hostmem.release(u8)


@pytest.mark.parametrize("kind", ["cartpole", "acrobot", "mountaincar"])
def test_in_kernel_action_check_rolls_the_step_back(kind):
    """carlb_env_step_host_checked: with page-locked actions the step kernel validates them itself; one
    invalid action anywhere -> AssertionError and EVERY env is exactly where it was (state, step counters,
    PCG64 streams -- also for the envs that had already auto-reset inside the rejected step), as the
    reference's env is when gymnasium's `action_space.contains` assert fires."""
    import carl_b200.envs as E
    from carl_b200 import hostmem

    cls = {"cartpole": E.CARLCartPole, "acrobot": E.CARLAcrobot, "mountaincar": E.CARLMountainCar}[kind]
    n, n_act = 2048, {"cartpole": 2, "acrobot": 3, "mountaincar": 3}[kind]
    a_env = cls(num_envs=n, autoreset=True, max_episode_steps=7)   # short episodes: resets inside the rejected step
    b_env = cls(num_envs=n, autoreset=True, max_episode_steps=7)
    a_env.reset(seed=5)
    b_env.reset(seed=5)
    rng = np.random.default_rng(3)
    pinned = hostmem.pinned_empty((n,), np.int64)
    for t in range(20):
        acts = rng.integers(0, n_act, size=n)
        if t in (6, 13):  # the truncation step (every env resets) and an ordinary one
            before = (a_env.state.clone(), a_env._elapsed.clone(), a_env._rng.clone(), a_env._sbt.clone(),
                      a_env._obs.clone(), a_env._reward.clone(), a_env._terminated.clone(), a_env._truncated.clone())
            host_before = [np.array(x) for x in (oa["obs"], ra, ta, tra)] if t > 0 else None
            pinned[...] = acts
            pinned[n // 2] = n_act if t == 6 else -1
            with pytest.raises(AssertionError, match="invalid action"):
                a_env.step(pinned)
            after = (a_env.state, a_env._elapsed, a_env._rng, a_env._sbt, a_env._obs, a_env._reward, a_env._terminated,
                     a_env._truncated)
            for x, y in zip(before, after):
                assert torch.equal(x, y)
            if host_before is not None:  # the caller's arrays of the previous step show that step again
                for x, y in zip(host_before, (oa["obs"], ra, ta, tra)):
                    np.testing.assert_array_equal(x, y)
        pinned[...] = acts
        oa, ra, ta, tra, _ = a_env.step(pinned)
        ob, rb, tb, trb, _ = b_env.step(acts)  # pageable -> host check + staged copy
        np.testing.assert_array_equal(oa["obs"], ob["obs"])
        np.testing.assert_array_equal(ta, tb)
        np.testing.assert_array_equal(tra, trb)
    assert torch.equal(a_env.state, b_env.state) and torch.equal(a_env._rng, b_env._rng)
    hostmem.release(pinned)


@pytest.mark.parametrize("kind,parts", [("cartpole", 2), ("cartpole", 3), ("pendulum", 4), ("acrobot", 2)])
def test_split_batch_step_async_equals_step(kind, parts):
    """step_async / step_wait (split-batch host stepping: the parts' kernels run on their own streams, the host polls
    a completion word per part) produces exactly what the synchronous env.step(numpy) produces -- whole-batch calls,
    per-part calls in any order, pinned and pageable actions, auto-resets included."""
    import carl_b200.envs as E
    from carl_b200 import hostmem

    cls = {"cartpole": E.CARLCartPole, "pendulum": E.CARLPendulum, "acrobot": E.CARLAcrobot}[kind]
    n = 1000 + 7
    a_env = cls(num_envs=n, autoreset=True, max_episode_steps=9)
    b_env = cls(num_envs=n, autoreset=True, max_episode_steps=9)
    a_env.async_parts = parts
    a_env.reset(seed=11)
    b_env.reset(seed=11)
    rng = np.random.default_rng(5)
    discrete = a_env._info.act_discrete
    pinned = hostmem.pinned_empty((n,) if discrete else (n, 1), np.int32 if discrete else np.float32)
    bounds = [a_env.part_range(p) for p in range(parts)]
    assert bounds[0][0] == 0 and bounds[-1][1] == n
    for t in range(24):
        acts = rng.integers(0, a_env._info.n_actions, size=n).astype(np.int32) if discrete else \
            rng.uniform(-2, 2, size=(n, 1)).astype(np.float32)
        ob, rb, tb, trb, _ = b_env.step(acts)
        if t % 3 == 0:      # whole batch, pageable actions
            a_env.step_async(acts)
            oa, ra, ta, tra, info = a_env.step_wait()
            np.testing.assert_array_equal(oa["obs"], ob["obs"])
            np.testing.assert_array_equal(ra, rb)
            np.testing.assert_array_equal(ta, tb)
            np.testing.assert_array_equal(tra, trb)
        else:               # part by part, page-locked actions, waited for in reverse order
            pinned[...] = acts
            for p, (lo, hi) in enumerate(bounds):
                a_env.step_async(pinned[lo:hi], part=p)
            for p in reversed(range(parts)):
                lo, hi = bounds[p]
                oa, ra, ta, tra, info = a_env.step_wait(part=p)
                np.testing.assert_array_equal(oa["obs"], ob["obs"][lo:hi])
                np.testing.assert_array_equal(ra, rb[lo:hi])
                np.testing.assert_array_equal(ta, tb[lo:hi])
                np.testing.assert_array_equal(tra, trb[lo:hi])
                assert len(info["context_id"]) == hi - lo
    assert torch.equal(a_env.state, b_env.state) and torch.equal(a_env._rng, b_env._rng)
    with pytest.raises(RuntimeError, match="no step in flight"):
        a_env.step_wait(part=0)
    hostmem.release(pinned)


def test_split_batch_invalid_action_rolls_back_only_that_part():
    import carl_b200.envs as E
    from carl_b200 import hostmem

    n = 2048
    env = E.CARLCartPole(num_envs=n, autoreset=True, max_episode_steps=5)
    ref = E.CARLCartPole(num_envs=n, autoreset=True, max_episode_steps=5)
    env.async_parts = 2
    env.reset(seed=3)
    ref.reset(seed=3)
    pinned = hostmem.pinned_empty((n,), np.int32)
    rng = np.random.default_rng(0)
    for t in range(8):
        acts = rng.integers(0, 2, size=n).astype(np.int32)
        pinned[...] = acts
        if t == 4:  # every env truncates and resets in this step; part 1 carries an invalid action
            before = (env.state.clone(), env._elapsed.clone(), env._rng.clone())
            pinned[n - 3] = 7
            env.step_async(pinned[:n // 2], part=0)
            env.step_async(pinned[n // 2:], part=1)
            env.step_wait(part=0)
            with pytest.raises(AssertionError, match="invalid action"):
                env.step_wait(part=1)
            # part 0 stepped, part 1 is exactly where it was
            assert torch.equal(env.state[n // 2:], before[0][n // 2:])
            assert torch.equal(env._elapsed[n // 2:], before[1][n // 2:])
            assert torch.equal(env._rng[:, n // 2:], before[2][:, n // 2:])
            pinned[n - 3] = acts[n - 3]
            env.step_async(pinned[n // 2:], part=1)  # retry the rejected part with valid actions
            env.step_wait(part=1)
        else:
            env.step_async(pinned)
            env.step_wait()
        ref.step(acts)
    assert torch.equal(env.state, ref.state) and torch.equal(env._rng, ref._rng) and torch.equal(env._elapsed, ref._elapsed)
    hostmem.release(pinned)
