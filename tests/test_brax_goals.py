"""Goal / language wrappers: the batched arithmetic against a literal
restatement of the reference wrapper's per-env loop (brax_walker_goal_wrapper.py:111-140), the
activation rule of carl_brax_env.py:195-223, and the sentences of BraxLanguageWrapper."""
import numpy as np
import pytest

from carl_b200.envs import brax_goals as bg


def reference_wrapper_rollout(goal_dir, goal_dist, radius, vels, dt):
    """One env, the reference's own statements."""
    direction_values = bg.DIRECTION_VALUES
    position = (0, 0)
    goal_position = np.array(direction_values[goal_dir]) * goal_dist
    out = []
    for v in vels:
        new_position = np.array(list(position)) + np.array([v[0], v[1]]) * dt
        current_distance_to_goal = np.linalg.norm(goal_position - new_position)
        previous_distance_to_goal = np.linalg.norm(goal_position - position)
        direction_reward = max(0, previous_distance_to_goal - current_distance_to_goal)
        position = new_position
        out.append((direction_reward, abs(current_distance_to_goal) <= radius))
    return out


def test_goal_step_matches_reference_loop():
    rng = np.random.default_rng(0)
    n, T, dt = 16, 30, 0.01
    dirs = rng.choice(list(bg.DIRECTION_VALUES), n)
    dist = rng.uniform(0.05, 0.5, n)
    radius = rng.uniform(0.01, 0.1, n)
    vels = rng.normal(0, 3.0, (T, n, 2))
    pos = np.zeros((n, 2))
    goal = bg.goal_positions(dirs, dist)
    got = []
    for t in range(T):
        pos, r, reached = bg.goal_step(pos, goal, radius, vels[t], dt)
        got.append((r.copy(), reached.copy()))
    for i in range(n):
        ref = reference_wrapper_rollout(int(dirs[i]), dist[i], radius[i], vels[:, i], dt)
        for t in range(T):
            assert got[t][0][i] == pytest.approx(ref[t][0], abs=1e-12)
            assert bool(got[t][1][i]) == bool(ref[t][1])
            assert got[t][0][i] >= 0  # the reference's own assertion (test_language_goals.py:140-166)


def test_goal_step_torch_equals_numpy():
    import torch

    rng = np.random.default_rng(1)
    n = 8
    pos, goal, rad, vel = np.zeros((n, 2)), rng.normal(0, 1, (n, 2)), rng.uniform(0.1, 1, n), rng.normal(0, 1, (n, 2))
    p1, r1, d1 = bg.goal_step(pos, goal, rad, vel, 0.01)
    p2, r2, d2 = bg.goal_step(torch.from_numpy(pos), torch.from_numpy(goal), torch.from_numpy(rad), torch.from_numpy(vel), 0.01)
    np.testing.assert_allclose(p2.numpy(), p1)
    np.testing.assert_allclose(r2.numpy(), r1)
    assert (d2.numpy() == d1).all()


def test_activation_rule():
    base = {"target_distance": 10.0, "target_direction": 1}
    assert not bg.goal_wrapper_active(None)
    assert not bg.goal_wrapper_active({0: {"gravity": -9.8}})
    assert not bg.goal_wrapper_active({0: dict(base), 1: dict(base)})
    assert bg.goal_wrapper_active({0: dict(base), 1: dict(base, target_distance=10.2)})
    assert bg.goal_wrapper_active({0: dict(base), 1: dict(base, target_direction=3)})
    # reference quirk: only increases relative to the FIRST context count
    assert not bg.goal_wrapper_active({0: dict(base), 1: dict(base, target_distance=5.0)})
    with pytest.raises(AssertionError):
        bg.goal_wrapper_active({0: dict(base), 1: {"target_distance": 3.0}})


def test_goal_descriptions():
    c = {"target_distance": 8.957275946170714, "target_direction": 112, "target_radius": 5.0}
    s = bg.goal_description(c)
    assert "8.957275946170714m" in s and "north north east" in s and "5.0 steps" in s
    assert bg.goal_description({"target_distance": 3, "target_direction": 4}) == "Move 3m west."
    assert set(bg.DIRECTION_VALUES) == set(bg.DIRECTION_NAMES)
    for v in bg.DIRECTION_VALUES.values():
        assert np.hypot(*v) == pytest.approx(1.0)


@pytest.mark.gpu
def test_goal_wrapper_on_device_rewards_nonnegative():
    """test/test_language_goals.py:129-166 shape: sampled target contexts, 10 x 10 random steps."""
    from carl_b200.context import CategoricalContextFeature, ContextSampler, NormalFloatContextFeature
    from carl_b200.envs import CARLBraxAnt, CARLBraxHalfcheetah
    from carl_b200.envs.brax import directions

    for cls in (CARLBraxAnt, CARLBraxHalfcheetah):
        sampler = ContextSampler(
            [NormalFloatContextFeature("target_distance", mu=9.8, sigma=1, upper=50, lower=-40),
             CategoricalContextFeature("target_direction", choices=directions)], context_space=cls.get_context_space(), seed=0)
        contexts = sampler.sample_contexts(n_contexts=10)
        env = cls(contexts=contexts, use_language_goals=True)
        assert env._goal_active and env.num_envs == 10
        for _ in range(3):
            state, info = env.reset()
            assert isinstance(state["obs"], dict) and len(state["obs"]["goal"]) == 10
            assert "Move within" in state["obs"]["goal"][0]
            for _ in range(10):
                a = np.stack([env.single_action_space.sample() for _ in range(10)])
                state, reward, te, tr, info = env.step(a)
                assert (reward >= 0).all() and "success" in info
    env = CARLBraxAnt()  # no varying target -> plain env reward
    assert not env._goal_active


@pytest.mark.gpu
@pytest.mark.parametrize("cls_name,body", [("CARLBraxAnt", "ant"), ("CARLBraxHopper", "hopper"), ("CARLBraxHumanoid", "humanoid")])
def test_goal_kernel_epilogue_matches_reference_wrapper(cls_name, body):
    """carlb_brax_goal_step (one launch after the step, float64) against the reference wrapper's own per-env
    statements driven by the SAME observations: reward to 1e-12, terminated / success identical; the goal
    state survives Brax auto-resets (the wrapper sits outside AutoResetWrapper) and is zeroed by reset()."""
    import torch

    import carl_b200.envs as E

    cls = getattr(E, cls_name)
    rng = np.random.default_rng(0)
    n, T = 64, 40
    dirs = rng.choice(list(bg.DIRECTION_VALUES), n)
    ctxs = {i: dict(cls.get_default_goal_context(), target_direction=int(dirs[i]), target_distance=float(0.02 + 0.02 * i),
                    target_radius=0.1) for i in range(n)}
    env = cls(contexts=ctxs, max_episode_steps=15)
    assert env._goal_active
    obs, info = env.reset(seed=0)
    assert (info["success"] == 0).all()
    idx, dt = bg.STATE_INDICES[body], bg.MJCF_TIMESTEP[body]
    pos = [(0, 0)] * n
    goal = [np.array(bg.DIRECTION_VALUES[int(dirs[i])]) * ctxs[i]["target_distance"] for i in range(n)]
    n_reached = 0
    for t in range(T):
        a = torch.from_numpy((rng.uniform(-1, 1, (n, env._info.act_dim)) * env._sysd["act_scale"]).astype(np.float32)).cuda()
        obs, r, te, tr, info = env.step(a)
        o = obs["obs"].cpu().numpy()
        brax_done = env._elapsed.cpu().numpy() == 0  # auto-reset this step (time limit / unhealthy)
        r, te, succ = r.cpu().numpy(), te.cpu().numpy(), info["success"].cpu().numpy()
        assert r.dtype == np.float64
        for i in range(n):
            new_position = np.array(list(pos[i])) + np.array([o[i, idx[0]], o[i, idx[1]]]) * dt
            cur = np.linalg.norm(goal[i] - new_position)
            prev = np.linalg.norm(goal[i] - pos[i])
            pos[i] = new_position
            assert r[i] == pytest.approx(max(0, prev - cur), abs=1e-12)
            reached = abs(cur) <= 0.1
            assert bool(succ[i]) == reached
            assert bool(te[i]) == (reached or bool(brax_done[i]))
            n_reached += int(reached)
    assert n_reached > 0
    np.testing.assert_allclose(env._goal_state["position"].cpu().numpy(), np.array(pos), atol=1e-12)
    # numpy actions -> numpy results through the same kernel
    obs, r, te, tr, info = env.step(rng.uniform(-1, 1, (n, env._info.act_dim)).astype(np.float32))
    assert isinstance(r, np.ndarray) and r.dtype == np.float64 and isinstance(info["success"], np.ndarray)
    env.reset()
    assert float(env._goal_state["position"].abs().max()) == 0.0


@pytest.mark.gpu
def test_language_goal_observation():
    import carl_b200.envs as E

    ctxs = {i: dict(E.CARLBraxAnt.get_default_goal_context(), target_direction=d, target_distance=10.0 + i)
            for i, d in enumerate((1, 112, 4))}
    env = E.CARLBraxAnt(contexts=ctxs, use_language_goals=True)
    obs, info = env.reset(seed=0)
    assert set(obs["obs"]) == {"obs", "goal"} and len(obs["obs"]["goal"]) == 3
    assert obs["obs"]["goal"][1] == bg.goal_description(ctxs[1])
    obs, r, te, tr, info = env.step(np.zeros((3, 8), np.float32))
    assert obs["obs"]["goal"][2] == bg.goal_description(ctxs[2]) and r.shape == (3,)
