"""SB3 / FlattenObservation adapter (reference usage examples/carl_with_sb3.py:22-28)."""
import types

import numpy as np
import pytest
import torch

from carl_b200 import adapters


def _fake_env(as_dict):
    info = types.SimpleNamespace(obs_dim=3)
    return types.SimpleNamespace(obs_context_features=["length", "gravity"], obs_context_as_dict=as_dict, _info=info)


def test_flatten_order_matches_gymnasium_dict_sorting():
    """gymnasium.spaces.Dict sorts keys: "context" precedes "obs", dict-valued context features are
    sorted by name; a vector-valued context keeps the declared feature order."""
    obs = np.arange(6, dtype=np.float32).reshape(2, 3)
    env = _fake_env(True)
    flat = adapters.flatten_observation(env, {"obs": obs, "context": {"length": np.array([0.5, 0.6]), "gravity": np.array([9.8, 9.9])}})
    np.testing.assert_allclose(flat, [[9.8, 0.5, 0, 1, 2], [9.9, 0.6, 3, 4, 5]], rtol=1e-6)
    assert [k for k, _ in adapters.flat_layout(env)] == ["context/gravity", "context/length", "obs"]
    env = _fake_env(False)
    flat = adapters.flatten_observation(env, {"obs": torch.from_numpy(obs), "context": torch.tensor([[0.5, 9.8], [0.6, 9.9]])})
    np.testing.assert_allclose(flat.numpy(), [[0.5, 9.8, 0, 1, 2], [0.6, 9.9, 3, 4, 5]], rtol=1e-6)
    assert [k for k, _ in adapters.flat_layout(env)] == ["context/length", "context/gravity", "obs"]


@pytest.mark.gpu
def test_sb3_vecenv_protocol_on_cartpole():
    from carl_b200.envs import CARLCartPole

    n = 512
    env = CARLCartPole(num_envs=n, autoreset=True, obs_context_features=["gravity", "length"])
    ref = CARLCartPole(num_envs=n, autoreset=True, obs_context_features=["gravity", "length"])
    venv = adapters.SB3VecEnv(env)
    assert venv.num_envs == n and venv.observation_space.shape == (6,) and venv.action_space.n == 2
    venv.seed(3)
    obs = venv.reset()
    ref_state, _ = ref.reset(seed=3)
    assert obs.shape == (n, 6) and obs.dtype == np.float32
    np.testing.assert_array_equal(obs[:, 2:], ref_state["obs"].cpu().numpy())
    np.testing.assert_allclose(obs[:, 0], 9.8, rtol=1e-6)   # context/gravity first (sorted), then length
    np.testing.assert_allclose(obs[:, 1], 0.5, rtol=1e-6)
    rng = np.random.default_rng(0)
    seen_done = 0
    for _ in range(40):
        a = rng.integers(0, 2, size=n)
        obs, rew, dones, infos = venv.step(a)
        r_state, r_rew, r_term, r_trunc, r_info = ref.step(torch.as_tensor(a, device="cuda"))
        np.testing.assert_array_equal(obs[:, 2:], r_state["obs"].cpu().numpy())
        np.testing.assert_array_equal(dones, (r_term | r_trunc).cpu().numpy())
        np.testing.assert_array_equal(rew, r_rew.cpu().numpy())
        assert len(infos) == n
        for i in np.nonzero(dones)[0]:
            seen_done += 1
            np.testing.assert_array_equal(infos[i]["terminal_observation"][2:], r_info["final_observation"][i].cpu().numpy())
            assert infos[i]["TimeLimit.truncated"] is False
            # the returned obs of a finished env is already the first obs of its next episode
            assert np.abs(obs[i, 2:]).max() <= 0.1 + 1e-6
        for i in np.nonzero(~dones)[0][:5]:
            assert infos[i] == {}
    assert seen_done > 20
    with pytest.raises(ValueError):
        adapters.SB3VecEnv(CARLCartPole(num_envs=4, autoreset=False))
