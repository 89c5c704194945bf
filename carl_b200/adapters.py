"""Adapters for the callers on the far side of the step path.

The reference is trained through ``gymnasium.wrappers.FlattenObservation`` + Stable-Baselines3
(``examples/carl_with_sb3.py:22-28``): SB3 wraps the single env into a ``DummyVecEnv`` and talks
to it through the ``VecEnv`` protocol (``reset() -> obs``, ``step_async`` / ``step_wait() ->
(obs, rewards, dones, infos)``, terminal observation in ``infos[i]["terminal_observation"]``, the
observation of a finished env already being the first one of its next episode). A batched
``CARLEnv`` with same-step autoreset IS that protocol, N envs wide, so the adapter is a thin
re-labelling: no physics, no copies beyond the device->host read SB3's numpy world needs.

Stable-Baselines3 / gymnasium are not imported (neither is installed here); the class is
duck-typed to the attributes and methods ``stable_baselines3.common.vec_env.VecEnv`` declares.
"""
from __future__ import annotations

from typing import Any, Sequence

import numpy as np
import torch

from carl_b200 import spaces

_NO_INFO: dict = {}


def flat_layout(env) -> list[tuple[str, int]]:
    """Column blocks of the flattened observation, in ``gymnasium.spaces.flatten`` order.

    ``gymnasium.spaces.Dict`` sorts its keys, so ``{"obs", "context"}`` (``carl_env.py:159-188``)
    flattens as context first, and a dict-valued context flattens in sorted feature-name order;
    a vector-valued context (``obs_context_as_dict=False``) keeps the order of
    ``obs_context_features``."""
    feats = list(env.obs_context_features)
    if env.obs_context_as_dict:
        feats = sorted(feats)
    return [(f"context/{k}", 1) for k in feats] + [("obs", int(env._info.obs_dim))]


def flatten_observation(env, state: dict[str, Any]):
    """``FlattenObservation`` of one batched CARL observation: ``[N, C + D]`` float32, on whatever
    device / array type the observation lives (torch tensor or numpy array)."""
    obs, ctx = state["obs"], state["context"]
    is_torch = isinstance(obs, torch.Tensor)
    cols = []
    if isinstance(ctx, dict):
        for k in sorted(ctx):
            c = ctx[k]
            if is_torch:
                cols.append(torch.as_tensor(c, dtype=torch.float32, device=obs.device).reshape(-1, 1))
            else:
                cols.append(np.asarray(c, dtype=np.float32).reshape(-1, 1))
    elif is_torch:
        cols.append(torch.as_tensor(ctx, dtype=torch.float32, device=obs.device).reshape(obs.shape[0], -1))
    else:
        cols.append(np.asarray(ctx, dtype=np.float32).reshape(obs.shape[0], -1))
    cols.append(obs if is_torch else np.asarray(obs, dtype=np.float32))
    return torch.cat(cols, dim=1) if is_torch else np.concatenate(cols, axis=1)


class SB3VecEnv:
    """``VecEnv``-protocol view of a batched ``CARLEnv`` (which must run with ``autoreset=True``).

    * ``observation_space``: the flattened single-env ``Box`` (``flatten=True``, what
      ``FlattenObservation`` yields) or the ``Dict`` space;
    * ``step_wait``: ``dones = terminated | truncated``; for every finished env
      ``infos[i] = {"terminal_observation": <flattened final obs>, "TimeLimit.truncated": bool}``
      (SB3's bootstrap convention); all other entries share one empty dict, so a step over tens of
      thousands of envs does not build tens of thousands of dicts."""

    metadata = {"render_modes": []}
    render_mode = None

    def __init__(self, env, flatten: bool = True):
        if not env._autoreset:
            raise ValueError("SB3VecEnv needs a CARLEnv created with autoreset=True (VecEnv resets finished envs itself)")
        self.env = env
        self.num_envs = int(env.num_envs)
        self.flatten = bool(flatten)
        self.action_space = env.single_action_space
        if self.flatten:
            width = sum(w for _, w in flat_layout(env))
            hi = np.full(width, np.inf, dtype=np.float32)
            self.observation_space = spaces.Box(-hi, hi, dtype=np.float32)
        else:
            self.observation_space = env.observation_space  # Dict{"obs", "context"} of one env instance
        self._actions = None
        self.reset_infos: list[dict] = [_NO_INFO] * self.num_envs

    # ------------------------------------------------------------------ protocol
    def _to_host(self, state):
        if self.flatten:
            return flatten_observation(self.env, state).cpu().numpy()
        ctx = state["context"]
        ctx = {k: v.cpu().numpy() for k, v in ctx.items()} if isinstance(ctx, dict) else ctx.cpu().numpy()
        return {"obs": state["obs"].cpu().numpy(), "context": ctx}

    def seed(self, seed: int | None = None) -> Sequence[int | None]:
        self._seed = seed
        return [None if seed is None else seed + i for i in range(self.num_envs)]

    def reset(self):
        state, _ = self.env.reset(seed=getattr(self, "_seed", None))
        self._seed = None
        return self._to_host(state)

    def step_async(self, actions) -> None:
        a = torch.as_tensor(np.asarray(actions))
        if self.env._info.act_discrete:
            a = a.to(torch.int64)
        else:
            a = a.to(torch.float32).reshape(self.num_envs, -1)
        self._actions = a.to(self.env.device, non_blocking=True)

    def step_wait(self):
        state, rew, term, trunc, info = self.env.step(self._actions)
        done = (term | trunc)
        done_h, trunc_h, term_h = done.cpu().numpy(), trunc.cpu().numpy(), term.cpu().numpy()
        infos: list[dict] = [_NO_INFO] * self.num_envs
        idx = np.nonzero(done_h)[0]
        if idx.size:
            sel = torch.as_tensor(idx, device=self.env.device)
            final = info["final_observation"].index_select(0, sel)
            if self.flatten:
                ctx = state["context"]
                ctx_sel = ({k: v.index_select(0, sel) for k, v in ctx.items()} if isinstance(ctx, dict)
                           else ctx.index_select(0, sel))
                final = flatten_observation(self.env, {"obs": final, "context": ctx_sel})
            final = final.cpu().numpy()
            infos = list(infos)
            for j, i in enumerate(idx):
                infos[i] = {"terminal_observation": final[j], "TimeLimit.truncated": bool(trunc_h[i] and not term_h[i])}
        return self._to_host(state), rew.cpu().numpy(), done_h, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self) -> None:
        pass

    # --------------------------------------------------- the rest of the VecEnv surface
    def get_attr(self, attr_name: str, indices=None) -> list:
        v = getattr(self.env, attr_name)
        return [v] * len(self._indices(indices))

    def set_attr(self, attr_name: str, value, indices=None) -> None:
        setattr(self.env, attr_name, value)

    def env_method(self, method_name: str, *args, indices=None, **kwargs) -> list:
        r = getattr(self.env, method_name)(*args, **kwargs)
        return [r] * len(self._indices(indices))

    def env_is_wrapped(self, wrapper_class, indices=None) -> list[bool]:
        return [False] * len(self._indices(indices))

    def get_images(self):
        raise NotImplementedError("rendering is outside the batched-step path")

    def render(self, mode: str | None = None):
        raise NotImplementedError("rendering is outside the batched-step path")

    def _indices(self, indices) -> list[int]:
        if indices is None:
            return list(range(self.num_envs))
        return [indices] if isinstance(indices, int) else list(indices)

    @property
    def unwrapped(self):
        return self
