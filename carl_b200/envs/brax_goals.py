"""Positional goals and language goals for the Brax walkers.

Reference: ``carl/envs/brax/brax_walker_goal_wrapper.py`` -- ``BraxWalkerGoalWrapper`` (:53-140)
replaces the env reward by the progress towards a goal point that it dead-reckons from two
observation entries times ``dt``; ``BraxLanguageWrapper`` (:143-178) adds a goal sentence to the
observation. Activated by ``CARLBraxEnv.__init__`` when the contexts vary ``target_distance`` /
``target_direction`` (``carl/envs/brax/carl_brax_env.py:195-223``).

Here the same arithmetic runs batched over the env instances (torch tensors on the env's device,
or numpy for the host path); it is an epilogue of the step, not part of the physics kernels.
"""
from __future__ import annotations

import numpy as np

# brax_walker_goal_wrapper.py:6-12 (note: the halfcheetah indices point at joint velocities, not the
# root velocity -- SURVEY App. E B7 -- reproduced as is)
STATE_INDICES = {"ant": [13, 14], "humanoid": [22, 23], "halfcheetah": [14, 15], "hopper": [5, 6], "walker2d": [8, 9]}

# the wrapper reads `sys.opt.timestep` of the freshly loaded MJCF (:107-109): the XML timestep,
# not the spring-backend override
MJCF_TIMESTEP = {"ant": 0.01, "halfcheetah": 0.01, "hopper": 0.002, "walker2d": 0.002, "humanoid": 0.003}

_c, _s = np.cos(22.5 * np.pi / 180), np.sin(22.5 * np.pi / 180)
DIRECTION_VALUES = {  # brax_walker_goal_wrapper.py:70-105
    3: [0, -1], 1: [0, 1], 2: [1, 0], 4: [-1, 0],
    34: [-np.sqrt(0.5), -np.sqrt(0.5)], 14: [-np.sqrt(0.5), np.sqrt(0.5)], 32: [np.sqrt(0.5), -np.sqrt(0.5)],
    12: [np.sqrt(0.5), np.sqrt(0.5)],
    334: [-_c, -_s], 434: [-_s, -_c], 114: [-_c, _s], 414: [-_s, _c], 332: [_c, -_s], 232: [_s, -_c], 112: [_c, _s],
    212: [_s, _c],
}
DIRECTION_NAMES = {
    1: "north", 3: "south", 2: "east", 4: "west", 12: "north east", 32: "south east", 14: "north west",
    34: "south west", 112: "north north east", 332: "south south east", 114: "north north west",
    334: "south south west", 212: "east north east", 232: "east south east", 414: "west north west",
    434: "west south west",
}


def goal_wrapper_active(contexts: dict | None) -> bool:
    """``carl_brax_env.py:195-223``: goals are on when the context set varies the target (only
    increases relative to the FIRST context count -- the reference takes ``max(c - base)``)."""
    if contexts is None:
        return False
    keys = list(contexts.keys())
    first = contexts[keys[0]]
    if "target_distance" not in first and "target_direction" not in first:
        return False
    assert all("target_direction" in contexts[k] for k in keys), "All contexts must have a 'target_direction' key"
    assert all("target_distance" in contexts[k] for k in keys), "All contexts must have a 'target_distance' key"
    base_dir, base_dist = first["target_direction"], first["target_distance"]
    max_diff_dir = max(c["target_direction"] - base_dir for c in contexts.values())
    max_diff_dist = max(c["target_distance"] - base_dist for c in contexts.values())
    return bool(max_diff_dir > 0.1 or max_diff_dist > 0.1)


def goal_wrapper_active_table(names: list[str], values: np.ndarray) -> bool:
    """The same gate for contexts handed over as a dense ``ContextTable`` (columns ``names``): on when the
    table carries both target columns and either one rises above its first row's value by more than 0.1."""
    if "target_distance" not in names or "target_direction" not in names or len(values) == 0:
        return False
    d = values[:, names.index("target_direction")]
    r = values[:, names.index("target_distance")]
    return bool((d - d[0]).max() > 0.1 or (r - r[0]).max() > 0.1)


def goal_positions(directions: np.ndarray, distances: np.ndarray) -> np.ndarray:
    """``np.array(direction_values[target_direction]) * target_distance`` per env (:115-118)."""
    d = np.array([DIRECTION_VALUES[int(k)] for k in directions], dtype=np.float64)
    return d * np.asarray(distances, dtype=np.float64)[:, None]


def goal_step(position, goal, radius, vel_xy, dt):
    """One ``BraxWalkerGoalWrapper.step`` (:124-140) for a batch. Works on numpy arrays and torch
    tensors alike. Returns (new_position, direction_reward, reached)."""
    new_position = position + vel_xy * dt
    if hasattr(position, "detach"):  # torch
        import torch

        cur = torch.linalg.norm(goal - new_position, dim=1)
        prev = torch.linalg.norm(goal - position, dim=1)
        reward = torch.clamp(prev - cur, min=0)
        reached = cur.abs() <= radius
    else:
        cur = np.linalg.norm(goal - new_position, axis=1)
        prev = np.linalg.norm(goal - position, axis=1)
        reward = np.maximum(0, prev - cur)
        reached = np.abs(cur) <= radius
    return new_position, reward, reached


def goal_description(context: dict) -> str:
    """``BraxLanguageWrapper.get_goal_desc`` (:167-178)."""
    if "target_radius" in context.keys():
        return (f"The distance to the goal is {context['target_distance']}m "
                f"{DIRECTION_NAMES[int(context['target_direction'])]}. Move within {context['target_radius']} steps of the goal.")
    return f"Move {context['target_distance']}m {DIRECTION_NAMES[int(context['target_direction'])]}."
