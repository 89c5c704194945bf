"""Contextual classic-control envs (batched).

Feature tables (names, bounds, defaults) are those of the reference classes:
``carl/envs/gymnasium/classic_control/carl_cartpole.py:16-42``, ``carl_pendulum.py:16-39``,
``carl_acrobot.py:16-69``, ``carl_mountaincar.py:16-51``, ``carl_mountaincarcontinuous.py:16-48``.
``_kernel_params`` is the batched ``CARLGymnasiumEnv._update_context``
(``carl/envs/gymnasium/carl_gymnasium_env.py:75-77``): which attribute each feature pokes and
what the gymnasium env does (or, in ``context_mode="reference"``, does not do) with it.
"""
from __future__ import annotations

import numpy as np

from carl_b200.context.context_space import ContextFeature, UniformFloatContextFeature
from carl_b200.envs.carl_env import CARLEnv


def _uf(name, lower, upper, default):
    return UniformFloatContextFeature(name, lower=lower, upper=upper, default_value=default)


class CARLGymnasiumEnv(CARLEnv):
    """Family adapter (reference: ``carl_gymnasium_env.py:19-77``). ``env_name`` is kept for
    registration parity; the physics is libcarlb's."""

    env_name: str
    render_mode: str = "rgb_array"

    @staticmethod
    def _cols(table: np.ndarray, names: list[str], wanted: list[str]) -> np.ndarray:
        return np.stack([table[:, names.index(w)] for w in wanted], axis=1)


class CARLCartPole(CARLGymnasiumEnv):
    env_name: str = "CartPole-v1"
    kind = "cartpole"
    metadata = {"render.modes": ["human", "rgb_array"]}

    @staticmethod
    def get_context_features() -> dict[str, ContextFeature]:
        return {
            "gravity": _uf("gravity", 0.1, np.inf, 9.8),
            "masscart": _uf("masscart", 0.1, 10, 1.0),
            "masspole": _uf("masspole", 0.01, 1, 0.1),
            "length": _uf("length", 0.05, 5, 0.5),
            "force_mag": _uf("force_mag", 1, 100, 10.0),
            "tau": _uf("tau", 0.002, 0.2, 0.02),
            "initial_state_lower": _uf("initial_state_lower", -np.inf, np.inf, -0.1),
            "initial_state_upper": _uf("initial_state_upper", -np.inf, np.inf, 0.1),
        }

    @classmethod
    def kernel_params(cls, table, names, context_mode="reference", explicit=None):
        # rows: gravity, masspole, length, force_mag, tau, total_mass, polemass_length, lo, hi
        g, mc, mp, ln, fm, tau, lo, hi = (table[:, names.index(k)] for k in (
            "gravity", "masscart", "masspole", "length", "force_mag", "tau", "initial_state_lower", "initial_state_upper"))
        if context_mode == "reference":
            # CartPoleEnv.__init__ caches total_mass / polemass_length from its own defaults and
            # CARL's setattr never refreshes them: masscart is inert.
            total_mass = np.full_like(g, 0.1 + 1.0)
            polemass_length = np.full_like(g, 0.1 * 0.5)
        else:
            total_mass = mp + mc
            polemass_length = mp * ln
        return np.stack([g, mp, ln, fm, tau, total_mass, polemass_length, lo, hi], axis=1)


class CARLPendulum(CARLGymnasiumEnv):
    env_name: str = "Pendulum-v1"
    kind = "pendulum"
    metadata = {"render_modes": ["human", "rgb_array"]}

    @staticmethod
    def get_context_features() -> dict[str, ContextFeature]:
        return {
            "gravity": _uf("gravity", -np.inf, np.inf, 8.0),
            "dt": _uf("dt", 0, np.inf, 0.05),
            "g": _uf("g", 0, np.inf, 10),
            "m": _uf("m", 1e-6, np.inf, 1),
            "l": _uf("l", 1e-6, np.inf, 1),
            "initial_angle_max": _uf("initial_angle_max", 0, np.inf, np.pi),
            "initial_velocity_max": _uf("initial_velocity_max", 0, np.inf, 1),
        }

    @classmethod
    def kernel_params(cls, table, names, context_mode="reference", explicit=None):
        # `gravity` is a dead attribute on PendulumEnv (the live one is `g`): carl_pendulum.py:18-26
        g = table[:, names.index("g")]
        if context_mode == "applied":
            # intended semantics: a context that SETS `gravity` changes the physics. `explicit[i, j]` says
            # whether context i named feature j itself (a deliberate gravity=8.0 counts); without that
            # information (bare tables) a value other than the default 8.0 is taken as set.
            j = names.index("gravity")
            grav = table[:, j]
            chosen = explicit[:, j] if explicit is not None else grav != 8.0
            g = np.where(chosen, grav, g)
        rest = cls._cols(table, names, ["m", "l", "dt", "initial_angle_max", "initial_velocity_max"])
        return np.concatenate([g[:, None], rest], axis=1)


class CARLAcrobot(CARLGymnasiumEnv):
    env_name: str = "Acrobot-v1"
    kind = "acrobot"
    metadata = {"render.modes": ["human", "rgb_array"]}

    @staticmethod
    def get_context_features() -> dict[str, ContextFeature]:
        return {
            "LINK_LENGTH_1": _uf("LINK_LENGTH_1", 0.1, 10, 1),
            "LINK_LENGTH_2": _uf("LINK_LENGTH_2", 0.1, 10, 1),
            "LINK_MASS_1": _uf("LINK_MASS_1", 0.1, 10, 1),
            "LINK_MASS_2": _uf("LINK_MASS_2", 0.1, 10, 1),
            "LINK_COM_POS_1": _uf("LINK_COM_POS_1", 0, 1, 0.5),
            "LINK_COM_POS_2": _uf("LINK_COM_POS_2", 0, 1, 0.5),
            "LINK_MOI": _uf("LINK_MOI", 0.1, 10, 1),
            "MAX_VEL_1": _uf("MAX_VEL_1", 0.4 * np.pi, 40 * np.pi, 4 * np.pi),
            "MAX_VEL_2": _uf("MAX_VEL_2", 0.9 * np.pi, 90 * np.pi, 9 * np.pi),
            "torque_noise_max": _uf("torque_noise_max", -1, 1, 0),
            "INITIAL_ANGLE_LOWER": _uf("INITIAL_ANGLE_LOWER", -np.inf, np.inf, -0.1),
            "INITIAL_ANGLE_UPPER": _uf("INITIAL_ANGLE_UPPER", -np.inf, np.inf, 0.1),
            "INITIAL_VELOCITY_LOWER": _uf("INITIAL_VELOCITY_LOWER", -np.inf, np.inf, -0.1),
            "INITIAL_VELOCITY_UPPER": _uf("INITIAL_VELOCITY_UPPER", -np.inf, np.inf, 0.1),
        }

    @classmethod
    def kernel_params(cls, table, names, context_mode="reference", explicit=None):
        # LINK_LENGTH_2 is render-only in AcrobotEnv._dsdt
        return cls._cols(table, names, [
            "LINK_MASS_1", "LINK_MASS_2", "LINK_LENGTH_1", "LINK_COM_POS_1", "LINK_COM_POS_2", "LINK_MOI",
            "MAX_VEL_1", "MAX_VEL_2", "torque_noise_max", "INITIAL_ANGLE_LOWER", "INITIAL_ANGLE_UPPER",
            "INITIAL_VELOCITY_LOWER", "INITIAL_VELOCITY_UPPER"])


class CARLMountainCar(CARLGymnasiumEnv):
    env_name: str = "MountainCar-v0"
    kind = "mountaincar"
    metadata = {"render.modes": ["human", "rgb_array"]}

    @staticmethod
    def get_context_features() -> dict[str, ContextFeature]:
        return {
            "min_position": _uf("min_position", -np.inf, np.inf, -1.2),
            "max_position": _uf("max_position", -np.inf, np.inf, 0.6),
            "max_speed": _uf("max_speed", 0, np.inf, 0.07),
            "goal_position": _uf("goal_position", -np.inf, np.inf, 0.45),
            "goal_velocity": _uf("goal_velocity", -np.inf, np.inf, 0),
            "force": _uf("force", -np.inf, np.inf, 0.001),
            "gravity": _uf("gravity", 0, np.inf, 0.0025),
            "min_position_start": _uf("min_position_start", -np.inf, np.inf, -0.6),
            "max_position_start": _uf("max_position_start", -np.inf, np.inf, -0.4),
            "min_velocity_start": _uf("min_velocity_start", -np.inf, np.inf, 0),
            "max_velocity_start": _uf("max_velocity_start", -np.inf, np.inf, 0),
        }

    @classmethod
    def kernel_params(cls, table, names, context_mode="reference", explicit=None):
        return cls._cols(table, names, [
            "min_position", "max_position", "max_speed", "goal_position", "goal_velocity", "force", "gravity",
            "min_position_start", "max_position_start", "min_velocity_start", "max_velocity_start"])


class CARLMountainCarContinuous(CARLGymnasiumEnv):
    env_name: str = "MountainCarContinuous-v0"
    kind = "mountaincar_cont"
    metadata = {"render.modes": ["human", "rgb_array"]}

    @staticmethod
    def get_context_features() -> dict[str, ContextFeature]:
        return {
            "min_position": _uf("min_position", -np.inf, np.inf, -1.2),
            "max_position": _uf("max_position", -np.inf, np.inf, 0.6),
            "max_speed": _uf("max_speed", 0, np.inf, 0.07),
            "goal_position": _uf("goal_position", -np.inf, np.inf, 0.5),
            "goal_velocity": _uf("goal_velocity", -np.inf, np.inf, 0),
            "power": _uf("power", -np.inf, np.inf, 0.0015),
            "min_position_start": _uf("min_position_start", -np.inf, np.inf, -0.6),
            "max_position_start": _uf("max_position_start", -np.inf, np.inf, -0.4),
            "min_velocity_start": _uf("min_velocity_start", -np.inf, np.inf, 0),
            "max_velocity_start": _uf("max_velocity_start", -np.inf, np.inf, 0),
        }

    @classmethod
    def kernel_params(cls, table, names, context_mode="reference", explicit=None):
        return cls._cols(table, names, [
            "min_position", "max_position", "max_speed", "goal_position", "goal_velocity", "power",
            "min_position_start", "max_position_start", "min_velocity_start", "max_velocity_start"])
