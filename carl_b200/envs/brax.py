"""Contextual Brax-locomotion envs (batched, spring pipeline in CUDA).

Feature tables are those of the reference classes: ``carl/envs/brax/carl_ant.py:20-49``,
``carl_halfcheetah.py:20-67``, ``carl_hopper.py:20-58``; ``directions`` is
``carl/envs/brax/brax_walker_goal_wrapper.py:33-50``.
"""
from __future__ import annotations

import numpy as np

from carl_b200.context.context_space import (
    CategoricalContextFeature,
    ContextFeature,
    UniformFloatContextFeature,
)
from carl_b200.envs.carl_env import CARLEnv
from carl_b200.utils.types import Context

directions = [1, 3, 2, 4, 12, 32, 14, 34, 112, 332, 114, 334, 212, 232, 414, 434]

DIRECTION_NAMES = {
    1: "north", 3: "south", 2: "east", 4: "west", 12: "north east", 32: "south east", 14: "north west",
    34: "south west", 112: "north north east", 332: "south south east", 114: "north north west",
    334: "south south west", 212: "east north east", 232: "east south east", 414: "west north west",
    434: "west south west",
}

# Context features CARLBraxEnv._update_context can apply (carl_brax_env.py:259-268) + every mass_<link>
REGISTERED_CFS = ["friction", "ang_damping", "gravity", "viscosity", "elasticity", "target_distance",
                  "target_direction", "target_radius"]


def _uf(name, lower, upper, default):
    return UniformFloatContextFeature(name, lower=lower, upper=upper, default_value=default)


def _common_features() -> dict[str, ContextFeature]:
    return {
        "gravity": _uf("gravity", -1000, -1e-6, -9.8),
        "friction": _uf("friction", 0, 100, 1),
        "elasticity": _uf("elasticity", 0, 100, 0),
        "ang_damping": _uf("ang_damping", -np.inf, np.inf, -0.05),
    }


def _goal_features() -> dict[str, ContextFeature]:
    return {
        "target_distance": _uf("target_distance", 0, np.inf, 100),
        "target_direction": CategoricalContextFeature("target_direction", choices=directions, default_value=1),
        "target_radius": _uf("target_radius", 0.1, np.inf, 5),
    }


def check_context(context: dict, registered_context_features: list[str]) -> None:
    """``carl_brax_env.py:104-112``."""
    for cfname in context.keys():
        if cfname not in registered_context_features and not cfname.startswith("mass_"):
            raise RuntimeError(
                f"Context feature {cfname} can not be updated in the brax system. Only "
                f"{registered_context_features} are possible."
            )


class CARLBraxEnv(CARLEnv):
    """Family adapter (reference: ``carl/envs/brax/carl_brax_env.py:115-336``)."""

    env_name: str
    backend: str = "spring"
    link_names: list[str] = []
    # `torso_mass` and `joint_stiffness` exist only in the reference's
    # stale v0 docs (docs/source/environments/data/context_definitions/CARLAnt.csv) but are what
    # BASELINE.json's configs name. `torso_mass` is an alias of `mass_torso`; `joint_stiffness` is an
    # extension: a per-env scale (1 = stock) of the spring backend's joint constraint stiffness.
    feature_aliases = {"torso_mass": "mass_torso"}
    extension_features = {"joint_stiffness": 1.0}

    def __init__(self, env=None, batch_size: int | None = None, contexts=None, obs_context_features=None,
                 obs_context_as_dict: bool = True, context_selector=None, context_selector_kwargs=None,
                 use_language_goals: bool = False, brax_tunables: dict | None = None, arithmetic: str = "strict",
                 reset_rng: str = "jax", **kwargs):
        """``carl_brax_env.py:119-236``. ``batch_size`` is the reference's name for the number of
        batched env instances (``brax.envs.create(batch_size=...)``, :163-167); here every instance
        may carry its own context."""
        if arithmetic not in ("strict", "fma"):
            raise ValueError(f"arithmetic must be 'strict' or 'fma', got {arithmetic!r}")
        # "strict": products and sums rounded separately (reproduces the float32 restatement of the reference
        # arithmetic to ~1e-6 per env-step: the parity mode). "fma": the same kernels built with FMA contraction,
        # as XLA compiles the reference's own -- faster, results at the float32 round-off floor of the algorithm.
        self.arithmetic = arithmetic
        if reset_rng not in ("jax", "philox"):
            raise ValueError(f"reset_rng must be 'jax' or 'philox', got {reset_rng!r}")
        # "jax": the reference's own reset-noise stream (JAX threefry2x32 keys consumed as wrappers.py / brax do:
        # PRNGKey(seed), `key1, key2 = split(key)` per reset, `split(key2, num_envs)[i]`, `split(rng, 3)`, uniform /
        # normal). Like the reference, an unseeded env draws from PRNGKey(0) (wrappers.py:41 `self.seed(0)`); unlike
        # it (`reset(seed=...)` is ignored there), an explicit seed re-keys the stream.
        # "philox": one Philox block per (seed, env, reset, index) -- the throughput-mode stream of round 1.
        self.reset_rng = reset_rng
        if batch_size is not None and batch_size != 1 and "num_envs" not in kwargs:
            kwargs["num_envs"] = int(batch_size)
        self.use_language_goals = use_language_goals
        from carl_b200.envs import brax_system as bs

        self._sysd = bs.build_system(bs.MODELS[self.env_name](), brax_tunables) if brax_tunables else bs.SYSTEMS[self.env_name]
        self.link_names = list(self._sysd["link_names"])
        from carl_b200.envs import brax_goals

        # carl_brax_env.py:195-223: goal wrapper only when the context set varies the target
        self._goal_active = brax_goals.goal_wrapper_active(contexts if isinstance(contexts, dict) else None)
        self._goal_state = None
        super().__init__(env=env, contexts=contexts, obs_context_features=obs_context_features,
                         obs_context_as_dict=obs_context_as_dict, context_selector=context_selector,
                         context_selector_kwargs=context_selector_kwargs, **kwargs)
        if self.context_mode == "reference" and len(self._table) > 1:
            vals, names_ = np.asarray(self._table.values), self._feature_names
            varying = [n for j, n in enumerate(names_)
                       if not n.startswith("target_") and np.ptp(vals[:, j]) > 0]
            if varying:
                import warnings

                warnings.warn(
                    f"{type(self).__name__}: the contexts vary {varying}, but context_mode='reference' reproduces the reference, "
                    "whose Brax context injection never reaches the physics (carl_brax_env.py:292 assigns the modified `sys` to "
                    "the gym shell, SURVEY §0.5): every env will step with the stock dynamics. Pass context_mode='applied' for "
                    "what the reference intends.", stacklevel=2)
        if contexts is not None and not isinstance(contexts, dict):
            # a dense ContextTable that varies the target behaves like the equivalent dict of contexts
            self._goal_active = brax_goals.goal_wrapper_active_table(list(contexts.names), np.asarray(contexts.values))

    def _default_autoreset(self) -> bool:
        return True  # brax.envs.create(auto_reset=True) wraps the env in AutoResetWrapper

    def _post_alloc(self) -> None:
        """Batched stand-in for ``mjcf.load`` (carl_brax_env.py:271-272): upload the system table."""
        import ctypes

        import numpy as np

        from carl_b200 import _native

        t = np.ascontiguousarray(self._sysd["table"], dtype=np.float32)
        self._sys_table_host = t
        _native.check(self._lib.carlb_brax_set_system(
            self._handle, t.ctypes.data_as(ctypes.c_void_p), int(t.size), 1 if self.context_mode == "reference" else 0))
        _native.check(self._lib.carlb_brax_set_arithmetic(self._handle, 1 if self.arithmetic == "fma" else 0))
        _native.check(self._lib.carlb_brax_set_reset_rng(self._handle, 1 if self.reset_rng == "jax" else 0,
                                                         self.global_num_envs))

    # ------------------------------------------------------------------ goal wrappers
    def _goal_reset(self, device_like, mask=None):
        """``BraxWalkerGoalWrapper.reset`` (brax_walker_goal_wrapper.py:111-122), batched. With a reset mask
        only the env instances being reset get a fresh position / goal / radius; the dead-reckoned position
        (hence the progress reward) of the others is left alone."""
        import torch

        from carl_b200.envs import brax_goals

        ids = self._context_ids
        names = self._feature_names
        vals = self._table.values[ids]
        goal = brax_goals.goal_positions(vals[:, names.index("target_direction")], vals[:, names.index("target_distance")])
        radius = vals[:, names.index("target_radius")]
        goal_t = torch.from_numpy(np.ascontiguousarray(goal, dtype=np.float64)).to(self.device).contiguous()
        radius_t = torch.from_numpy(np.ascontiguousarray(radius, dtype=np.float64)).to(self.device).contiguous()
        if mask is not None and self._goal_state is not None:
            m = torch.from_numpy(np.asarray(mask, dtype=bool)).to(self.device)
            g = self._goal_state
            g["position"][m] = 0.0
            g["goal"][m] = goal_t[m]
            g["radius"][m] = radius_t[m]
            g["success"][m] = 0
            g["reward"][m] = 0.0
        else:
            self._goal_state = dict(
                position=torch.zeros(self.num_envs, 2, dtype=torch.float64, device=self.device),
                goal=goal_t, radius=radius_t,
                reward=torch.zeros(self.num_envs, dtype=torch.float64, device=self.device),
                success=torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device),
                dt=brax_goals.MJCF_TIMESTEP[self.env_name],
                idx=brax_goals.STATE_INDICES[self.env_name],
            )
        self._goal_strings = None

    def _goal_strings_list(self):
        from carl_b200.envs import brax_goals

        if self._goal_strings is None:
            ctxs = self.contexts
            keys = self._table.keys
            self._goal_strings = [brax_goals.goal_description(ctxs[keys[int(i)]]) for i in self._context_ids]
        return self._goal_strings

    def reset(self, *, seed=None, options=None, mask=None):
        if seed is None and not self._seeded and self.reset_rng == "jax":
            seed = 0  # the reference's shells are keyed with PRNGKey(0) at construction (wrappers.py:41,80-81)
        state, info = super().reset(seed=seed, options=options, mask=mask)
        if self._goal_active:
            mask_np = None
            if mask is not None:
                import torch

                mask_np = (mask.detach().cpu().numpy() if isinstance(mask, torch.Tensor) else np.asarray(mask)).astype(bool)
            self._goal_reset(state["obs"], mask_np)
            info["success"] = self._goal_state["success"].to("cpu").numpy().astype(np.int64)
            if self.use_language_goals:
                state = {"obs": {"obs": state["obs"], "goal": self._goal_strings_list()}, "context": state["context"]}
        return state, info

    def step(self, action):
        import torch

        state, reward, te, tr, info = super().step(action)
        if not self._goal_active:
            return state, reward, te, tr, info
        from carl_b200.envs import brax_goals

        from carl_b200 import _native

        g = self._goal_state
        host = isinstance(state["obs"], np.ndarray)
        # one launch on the step's stream: dead reckoning, progress reward, radius termination (OR-ed into
        # the handle's terminated flags), success -- float64 like the reference wrapper's NumPy scalars
        _native.check(self._lib.carlb_brax_goal_step(
            self._handle, int(g["idx"][0]), int(g["idx"][1]), float(g["dt"]), g["position"].data_ptr(), g["goal"].data_ptr(),
            g["radius"].data_ptr(), g["reward"].data_ptr(), g["success"].data_ptr(), self._stream()))
        r, te_t = g["reward"], self._terminated_b
        info["success"] = g["success"].to(torch.int64)
        if host:
            reward, te = r.cpu().numpy(), te_t.cpu().numpy()
            info["success"] = info["success"].cpu().numpy()
        else:
            reward, te = r, te_t
        if self.use_language_goals:
            state = {"obs": {"obs": state["obs"], "goal": self._goal_strings_list()}, "context": state["context"]}
        return state, reward, te, tr, info

    def reset_from_q(self, q, qd, mask=None):
        """Parity-mode reset: ``pipeline_init(q, qd)`` from caller-supplied generalized coordinates
        (``float32[num_envs, n_q]`` / ``[num_envs, n_qd]``) instead of the noise draws of the
        reference's ``Env.reset`` (its JAX PRNG stream cannot be reproduced without JAX)."""
        import torch

        from carl_b200 import _native

        changed = self._progress_instance(None)
        if changed.any():
            self._update_context(changed if not changed.all() else None)
        qt = torch.as_tensor(q, dtype=torch.float32, device=self.device).contiguous()
        qdt = torch.as_tensor(qd, dtype=torch.float32, device=self.device).contiguous()
        assert qt.shape == (self.num_envs, self._sysd["n_q"]) and qdt.shape == (self.num_envs, self._sysd["n_qd"])
        mt = None if mask is None else torch.as_tensor(mask, dtype=torch.uint8, device=self.device)
        if not self._seeded:
            _native.check(self._lib.carlb_env_seed(self._handle, 0, self._stream()))
            self._seeded = True
        _native.check(self._lib.carlb_brax_reset_from_q(
            self._handle, None if mt is None else mt.data_ptr(), qt.data_ptr(), qdt.data_ptr(), self._stream()))
        self._has_reset = True
        return self._add_context_to_state(self._obs), {"context_id": self.context_id}

    @classmethod
    def get_default_context(cls) -> Context:
        """``carl_brax_env.py:308-324``: default context without the goal features."""
        default_context = cls.get_context_space().get_default_context()
        for k in ("target_distance", "target_direction", "target_radius"):
            default_context.pop(k, None)
        return default_context

    @classmethod
    def get_default_goal_context(cls) -> Context:
        """``carl_brax_env.py:326-336``."""
        return cls.get_context_space().get_default_context()

    @classmethod
    def kernel_params(cls, table, names, context_mode="reference", explicit=None):
        """Batched ``CARLBraxEnv._update_context`` (``carl_brax_env.py:255-292``).

        rows: gravity, friction, elasticity, ang_damping, joint-stiffness scale, mass_<link> per link.
        ``context_mode="reference"``: the reference assigns the modified ``sys`` to the gym shell,
        whose jitted step never reads it -- the physics sees the *stock* system
        (MJCF gravity/friction/masses) whatever the context says. ``"applied"``: the intended
        semantics, including the `viscosity` -> `ang_damping` overwrite (:278-279)."""
        from carl_b200.envs import brax_system as bs

        sysd = bs.SYSTEMS[cls.env_name]
        m = table.shape[0]
        rows = np.empty((m, 5 + len(sysd["link_names"])), dtype=np.float64)
        if context_mode == "reference":
            rows[:, 0] = sysd["stock_gravity"]
            rows[:, 1] = sysd["stock_friction"]
            rows[:, 2] = sysd["stock_elasticity"]
            rows[:, 3] = sysd["stock_ang_damping"]
            rows[:, 4] = 1.0
            rows[:, 5:] = np.asarray(sysd["stock_masses"])[None, :]
            return rows
        col = lambda k: table[:, names.index(k)]
        check_context({n: 0 for n in names if n not in cls.extension_features}, REGISTERED_CFS)
        rows[:, 0] = col("gravity")
        rows[:, 1] = col("friction")
        rows[:, 2] = col("elasticity")
        # "viscosity" in context overwrites ang_damping after "ang_damping" was applied (:276-279)
        rows[:, 3] = col("viscosity") if "viscosity" in names else col("ang_damping")
        rows[:, 4] = col("joint_stiffness") if "joint_stiffness" in names else 1.0
        for j, ln in enumerate(sysd["link_names"]):
            key = f"mass_{ln}"
            rows[:, 5 + j] = col(key) if key in names else sysd["stock_masses"][j]
        return rows


class CARLBraxAnt(CARLBraxEnv):
    env_name: str = "ant"
    kind = "brax_ant"
    asset_path: str = "envs/assets/ant.xml"
    metadata = {"render_modes": []}

    @staticmethod
    def get_context_features() -> dict[str, ContextFeature]:
        f = _common_features()
        f["mass_torso"] = _uf("mass_torso", 1e-6, np.inf, 10)
        f["viscosity"] = _uf("viscosity", 0, np.inf, 0)
        f.update(_goal_features())
        return f


class CARLBraxHalfcheetah(CARLBraxEnv):
    env_name: str = "halfcheetah"
    kind = "brax_halfcheetah"
    asset_path: str = "envs/assets/half_cheetah.xml"
    metadata = {"render_modes": []}

    @staticmethod
    def get_context_features() -> dict[str, ContextFeature]:
        f = _common_features()
        f["viscosity"] = _uf("viscosity", 0, np.inf, 0)
        f["mass_torso"] = _uf("mass_torso", 1e-6, np.inf, 10)
        f["mass_bthigh"] = _uf("mass_bthigh", 1e-6, np.inf, 1.5435146)
        f["mass_bshin"] = _uf("mass_bshin", 1e-6, np.inf, 1.5874476)
        f["mass_bfoot"] = _uf("mass_bfoot", 1e-6, np.inf, 1.0953975)
        f["mass_fthigh"] = _uf("mass_fthigh", 1e-6, np.inf, 1.4380753)
        f["mass_fshin"] = _uf("mass_fshin", 1e-6, np.inf, 1.2008368)
        f["mass_ffoot"] = _uf("mass_ffoot", 1e-6, np.inf, 0.8845188)
        f.update(_goal_features())
        return f


class CARLBraxHopper(CARLBraxEnv):
    env_name: str = "hopper"
    kind = "brax_hopper"
    asset_path: str = "envs/assets/hopper.xml"
    metadata = {"render_modes": []}

    @staticmethod
    def get_context_features() -> dict[str, ContextFeature]:
        f = _common_features()
        f["viscosity"] = _uf("viscosity", 0, np.inf, 0)
        f["mass_torso"] = _uf("mass_torso", 1e-6, np.inf, 10)
        f["mass_thigh"] = _uf("mass_thigh", 1e-6, np.inf, 4.0578904)
        f["mass_leg"] = _uf("mass_leg", 1e-6, np.inf, 2.7813568)
        f["mass_foot"] = _uf("mass_foot", 1e-6, np.inf, 5.3155746)
        f.update(_goal_features())
        return f


class CARLBraxWalker2d(CARLBraxEnv):
    """``carl/envs/brax/carl_walker2d.py:14-67`` (same kernels, a new system table)."""

    env_name: str = "walker2d"
    kind = "brax_walker2d"
    asset_path: str = "envs/assets/walker2d.xml"
    metadata = {"render_modes": []}

    @staticmethod
    def get_context_features() -> dict[str, ContextFeature]:
        f = _common_features()
        f["viscosity"] = _uf("viscosity", 0, np.inf, 0)
        f["mass_torso"] = _uf("mass_torso", 1e-6, np.inf, 10)
        f["mass_thigh"] = _uf("mass_thigh", 1e-6, np.inf, 4.0578904)
        f["mass_leg"] = _uf("mass_leg", 1e-6, np.inf, 2.7813568)
        f["mass_foot"] = _uf("mass_foot", 1e-6, np.inf, 3.1667254)
        f["mass_thigh_left"] = _uf("mass_thigh_left", 1e-6, np.inf, 4.0578904)
        f["mass_leg_left"] = _uf("mass_leg_left", 1e-6, np.inf, 2.7813568)
        f["mass_foot_left"] = _uf("mass_foot_left", 1e-6, np.inf, 3.1667254)
        f.update(_goal_features())
        return f


class CARLBraxInvertedPendulum(CARLBraxEnv):
    """``carl/envs/brax/carl_inverted_pendulum.py:9-39``: a new system table + the slide joint of the cart."""

    env_name: str = "inverted_pendulum"
    kind = "brax_inverted_pendulum"
    asset_path: str = "envs/assets/inverted_pendulum.xml"
    metadata = {"render_modes": []}

    @staticmethod
    def get_context_features() -> dict[str, ContextFeature]:
        f = _common_features()
        del f["ang_damping"]
        f["mass_cart"] = _uf("mass_cart", 1e-6, np.inf, 1)
        f["mass_pole"] = _uf("mass_pole", 1e-6, np.inf, 1)
        f["ang_damping"] = _uf("ang_damping", -np.inf, np.inf, -0.05)
        f["viscosity"] = _uf("viscosity", 0, np.inf, 0)
        return f


class CARLBraxInvertedDoublePendulum(CARLBraxEnv):
    """``carl/envs/brax/carl_inverted_double_pendulum.py:9-42``. Bug-compatible: the reference registers the
    ``mass_pole2`` entry with the feature NAME ``mass_pole`` (:33-35), so default contexts carry no ``mass_pole2``."""

    env_name: str = "inverted_double_pendulum"
    kind = "brax_inverted_double_pendulum"
    asset_path: str = "envs/assets/inverted_double_pendulum.xml"
    metadata = {"render_modes": []}

    @staticmethod
    def get_context_features() -> dict[str, ContextFeature]:
        f = _common_features()
        del f["ang_damping"]
        f["mass_cart"] = _uf("mass_cart", 1e-6, np.inf, 1)
        f["mass_pole"] = _uf("mass_pole", 1e-6, np.inf, 1)
        f["mass_pole2"] = _uf("mass_pole", 1e-6, np.inf, 1)
        f["ang_damping"] = _uf("ang_damping", -np.inf, np.inf, -0.05)
        f["viscosity"] = _uf("viscosity", 0, np.inf, 0)
        return f


class CARLBraxReacher(CARLBraxEnv):
    """``carl/envs/brax/carl_reacher.py:9-42``: two-link planar arm + a target body on two slide joints."""

    env_name: str = "reacher"
    kind = "brax_reacher"
    asset_path: str = "envs/assets/reacher.xml"
    metadata = {"render_modes": []}

    @staticmethod
    def get_context_features() -> dict[str, ContextFeature]:
        f = _common_features()
        f["viscosity"] = _uf("viscosity", 0, np.inf, 0)
        f["mass_body0"] = _uf("mass_body0", 1e-6, np.inf, 0.03560472)
        f["mass_body1"] = _uf("mass_body1", 1e-6, np.inf, 0.03979351)
        return f


def _humanoid_features() -> dict[str, ContextFeature]:
    """``carl/envs/brax/carl_humanoid.py:19-85`` (``carl_humanoidstandup.py:14-71`` lists the same physics and mass
    features without the goal block)."""
    f = _common_features()
    f["viscosity"] = _uf("viscosity", 0, np.inf, 0)
    for name, m in (("torso", 10), ("lwaist", 2.2619467), ("pelvis", 6.6161942), ("right_thigh", 4.751751),
                    ("right_shin", 4.522842), ("left_thigh", 4.751751), ("left_shin", 4.522842),
                    ("right_upper_arm", 1.6610805), ("right_lower_arm", 1.2295402), ("left_upper_arm", 1.6610805),
                    ("left_lower_arm", 1.2295402)):
        f[f"mass_{name}"] = _uf(f"mass_{name}", 1e-6, np.inf, m)
    return f


class CARLBraxHumanoid(CARLBraxEnv):
    """``carl/envs/brax/carl_humanoid.py:14-85``: 11 links on stacked 2- / 3-dof hinges, 17 actuators, the 244-entry
    observation of ``brax.envs.humanoid`` (q[2:], qd, cinert, cvel, actuator torques), centre-of-mass forward reward."""

    env_name: str = "humanoid"
    kind = "brax_humanoid"
    asset_path: str = "envs/assets/humanoid.xml"
    metadata = {"render_modes": []}

    @staticmethod
    def get_context_features() -> dict[str, ContextFeature]:
        f = _humanoid_features()
        f.update(_goal_features())
        return f


class CARLBraxHumanoidStandup(CARLBraxEnv):
    """``carl/envs/brax/carl_humanoidstandup.py:9-71``: the humanoid lying on its back; reward = torso height / dt + 1
    - 0.01 |a|^2, episodes end only by the time limit."""

    env_name: str = "humanoidstandup"
    kind = "brax_humanoidstandup"
    asset_path: str = "envs/assets/humanoidstandup.xml"
    metadata = {"render_modes": []}

    @staticmethod
    def get_context_features() -> dict[str, ContextFeature]:
        return _humanoid_features()


class CARLBraxPusher(CARLBraxEnv):
    """``carl/envs/brax/carl_pusher.py:11-103``: a 7-dof arm pushes a ball towards a goal marker; gripper-vs-ball
    contacts are body-vs-body pairs of the kernel's contact phase. Like the reference, the ``goal_position_*``
    features are part of the context (and of the observed context) but do not move the goal: the reference stores
    them in ``env._goal_pos``, an attribute nothing reads (:93-103; its own notebook prints the default goal for a
    changed context)."""

    env_name: str = "pusher"
    kind = "brax_pusher"
    asset_path: str = "envs/assets/pusher.xml"
    metadata = {"render_modes": []}
    GOAL_FEATURES = ("goal_position_x", "goal_position_y", "goal_position_z")

    @staticmethod
    def get_context_features() -> dict[str, ContextFeature]:
        f = _common_features()
        f["viscosity"] = _uf("viscosity", 0, np.inf, 0)
        for name, m in (("r_shoulder_pan_link", 7.2935214), ("r_shoulder_lift_link", np.pi), ("r_upper_arm_roll_link", 1.7140529),
                        ("r_elbow_flex_link", 4.0715042e-01), ("r_forearm_roll_link", 9.2818356e-01),
                        ("r_wrist_flex_link", 5.0265482e-03), ("r_wrist_roll_link", 1.8346901e-01), ("object", 1.8325957e-03)):
            f[f"mass_{name}"] = _uf(f"mass_{name}", 1e-6, np.inf, m)
        f["goal_position_x"] = _uf("goal_position_x", 0, np.inf, 0.45)
        f["goal_position_y"] = _uf("goal_position_y", 0, np.inf, 0.05)
        f["goal_position_z"] = _uf("goal_position_z", 0, np.inf, 0.05)
        return f

    @classmethod
    def kernel_params(cls, table, names, context_mode="reference", explicit=None):
        """``carl_pusher.py:93-103``: the goal features are taken out of the context before the family's
        ``_update_context`` runs (they never reach ``check_context`` nor the physics)."""
        keep = [j for j, n in enumerate(names) if n not in cls.GOAL_FEATURES]
        return super().kernel_params(np.asarray(table)[:, keep], [names[j] for j in keep], context_mode, explicit)


# Every body of carl/envs/brax/__init__.py is built.
UNSUPPORTED_BODIES = ()
