"""Batched contextual env base class.

``CARLEnv`` keeps the reference's gym-style surface (``carl/envs/carl_env.py:19-342``:
``contexts`` / ``context`` / ``context_id`` / ``context_selector`` / ``obs_context_features`` /
``obs_context_as_dict`` / ``observation_space`` / ``reset`` / ``step`` and the class-level
``get_context_features`` / ``get_context_space`` / ``get_default_context``) but steps
``num_envs`` independent context instances per call, gymnasium-``VectorEnv`` shaped like the
reference's only batched precedent (``carl/envs/brax/wrappers.py:93-158``): ``reset() ->
(obs, info)``, ``step(actions[N, ...]) -> (obs, reward[N], terminated[N], truncated[N], info)``.

The physics runs in hand-written CUDA (libcarlb, ``include/carlb.h``) on caller-owned device
buffers held here as torch tensors (device-buffer containers only). There is no CPU path.

Differences from the reference that a user must know (all documented in DESIGN.md):

* ``num_envs`` defaults to ``len(contexts)``; the context selector is *shared* by the batch and
  queried once per env instance being reset, in env-index order (so a first ``reset()`` with the
  default round-robin selector binds env i to context i).
* returned tensors are views of persistent device buffers: they are overwritten by the next call.
* numpy actions in -> numpy results out (pinned host buffers, copies inside the call);
  torch CUDA actions in -> torch CUDA results out (no host traffic, stream-ordered).
"""
from __future__ import annotations

import abc
import ctypes
import functools
import inspect
import os
from typing import Any

import numpy as np
import torch

from carl_b200 import _native, hostmem, spaces
from carl_b200.context.context_space import ContextFeature, ContextSpace
from carl_b200.context.selection import AbstractSelector, RoundRobinSelector
from carl_b200.utils.types import Context, Contexts

_TORCH_ACT = {
    torch.int32: _native.ACT_I32, torch.int64: _native.ACT_I64, torch.uint8: _native.ACT_U8,
    torch.float32: _native.ACT_F32,
}
_NP_ACT = {
    np.dtype(np.int32): _native.ACT_I32, np.dtype(np.int64): _native.ACT_I64, np.dtype(np.uint8): _native.ACT_U8,
    np.dtype(np.float32): _native.ACT_F32,
}
_UNSIGNED = {1: np.uint8, 4: np.uint32, 8: np.uint64}


# NVTX ranges around reset / step / rollout for nsys / ncu timelines (SURVEY §5 "tracing"): CARLB_NVTX=1.
# Off by default: a range push / pop pair costs ~1 us of host time per call.
_NVTX = os.environ.get("CARLB_NVTX", "0") == "1"


def _nvtx(name):
    def deco(fn):
        if not _NVTX:
            return fn

        @functools.wraps(fn)
        def wrapped(self, *a, **k):
            torch.cuda.nvtx.range_push(f"carlb.{self.kind}.{name}")
            try:
                return fn(self, *a, **k)
            finally:
                torch.cuda.nvtx.range_pop()
        return wrapped
    return deco


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


class ContextTable:
    """Dense context set: ``float64[M, F]`` + feature names (+ the original keys).

    The reference's ``Contexts`` is a dict of dicts (``carl/utils/types.py:5-6``); building 65 536
    Python dicts is seconds of host work, so samplers can hand the batch env a table directly
    (``ContextSampler.sample_context_table``). ``to_contexts()`` gives the dict view back."""

    def __init__(self, names: list[str], values: np.ndarray, keys: list[Any] | None = None):
        self.names = list(names)
        self.values = np.ascontiguousarray(values, dtype=np.float64)
        assert self.values.ndim == 2 and self.values.shape[1] == len(self.names)
        self.keys = list(range(self.values.shape[0])) if keys is None else list(keys)
        self._index = None

    def index_of(self, key) -> int:
        if self._index is None:
            self._index = {k: i for i, k in enumerate(self.keys)}
        return self._index[key]

    def __len__(self) -> int:
        return self.values.shape[0]

    def to_contexts(self) -> Contexts:
        return {k: {n: float(v) for n, v in zip(self.names, row)} for k, row in zip(self.keys, self.values)}

    def keys_view(self):
        return self.keys


class CARLEnv(abc.ABC):
    """Batched drop-in for ``carl.envs.carl_env.CARLEnv`` (reference ``carl_env.py:19``)."""

    kind: str  # libcarlb env kind name (see _native.KIND)
    metadata: dict = {"render_modes": []}
    # Legacy names of CARL's v0 docs (docs/source/environments/data/context_definitions/*.csv) that
    # the reference's code no longer has: aliases are renamed on entry, extension
    # features are accepted in a context (with a default) without being part of the context space.
    feature_aliases: dict = {}
    extension_features: dict = {}

    def __init__(
        self,
        env: Any = None,
        contexts: Contexts | ContextTable | None = None,
        obs_context_features: list[str] | None = None,
        obs_context_as_dict: bool = True,
        context_selector: AbstractSelector | type[AbstractSelector] | None = None,
        context_selector_kwargs: dict | None = None,
        *,
        num_envs: int | None = None,
        device: str | torch.device | int = "cuda",
        dtype: str = "float32",
        context_mode: str = "reference",
        autoreset: bool | None = None,
        max_episode_steps: int | None = None,
        shard: tuple[int, int] | None = None,
        validate_actions: bool = True,
        **kwargs,
    ):
        """Parameters follow ``carl_env.py:20-74``; the keyword-only ones are new.

        num_envs: env instances in the (global) batch; default ``len(contexts)``.
        device: CUDA device of this shard.
        dtype: ``"float32"`` (throughput) or ``"float64"`` (reference precision, classic control).
        context_mode: ``"reference"`` reproduces what the reference's context injection actually
            does, ``"applied"`` what it intends.
        autoreset: same-step auto-reset inside the step kernel (default: off for classic control
            as in the reference, on for Brax where ``brax.envs.create`` adds AutoResetWrapper).
        shard: ``(rank, world_size)`` -- this object owns the contiguous slice of the global batch
            that ``carl_b200.parallel.shard_range`` assigns to ``rank``.
        validate_actions: host-buffer path only -- assert that discrete actions lie in the action
            space, as gymnasium's envs do on every step; device tensors are never validated (no sync).
        """
        if env is not None:
            raise ValueError("carl_b200 envs own their physics; passing a gymnasium/brax `env` is not supported.")
        if kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")
        self._lib = _native.load()  # fails loudly when libcarlb is missing
        self.obs_context_as_dict = obs_context_as_dict
        if dtype not in ("float32", "float64"):
            raise ValueError(f"dtype must be 'float32' or 'float64', got {dtype!r}")
        if context_mode not in ("reference", "applied"):
            raise ValueError(f"context_mode must be 'reference' or 'applied', got {context_mode!r}")
        self.dtype = dtype
        self.context_mode = context_mode
        self._info = _native.query_env(_native.KIND[self.kind])

        if contexts is None:
            contexts = {0: self.get_default_context()}
        self.contexts = contexts  # setter fills defaults and builds the dense table
        self.context: Context | None = None
        if obs_context_features is None:
            # the reference takes the keys of the first (default-filled) context (carl_env.py:86-88)
            obs_context_features = [n for n in self._feature_names if n not in self.extension_features]
        self.obs_context_features = obs_context_features

        # Context selector (carl_env.py:90-108)
        sel_contexts = _LazyContexts(self)
        if context_selector is None:
            self.context_selector = RoundRobinSelector(contexts=sel_contexts)
        elif isinstance(context_selector, AbstractSelector):
            self.context_selector = context_selector
        elif inspect.isclass(context_selector) and issubclass(context_selector, AbstractSelector):
            if context_selector_kwargs is None:
                context_selector_kwargs = {}
            _context_selector_kwargs = {"contexts": sel_contexts}
            context_selector_kwargs.update(_context_selector_kwargs)
            self.context_selector = context_selector(**context_selector_kwargs)
        else:
            raise ValueError(
                f"Context selector must be None or an AbstractSelector class or instance. "
                f"Got type {type(context_selector)}."
            )

        # batch geometry
        self.global_num_envs = int(num_envs) if num_envs is not None else len(self._table)
        if self.global_num_envs <= 0:
            raise ValueError("num_envs must be positive")
        if shard is None:
            self.rank, self.world_size = 0, 1
        else:
            self.rank, self.world_size = int(shard[0]), int(shard[1])
        from carl_b200.parallel import shard_range

        self.env_lo, self.env_hi = shard_range(self.global_num_envs, self.rank, self.world_size)
        self.num_envs = self.env_hi - self.env_lo
        if self.num_envs <= 0:
            raise ValueError(f"rank {self.rank} of {self.world_size} owns no env instances")

        self.device = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if self.device.type != "cuda":
            raise ValueError("carl_b200 runs on CUDA devices only (no CPU fallback)")
        if not torch.cuda.is_available():
            raise RuntimeError("carl_b200 needs a CUDA device: torch.cuda.is_available() is False")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())

        # spaces (carl_env.py:77,110-112)
        self.base_observation_space = self._single_observation_space()
        self.single_action_space = self._single_action_space()
        self.action_space = self._batched_action_space()
        self.observation_space = self.get_observation_space(obs_context_feature_names=self.obs_context_features)

        self._autoreset = self._default_autoreset() if autoreset is None else bool(autoreset)
        self._max_episode_steps = int(max_episode_steps) if max_episode_steps else self._info.default_max_steps
        self._alloc()
        self._context_ids = np.full(self.num_envs, -1, dtype=np.int64)  # per-env current context id
        self._reset_counts = np.zeros(self.num_envs, dtype=np.int64)
        self._rr_offset = np.zeros(self.num_envs, dtype=np.int64)
        self._ctx_obs_cache = None
        self._ctx_obs_host_cache = None
        self._ids_view = None
        self._seeded = False
        self._seed_value = None
        self._has_reset = False
        self._validate_actions = bool(validate_actions)
        self._host_io = None
        self._async_parts = min(2, self.num_envs)
        self._async_pending = [False] * 8
        self._async_cache = None

    # ------------------------------------------------------------------ contexts
    @property
    def contexts(self) -> Contexts:
        if self._contexts_dict is None:
            self._contexts_dict = self._table.to_contexts()
        return self._contexts_dict

    @contexts.setter
    def contexts(self, contexts: Contexts | ContextTable) -> None:
        """``carl_env.py:122-137``: every context is filled with the default values."""
        context_space = self.get_context_space()
        space_defaults = context_space.get_default_context()
        defaults = dict(space_defaults)
        # features registered under a key that differs from their NAME (the reference's `mass_pole2`,
        # carl_inverted_double_pendulum.py:33-35) are missing from the name-keyed defaults but still settable
        for key, feat in self.get_context_features().items():
            if key not in defaults:
                defaults[key] = feat.default_value
        defaults.update(self.extension_features)
        names = list(defaults.keys())
        self._feature_names = names
        alias = self.feature_aliases
        if isinstance(contexts, ContextTable):
            cnames = [alias.get(n, n) for n in contexts.names]
            unknown = [n for n in cnames if n not in defaults]
            if unknown:
                raise ValueError(f"Unknown context features {unknown}")
            vals = np.empty((len(contexts), len(names)), dtype=np.float64)
            for j, n in enumerate(names):
                vals[:, j] = contexts.values[:, cnames.index(n)] if n in cnames else float(defaults[n])
            self._table = ContextTable(names, vals, contexts.keys)
            self._contexts_dict = None
            # which features each context named itself (kernel_params of "applied" mode may need to tell a
            # deliberately chosen default value from a filled-in one)
            self._explicit = np.tile(np.array([n in cnames for n in names], dtype=bool), (len(contexts), 1))
        else:
            renamed = {k: {alias.get(n, n): v for n, v in c.items()} for k, c in contexts.items()}
            filled = {}
            for k, c in renamed.items():
                f = context_space.insert_defaults({n: v for n, v in c.items() if n in space_defaults or n not in defaults})
                for n in defaults:
                    if n not in space_defaults and n not in self.extension_features:
                        f[n] = c.get(n, defaults[n])
                for n, d in self.extension_features.items():
                    if n in c:
                        f[n] = c[n]
                filled[k] = f
            for k, c in filled.items():
                unknown = [n for n in c if n not in defaults]
                if unknown:
                    raise ValueError(f"Unknown context features {unknown} in context {k!r}")
            vals = np.array([[float(c.get(n, defaults[n])) for n in names] for c in filled.values()], dtype=np.float64)
            self._table = ContextTable(names, vals.reshape(len(filled), len(names)), list(filled.keys()))
            self._contexts_dict = filled
            self._explicit = np.array([[n in c for n in names] for c in renamed.values()], dtype=bool).reshape(len(filled), len(names))
        self._params_table = None  # rebuilt lazily by _update_context

    @property
    def context_table(self) -> ContextTable:
        return self._table

    @property
    def context_id(self):
        """Current context id: an int for a single env instance, else ``int64[num_envs]``."""
        view = getattr(self, "_ids_view", None)  # per-step fast path: dropped whenever the ids change
        if view is None:
            if getattr(self, "_context_ids", None) is None or (self._context_ids < 0).all():
                return self.context_selector.context_id
            view = self._ids_view = self._context_ids.copy()
            view.setflags(write=False)
        return int(view[0]) if self.num_envs == 1 else view

    @context_id.setter
    def context_id(self, new_id) -> None:
        """``carl_env.py:139-157``: switch context immediately (an int applies to every env)."""
        ids = np.broadcast_to(np.asarray(new_id, dtype=np.int64), (self.num_envs,)).copy()
        valid = np.isin(ids, np.asarray(self.context_selector.context_ids))
        assert valid.all(), "Unknown ID, this context does not exist in the context set."
        self.context_selector.context_id = int(ids[-1])
        self.context_selector.context = self.context_selector.contexts[
            self.context_selector.contexts_keys[int(ids[-1])]
        ]
        self._context_ids = ids
        # a round-robin selector continues from the assigned id (selection.py:116-118)
        nxt = np.arange(self.num_envs) + self.env_lo + self._reset_counts * self.global_num_envs
        self._rr_offset = (ids + 1 - nxt) % len(self._table)
        self._refresh_context_view()
        self._update_context()

    def _refresh_context_view(self) -> None:
        ids = self._context_ids
        vals = self._table.values[ids]
        if self.num_envs == 1:
            self.context = {n: _pyval(v) for n, v in zip(self._feature_names, vals[0])}
        else:
            self.context = {n: vals[:, j].copy() for j, n in enumerate(self._feature_names)}
        self._ctx_obs_cache = None
        self._ctx_obs_host_cache = None
        self._ids_view = None
        if getattr(self, "_async_cache", None) is not None:  # per-part context views of the split-batch API
            self._async_cache["ctx"] = [None] * len(self._async_cache["ctx"])
            self._async_cache["ids"] = [None] * len(self._async_cache["ctx"])

    # -------------------------------------------------------------------- spaces
    def get_observation_space(self, obs_context_feature_names: list[str] | None = None):
        """``carl_env.py:159-188``: ``Dict{"obs": base, "context": ...}`` (single-env spaces)."""
        context_space = self.get_context_space()
        obs_space_context = context_space.to_gymnasium_space(
            context_feature_names=obs_context_feature_names, as_dict=self.obs_context_as_dict
        )
        return spaces.Dict({"obs": self.base_observation_space, "context": obs_space_context})

    def _single_observation_space(self):
        hi = np.full(self._info.obs_dim, np.inf, dtype=np.float32)
        return spaces.Box(-hi, hi, dtype=np.float32)

    def _single_action_space(self):
        if self._info.act_discrete:
            return spaces.Discrete(self._info.n_actions)
        lo = np.full(self._info.act_dim, self._info.act_low, dtype=np.float32)
        hi = np.full(self._info.act_dim, self._info.act_high, dtype=np.float32)
        return spaces.Box(lo, hi, dtype=np.float32)

    def _batched_action_space(self):
        if self._info.act_discrete:
            return spaces.Box(
                low=np.zeros(self.num_envs), high=np.full(self.num_envs, self._info.n_actions - 1), dtype=np.int64
            )
        return spaces.batch_box(self.single_action_space, self.num_envs)

    @staticmethod
    @abc.abstractmethod
    def get_context_features() -> dict[str, ContextFeature]:
        """``carl_env.py:190-202``."""
        ...

    @classmethod
    def get_context_space(cls) -> ContextSpace:
        """``carl_env.py:204-214``."""
        return ContextSpace(cls.get_context_features())

    @classmethod
    def get_default_context(cls) -> Context:
        """``carl_env.py:216-226``."""
        return cls.get_context_space().get_default_context()

    def _default_autoreset(self) -> bool:
        return False

    # ------------------------------------------------------------------- buffers
    def _alloc(self) -> None:
        n, info, dev = self.num_envs, self._info, self.device
        tdt = torch.float32 if self.dtype == "float32" else torch.float64
        self._precision = _native.F32 if self.dtype == "float32" else _native.F64
        is_brax = self.kind.startswith("brax")
        z = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype, device=dev)
        self._state = z(n, info.state_words, dtype=(torch.float32 if is_brax else tdt))
        # classic: SoA rows ctx[P][n] (one thread per env); Brax: AoS ctx[n][P] (one warp per env reads
        # its P scalars with one coalesced load and broadcasts them with shuffles)
        ctx_shape = (n, info.n_param_rows) if is_brax else (info.n_param_rows, n)
        self._ctx = z(*ctx_shape, dtype=(torch.float32 if is_brax else tdt))
        self._ctx_aos = is_brax
        self._elapsed = z(n, dtype=torch.int32)
        self._sbt = z(n, dtype=torch.uint8)
        self._rng = z(4, n, dtype=torch.int64)
        # obs | reward | terminated | truncated back to back: the host path moves them with ONE copy
        ob, rb = n * info.obs_dim * 4, n * 4
        self._out = z(ob + rb + 2 * n, dtype=torch.uint8)
        self._obs = self._out[:ob].view(torch.float32).view(n, info.obs_dim)
        self._reward = self._out[ob:ob + rb].view(torch.float32)
        self._terminated = self._out[ob + rb:ob + rb + n]
        self._truncated = self._out[ob + rb + n:ob + rb + 2 * n]
        self._terminated_b = self._terminated.view(torch.bool)  # cached views: step() returns these
        self._truncated_b = self._truncated.view(torch.bool)
        self._final_obs = z(n, info.obs_dim, dtype=torch.float32)
        self._first_state = z(n, info.state_words, dtype=torch.float32) if is_brax else None
        self._first_obs = z(n, info.obs_dim, dtype=torch.float32) if is_brax else None
        self._act_staging = z(n * max(1, info.act_dim), dtype=torch.int64)
        self._handle = ctypes.c_void_p()
        _native.check(self._lib.carlb_env_create(
            _native.KIND[self.kind], n, self._precision, self.device.index, self.env_lo, ctypes.byref(self._handle)))
        b = _native.Buffers(
            state=_ptr(self._state), ctx=_ptr(self._ctx), elapsed=_ptr(self._elapsed), sbt=_ptr(self._sbt),
            rng=_ptr(self._rng), obs=_ptr(self._obs), reward=_ptr(self._reward), terminated=_ptr(self._terminated),
            truncated=_ptr(self._truncated), final_obs=_ptr(self._final_obs), first_state=_ptr(self._first_state),
            first_obs=_ptr(self._first_obs), act_staging=_ptr(self._act_staging),
        )
        _native.check(self._lib.carlb_env_bind(self._handle, ctypes.byref(b)))
        _native.check(self._lib.carlb_env_configure(
            self._handle, self._max_episode_steps,
            _native.AUTORESET_SAME_STEP if self._autoreset else _native.AUTORESET_NONE))
        self._post_alloc()

    def _post_alloc(self) -> None:
        """Family hook run once the native handle exists (Brax uploads its system table)."""

    def close(self) -> None:
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            self._lib.carlb_env_destroy(h)
            self._handle = ctypes.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # ------------------------------------------------------- context -> kernel
    @classmethod
    @abc.abstractmethod
    def kernel_params(cls, table: np.ndarray, names: list[str], context_mode: str = "reference",
                      explicit: np.ndarray | None = None) -> np.ndarray:
        """Map the dense context table ``float64[M, F]`` onto the kernel-parameter columns
        ``float64[M, P]`` (the batched form of ``_update_context``, ``carl_env.py:307-319``).
        Pure host logic (testable without a GPU)."""
        ...

    def _update_context(self, mask: np.ndarray | None = None) -> None:
        """Upload the kernel-parameter rows of the envs whose context id changed."""
        if self._params_table is None:
            self._params_table = np.ascontiguousarray(
                self.kernel_params(self._table.values, self._feature_names, self.context_mode, explicit=self._explicit))
            assert self._params_table.shape == (len(self._table), self._info.n_param_rows)
        ids = self._context_ids
        np_dt = np.float32 if self._ctx.dtype == torch.float32 else np.float64
        if mask is None or mask.all():
            rows = self._params_table[ids].astype(np_dt)
            rows = np.ascontiguousarray(rows if self._ctx_aos else rows.T)
            self._ctx.copy_(torch.from_numpy(rows), non_blocking=False)
        else:
            idx = np.nonzero(mask)[0]
            rows = self._params_table[ids[idx]].astype(np_dt)
            idx_t = torch.from_numpy(idx).to(self.device)
            if self._ctx_aos:
                self._ctx[idx_t] = torch.from_numpy(np.ascontiguousarray(rows)).to(self.device)
            else:
                self._ctx[:, idx_t] = torch.from_numpy(np.ascontiguousarray(rows.T)).to(self.device)

    def _progress_instance(self, mask: np.ndarray | None = None) -> np.ndarray:
        """``carl_env.py:228-243`` batched: one ``select()`` per env being reset, env order."""
        n_sel = self.num_envs if mask is None else int(mask.sum())
        sel = self.context_selector
        if type(sel) is RoundRobinSelector:
            # Round robin per env instance: the k-th reset of (global) env i selects context
            # (i + k * N) mod M. For lock-step full resets this is exactly what one shared
            # RoundRobinSelector queried in env order returns; unlike it, a partial (masked) reset
            # does not depend on which other envs happen to reset (N contexts <-> N envs stay bound).
            local = np.arange(self.num_envs) if mask is None else np.nonzero(mask)[0]
            k = self._reset_counts[local]
            new_ids = ((local + self.env_lo) + k * self.global_num_envs + self._rr_offset[local]) % len(self._table)
            self._reset_counts[local] += 1
            sel.n_calls += self.global_num_envs if mask is None else n_sel
            if len(new_ids):
                sel.context_id = int(new_ids[-1])
        # envs owned by other shards consume selector calls too, so ids do not depend on sharding
        elif mask is None and self.world_size > 1:
            all_ids = sel.select_batch(self.global_num_envs)
            new_ids = all_ids[self.env_lo:self.env_hi]
        else:
            new_ids = sel.select_batch(n_sel)
        changed = np.zeros(self.num_envs, dtype=bool)
        if mask is None:
            changed = new_ids != self._context_ids
            self._context_ids = np.asarray(new_ids, dtype=np.int64).copy()
        else:
            idx = np.nonzero(mask)[0]
            changed[idx] = new_ids != self._context_ids[idx]
            self._context_ids[idx] = new_ids
        self._refresh_context_view()
        return changed

    # ------------------------------------------------------------- reset / step
    @_nvtx("reset")
    def reset(self, *, seed: int | None = None, options: dict[str, Any] | None = None,
              mask: np.ndarray | torch.Tensor | None = None):
        """``carl_env.py:245-274``: select contexts, re-inject those that changed, reset, dict obs.

        seed: env i of the *global* batch is seeded with ``seed + i`` (gymnasium vector-env
        convention), reproducing ``np.random.Generator(PCG64(SeedSequence(seed + i)))`` on device.
        mask: optional ``bool[num_envs]`` -- reset only those env instances."""
        mask_np = None
        if mask is not None:
            mask_np = (mask.detach().cpu().numpy() if isinstance(mask, torch.Tensor) else np.asarray(mask)).astype(bool)
            assert mask_np.shape == (self.num_envs,)
        changed = self._progress_instance(mask_np)
        if changed.any():
            self._update_context(changed if not changed.all() else None)
        st = self._stream()
        if seed is not None:
            _native.check(self._lib.carlb_env_seed(self._handle, int(seed), st))
            self._seeded = True
            self._seed_value = int(seed)
        elif not self._seeded:
            # gymnasium seeds from OS entropy on the first unseeded reset
            entropy = int(np.random.SeedSequence().generate_state(1, np.uint64)[0] >> 1)
            _native.check(self._lib.carlb_env_seed(self._handle, entropy, st))
            self._seeded = True
            self._seed_value = entropy
        mask_t = None
        if mask_np is not None:
            mask_t = torch.from_numpy(mask_np.astype(np.uint8)).to(self.device)
        _native.check(self._lib.carlb_env_reset(self._handle, _ptr(mask_t), st))
        self._has_reset = True
        state = self._add_context_to_state(self._obs)
        info = {"context_id": self.context_id}
        return state, info

    def _context_obs_tensors(self):
        if self._ctx_obs_cache is None:
            ids = self._context_ids
            cols = [self._feature_names.index(k) for k in self.obs_context_features]
            vals = self._table.values[ids][:, cols].astype(np.float32) if cols else np.zeros((self.num_envs, 0), np.float32)
            t = torch.from_numpy(np.ascontiguousarray(vals)).to(self.device)
            if self.obs_context_as_dict:
                self._ctx_obs_cache = {k: t[:, j] for j, k in enumerate(self.obs_context_features)}
            else:
                self._ctx_obs_cache = t
        return self._ctx_obs_cache

    def _add_context_to_state(self, state: Any) -> dict[str, Any]:
        """``carl_env.py:276-305``: ``{"obs": state, "context": dict | vector}`` (batched columns)."""
        return {"obs": state, "context": self._context_obs_tensors()}

    def _check_actions(self, n_expected: int, shape: tuple) -> None:
        a = self._info.act_dim
        ok = shape in ((n_expected,), (n_expected, a)) if a == 1 else shape == (n_expected, a)
        assert ok, f"actions must have shape ({n_expected}, {a}), got {shape}"

    @_nvtx("step")
    def step(self, action: Any):
        """``carl_env.py:321-342`` batched. torch CUDA actions -> device results;
        numpy / list actions -> host (numpy) results through pinned buffers."""
        if not self._has_reset:  # gymnasium's OrderEnforcing wrapper raises ResetNeeded here
            raise RuntimeError("Cannot call env.step() before calling env.reset()")
        if isinstance(action, torch.Tensor) and action.is_cuda:
            self._check_actions(self.num_envs, tuple(action.shape))
            if action.dtype not in _TORCH_ACT or (self._info.act_discrete and action.dtype.is_floating_point):
                # as the host path does: a float tensor of integral action values is accepted for discrete envs
                action = action.to(torch.int32 if self._info.act_discrete else torch.float32)
            if not self._info.act_discrete and action.dtype != torch.float32:
                action = action.to(torch.float32)
            action = action.contiguous()
            _native.check(self._lib.carlb_env_step(self._handle, action.data_ptr(), _TORCH_ACT[action.dtype], self._stream()))
            obs, rew, term, trunc = self._obs, self._reward, self._terminated_b, self._truncated_b
            state = self._add_context_to_state(obs)
            info = {"context_id": self.context_id}
            if self._autoreset:
                info["final_observation"] = self._final_obs
            return state, rew, term, trunc, info
        return self._step_host(action)

    def _ensure_host_io(self):
        if self._host_io is None:
            n, info = self.num_envs, self._info
            pin = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype).pin_memory()
            ob, rb = n * info.obs_dim * 4, n * 4
            out = pin(ob + rb + 2 * n, dtype=torch.uint8)  # same packing as the device side
            act = pin(n * max(1, info.act_dim), dtype=torch.int64)
            obs, reward = out[:ob].view(torch.float32).view(n, info.obs_dim), out[ob:ob + rb].view(torch.float32)
            term, trunc = out[ob + rb:ob + rb + n], out[ob + rb + n:ob + rb + 2 * n]
            act_bytes = act.numpy().view(np.uint8)
            self._host_io = dict(
                act=act, out=out, obs=obs, reward=reward, term=term, trunc=trunc,
                # everything the per-step path touches is resolved once: raw pointers, numpy views of the
                # pinned result block, one staging view per accepted action dtype
                ptrs=(act.data_ptr(), obs.data_ptr(), reward.data_ptr(), term.data_ptr(), trunc.data_ptr()),
                np_obs=obs.numpy(), np_reward=reward.numpy(), np_term=term.numpy().view(np.bool_),
                np_trunc=trunc.numpy().view(np.bool_),
                staged={dt: act_bytes[:n * max(1, info.act_dim) * np.dtype(dt).itemsize].view(dt) for dt in _NP_ACT},
            )
        return self._host_io

    def _step_host(self, action: Any):
        io = self._ensure_host_io()
        a = np.asarray(action.cpu() if isinstance(action, torch.Tensor) else action)
        self._check_actions(self.num_envs, tuple(a.shape))
        if self._info.act_discrete:
            if a.dtype not in _NP_ACT or a.dtype == np.float32:
                a = a.astype(np.int64)
        elif a.dtype != np.float32:
            a = a.astype(np.float32)
        flat = np.ascontiguousarray(a).reshape(-1)
        n_act = self._info.n_actions if (self._validate_actions and self._info.act_discrete) else 0
        p = io["ptrs"]
        src = flat.ctypes.data
        # an array from carl_b200.hostmem.pinned_empty is read in place by the step kernel (range check
        # only); anything else takes one native pass: copy into the page-locked staging block + check
        if hostmem.is_pinned(src, flat.nbytes):
            # one call: the step kernel validates the actions while it reads them and rolls the step back
            # if one is out of range (carlb_env_step_host_checked) -- no host pass over the array
            rc = self._lib.carlb_env_step_host_checked(self._handle, src, _NP_ACT[a.dtype], n_act, p[1], p[2], p[3], p[4],
                                                       self._stream())
            if rc != 0:
                msg = _native.last_error()
                if msg.startswith("invalid action"):
                    raise AssertionError(msg)  # gymnasium: `assert self.action_space.contains(action)`
                _native.check(rc)
        else:
            if self._lib.carlb_stage_actions(p[0], src, flat.size, _NP_ACT[a.dtype], n_act) != 0:
                raise AssertionError(_native.last_error())
            _native.check(self._lib.carlb_env_step_host(self._handle, p[0], _NP_ACT[a.dtype], p[1], p[2], p[3], p[4],
                                                        self._stream()))
        state = {"obs": io["np_obs"], "context": self._context_obs_host()}
        return state, io["np_reward"], io["np_term"], io["np_trunc"], {"context_id": self.context_id}

    # ------------------------------------------------- split-batch (EnvPool-style) host stepping
    @property
    def async_parts(self) -> int:
        """Number of contiguous parts ``step_async`` / ``step_wait`` cut the batch into (default 2)."""
        return self._async_parts

    @async_parts.setter
    def async_parts(self, k: int) -> None:
        if any(self._async_pending):
            raise RuntimeError("async_parts cannot change while a step is in flight")
        k = int(k)
        if not (1 <= k <= 8 and k <= self.num_envs):
            raise ValueError("async_parts must lie in [1, min(8, num_envs)]")
        self._async_parts = k
        self._async_cache = None

    def part_range(self, part: int) -> tuple[int, int]:
        """``[lo, hi)`` of part ``part`` among this object's env instances."""
        k = self._async_parts
        return self.num_envs * part // k, self.num_envs * (part + 1) // k

    def _async_state(self):
        if self._async_cache is None:
            io = self._ensure_host_io()
            k = self._async_parts
            bounds = [self.part_range(p) for p in range(k)]
            streams = [torch.cuda.Stream(self.device) for _ in range(k)]
            p = io["ptrs"]
            self._async_cache = dict(
                streams=streams, bounds=bounds,
                # everything the per-call path needs, resolved once
                begin=self._lib.carlb_env_step_host_begin, end=self._lib.carlb_env_step_host_end,
                tail=[(p[1], p[2], p[3], p[4], s_.cuda_stream) for s_ in streams],
                n_act=self._info.n_actions if (self._validate_actions and self._info.act_discrete) else 0,
                views=[(io["np_obs"][lo:hi], io["np_reward"][lo:hi], io["np_term"][lo:hi], io["np_trunc"][lo:hi])
                       for lo, hi in bounds],
                ctx=[None] * k, ids=[None] * k)
        return self._async_cache

    def step_async(self, action: Any, part: int | None = None) -> None:
        """Split-batch stepping with host buffers (EnvPool ``send`` / SB3 ``VecEnv.step_async``): enqueue the step
        of one part (``action`` = that part's actions) or, with ``part=None``, of every part (``action`` = the whole
        batch) and return at once. ``step_wait(part)`` hands back the part's results; while the host consumes them
        and computes the part's next actions, the other parts' results are crossing PCIe. Classic-control envs."""
        if not self._has_reset:
            raise RuntimeError("Cannot call env.step_async() before calling env.reset()")
        st = self._async_state()
        io = self._host_io
        if part is None:
            a = np.asarray(action)
            self._check_actions(self.num_envs, tuple(a.shape))
            flat = a.reshape(self.num_envs, -1)
            for p_, (lo, hi) in enumerate(st["bounds"]):
                self.step_async(flat[lo:hi].reshape((hi - lo,) + tuple(a.shape[1:])), part=p_)
            return
        lo, hi = st["bounds"][part]
        a = action if type(action) is np.ndarray else np.asarray(action)
        adim = self._info.act_dim
        if not (a.shape == (hi - lo,) and adim == 1) and a.shape != (hi - lo, adim):
            self._check_actions(hi - lo, tuple(a.shape))
        if self._info.act_discrete:
            if a.dtype not in _NP_ACT or a.dtype == np.float32:
                a = a.astype(np.int64)
        elif a.dtype != np.float32:
            a = a.astype(np.float32)
        if not a.flags.c_contiguous:
            a = np.ascontiguousarray(a)
        n_act = st["n_act"]
        src = a.ctypes.data
        if not hostmem.is_pinned(src, a.nbytes):
            # pageable actions: one native pass copies them into the part's slice of the page-locked staging block
            # (and range-checks them; the kernel checks again, which costs nothing)
            adim = max(1, adim)
            dst = io["staged"][a.dtype][lo * adim:hi * adim]
            if self._lib.carlb_stage_actions(dst.ctypes.data, src, a.size, _NP_ACT[a.dtype], n_act) != 0:
                raise AssertionError(_native.last_error())
            src = dst.ctypes.data
        rc = st["begin"](self._handle, part, self._async_parts, src, _NP_ACT[a.dtype], n_act, *st["tail"][part])
        if rc != 0:
            _native.check(rc)
        self._async_pending[part] = True

    def step_wait(self, part: int | None = None):
        """Results of the step ``step_async`` enqueued for ``part`` (views of the page-locked result arrays
        restricted to the part's envs) -- or, with ``part=None``, of the whole batch once every part has landed."""
        st = self._async_state()
        parts = range(self._async_parts) if part is None else (part,)
        for p_ in parts:
            if not self._async_pending[p_]:
                raise RuntimeError(f"step_wait: part {p_} has no step in flight")
            rc = st["end"](self._handle, p_)
            self._async_pending[p_] = False
            if rc != 0:
                msg = _native.last_error()
                if msg.startswith("invalid action"):
                    raise AssertionError(msg)
                _native.check(rc)
        io = self._host_io
        if part is None:
            state = {"obs": io["np_obs"], "context": self._context_obs_host()}
            return state, io["np_reward"], io["np_term"], io["np_trunc"], {"context_id": self.context_id}
        lo, hi = st["bounds"][part]
        if st["ctx"][part] is None:
            full = self._context_obs_host()
            if isinstance(full, dict):
                st["ctx"][part] = full if self.num_envs == 1 else {k: v[lo:hi] for k, v in full.items()}
            else:
                st["ctx"][part] = full[lo:hi]
        if st["ids"][part] is None:
            ids = self.context_id
            st["ids"][part] = ids if np.isscalar(ids) else ids[lo:hi]
        obs, rew, term, trunc = st["views"][part]
        return {"obs": obs, "context": st["ctx"][part]}, rew, term, trunc, {"context_id": st["ids"][part]}

    def _context_obs_host(self):
        if self._ctx_obs_host_cache is None:  # contexts only change at reset / context_id assignment
            ids = self._context_ids
            cols = [self._feature_names.index(k) for k in self.obs_context_features]
            vals = self._table.values[ids][:, cols]
            if self.obs_context_as_dict:
                if self.num_envs == 1:
                    self._ctx_obs_host_cache = {k: _pyval(vals[0, j]) for j, k in enumerate(self.obs_context_features)}
                else:
                    self._ctx_obs_host_cache = {k: np.ascontiguousarray(vals[:, j]) for j, k in enumerate(self.obs_context_features)}
            else:
                self._ctx_obs_host_cache = vals.astype(np.float32)
        return self._ctx_obs_host_cache

    # ----------------------------------------------------------- fused rollout
    @_nvtx("rollout")
    def rollout(self, n_steps: int, policy_seed: int = 0, step_base: int = 0, actions: torch.Tensor | None = None,
                record: bool = False):
        """Advance every env ``n_steps`` steps in ONE launch (state stays in registers).

        actions=None: synthetic uniform random policy from Philox4x32-10 keyed by
        ``(policy_seed, global env id, step_base + t)``. record=True returns the trajectory
        ``{"obs": [K,N,D], "actions": [K,N(,A)], "reward": [K,N], "done": [K,N] (bit0 term, bit1 trunc)}``."""
        if not self._has_reset:
            raise RuntimeError("Cannot call env.rollout() before calling env.reset()")
        n, info, dev = self.num_envs, self._info, self.device
        traj_t = None
        traj = None
        if record:
            adt = torch.int32 if info.act_discrete else torch.float32
            ashape = (n_steps, n) if info.act_dim == 1 else (n_steps, n, info.act_dim)
            traj_t = dict(
                obs=torch.empty(n_steps, n, info.obs_dim, dtype=torch.float32, device=dev),
                actions=torch.empty(*ashape, dtype=adt, device=dev),
                reward=torch.empty(n_steps, n, dtype=torch.float32, device=dev),
                done=torch.empty(n_steps, n, dtype=torch.uint8, device=dev),
            )
            traj = _native.Traj(obs=_ptr(traj_t["obs"]), actions=_ptr(traj_t["actions"]),
                                reward=_ptr(traj_t["reward"]), done=_ptr(traj_t["done"]))
        act_ptr, act_dt = None, _native.ACT_I32
        if actions is not None:
            assert actions.is_cuda and actions.shape[0] == n_steps and actions.shape[1] == n
            actions = actions.contiguous()
            act_ptr, act_dt = actions.data_ptr(), _TORCH_ACT[actions.dtype]
        _native.check(self._lib.carlb_env_rollout(
            self._handle, int(n_steps), int(policy_seed), int(step_base), act_ptr, act_dt,
            ctypes.byref(traj) if traj is not None else None, self._stream()))
        return traj_t

    # ------------------------------------------------------------ checkpointing
    def state_dict(self) -> dict[str, Any]:
        """Everything needed to resume an uninterrupted run: env state, counters, RNG streams (classic: the
        PCG64 words; Brax: the reset seed + per-env episode counters), the context binding AND the selection
        state (per-env round-robin counters, the selector's ``context_id`` / ``n_calls``), the goal wrapper's
        dead-reckoned positions. ``load_state_dict`` checks kind, batch geometry and precision."""
        sel = self.context_selector
        goal = None
        if getattr(self, "_goal_state", None) is not None:
            goal = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in self._goal_state.items()}
        return {
            "meta": {"kind": self.kind, "num_envs": self.num_envs, "global_num_envs": self.global_num_envs,
                     "env_lo": self.env_lo, "dtype": self.dtype, "n_contexts": len(self._table)},
            "state": self._state.clone(), "elapsed": self._elapsed.clone(), "sbt": self._sbt.clone(),
            "rng": self._rng.clone(), "obs": self._obs.clone(), "reward": self._reward.clone(),
            "terminated": self._terminated.clone(), "truncated": self._truncated.clone(),
            "context_ids": self._context_ids.copy(), "reset_counts": self._reset_counts.copy(),
            "rr_offset": self._rr_offset.copy(),
            "selector": {"context_id": getattr(sel, "context_id", None), "n_calls": getattr(sel, "n_calls", None)},
            "seed": self._seed_value, "seeded": self._seeded, "has_reset": self._has_reset,
            "first_state": None if self._first_state is None else self._first_state.clone(),
            "first_obs": None if self._first_obs is None else self._first_obs.clone(),
            "goal_state": goal,
        }

    def load_state_dict(self, sd: dict[str, Any]) -> None:
        meta = sd.get("meta")
        if meta is not None:
            mine = {"kind": self.kind, "num_envs": self.num_envs, "global_num_envs": self.global_num_envs,
                    "env_lo": self.env_lo, "dtype": self.dtype, "n_contexts": len(self._table)}
            for k, v in mine.items():
                assert meta.get(k) == v, f"state dict was saved from a different env ({k}: {meta.get(k)!r} != {v!r})"
        for name, buf in (("state", self._state), ("elapsed", self._elapsed), ("sbt", self._sbt), ("rng", self._rng),
                          ("obs", self._obs)):
            t = sd[name]
            assert tuple(t.shape) == tuple(buf.shape) and t.dtype == buf.dtype, \
                f"state dict entry {name!r}: {tuple(t.shape)} {t.dtype}, expected {tuple(buf.shape)} {buf.dtype}"
        if self._first_state is not None and sd.get("seed") is not None:
            # Brax: the reset noise is keyed by (seed, global env id, episode counter). Re-key the handle FIRST
            # (seeding zeroes the counters), then restore the counters with the rng buffer below.
            _native.check(self._lib.carlb_env_seed(self._handle, int(sd["seed"]), self._stream()))
        self._state.copy_(sd["state"]); self._elapsed.copy_(sd["elapsed"]); self._sbt.copy_(sd["sbt"])
        self._rng.copy_(sd["rng"]); self._obs.copy_(sd["obs"])
        for name, buf in (("reward", self._reward), ("terminated", self._terminated), ("truncated", self._truncated)):
            if sd.get(name) is not None:
                buf.copy_(sd[name])
        if self._first_state is not None and sd.get("first_state") is not None:
            self._first_state.copy_(sd["first_state"]); self._first_obs.copy_(sd["first_obs"])
        self._context_ids = np.asarray(sd["context_ids"], dtype=np.int64).copy()
        if sd.get("reset_counts") is not None:
            self._reset_counts = np.asarray(sd["reset_counts"], dtype=np.int64).copy()
            self._rr_offset = np.asarray(sd["rr_offset"], dtype=np.int64).copy()
        sel_sd = sd.get("selector") or {}
        for k in ("context_id", "n_calls"):
            if sel_sd.get(k) is not None and hasattr(self.context_selector, k):
                setattr(self.context_selector, k, sel_sd[k])
        if sd.get("goal_state") is not None:
            self._goal_state = {k: (v.clone().to(self.device) if isinstance(v, torch.Tensor) else v)
                                for k, v in sd["goal_state"].items()}
            self._goal_strings = None
        self._refresh_context_view()
        self._update_context()
        self._seed_value = sd.get("seed", self._seed_value)
        self._seeded = bool(sd.get("seeded", True))
        self._has_reset = bool(sd.get("has_reset", True))

    # raw views for tests / learners
    @property
    def state(self) -> torch.Tensor:
        return self._state

    @property
    def unwrapped(self):
        return self

    def render(self):
        raise NotImplementedError("rendering is out of scope of the batched-step engine (DESIGN.md)")


class _LazyContexts(dict):
    """Dict view of the env's context table handed to selectors without materialising
    ``len(contexts)`` Python dicts up front (they only need ``len``, ``keys`` and item access)."""

    def __init__(self, env: CARLEnv):
        super().__init__()
        self._env = env

    def __len__(self):
        return len(self._env._table)

    def keys(self):
        return self._env._table.keys

    def __iter__(self):
        return iter(self._env._table.keys)

    def __getitem__(self, k):
        t = self._env._table
        if self._env._contexts_dict is not None:
            return self._env._contexts_dict[k]
        row = t.values[t.index_of(k)]
        return {n: _pyval(v) for n, v in zip(t.names, row)}

    def __contains__(self, k):
        return k in self._env._table.keys

    def items(self):
        return self._env.contexts.items()

    def values(self):
        return self._env.contexts.values()


def _pyval(v):
    f = float(v)
    return f
