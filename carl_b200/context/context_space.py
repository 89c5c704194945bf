"""Context features and the context space.

Mirrors the plugin surface of the reference's ``carl/context/context_space.py:23-229``.
The reference aliases ConfigSpace hyperparameter classes (``context_space.py:23-28``);
ConfigSpace is a third-party dependency that is absent here, so the feature classes
are re-implemented with the constructor keywords CARL's env tables use
(``name, lower, upper, default_value`` / ``mu, sigma`` / ``choices``).

Sampling follows the stream pinned by the reference's notebooks:
one ``numpy.random.RandomState`` per sampler, whole-column draws.
"""
from __future__ import annotations

from typing import Any, List, Sequence

import numpy as np

from carl_b200 import spaces
from carl_b200.utils.types import Context, Contexts


class ContextFeature:
    """Base class of a context feature (reference: ConfigSpace ``Hyperparameter``)."""

    def __init__(self, name: str, default_value: Any = None, meta: dict | None = None):
        if not isinstance(name, str):
            raise TypeError(f"Context feature name must be a str, got {type(name)}.")
        self.name = name
        self.default_value = default_value
        self.meta = meta

    # -- sampling -----------------------------------------------------------------
    def sample_column(self, rs: np.random.RandomState, size: int) -> np.ndarray:
        """Draw ``size`` values at once from ``rs`` (one column of a context table)."""
        raise NotImplementedError

    def rvs(self, size: int | None = None, random_state: Any = None) -> Any:
        """Unseeded draw, as ``ContextSpace.sample_contexts`` uses it (``context_space.py:222``)."""
        if isinstance(random_state, np.random.RandomState):
            rs = random_state
        else:
            rs = np.random.RandomState(random_state)
        col = self.sample_column(rs, 1 if size is None else size)
        return col[0] if size is None else col

    def __repr__(self) -> str:
        return f"{type(self).__name__}({self.name!r}, default={self.default_value!r})"


class NumericalContextFeature(ContextFeature):
    def __init__(self, name, lower, upper, default_value=None, log=False, meta=None):
        super().__init__(name, default_value, meta)
        self.lower = lower
        self.upper = upper
        self.log = bool(log)
        if lower is not None and upper is not None and lower > upper:
            raise ValueError(f"{name}: lower {lower} > upper {upper}")

    def legal(self, v) -> bool:
        lo = -np.inf if self.lower is None else self.lower
        hi = np.inf if self.upper is None else self.upper
        return bool(lo <= v <= hi)


class UniformFloatContextFeature(NumericalContextFeature):
    """Reference alias of ``UniformFloatHyperparameter`` (``context_space.py:26``)."""

    def __init__(self, name, lower, upper, default_value=None, log=False, meta=None):
        super().__init__(name, lower, upper, default_value, log, meta)
        if self.default_value is None:
            if np.isfinite(lower) and np.isfinite(upper):
                self.default_value = (
                    float(np.exp((np.log(lower) + np.log(upper)) / 2)) if log else (lower + upper) / 2
                )

    def sample_column(self, rs, size):
        if not (np.isfinite(self.lower) and np.isfinite(self.upper)):
            raise ValueError(
                f"Cannot sample uniformly from unbounded context feature {self.name!r} "
                f"[{self.lower}, {self.upper}]."
            )
        u = rs.uniform(size=size)
        if self.log:
            lo, hi = np.log(self.lower), np.log(self.upper)
            return np.exp(lo + u * (hi - lo))
        return self.lower + u * (self.upper - self.lower)


class UniformIntegerContextFeature(NumericalContextFeature):
    """Reference alias of ``UniformIntegerHyperparameter`` (``context_space.py:27``)."""

    def __init__(self, name, lower, upper, default_value=None, log=False, meta=None):
        super().__init__(name, int(lower), int(upper), default_value, log, meta)
        if self.default_value is None:
            self.default_value = int(round((self.lower + self.upper) / 2))

    def sample_column(self, rs, size):
        u = rs.uniform(size=size)
        v = np.floor(self.lower + u * (self.upper - self.lower + 1)).astype(np.int64)
        return np.clip(v, self.lower, self.upper)


class NormalFloatContextFeature(NumericalContextFeature):
    """Reference alias of ``NormalFloatHyperparameter`` (``context_space.py:25``)."""

    def __init__(self, name, mu, sigma, lower=None, upper=None, default_value=None, log=False, meta=None):
        super().__init__(name, lower, upper, default_value, log, meta)
        self.mu = mu
        self.sigma = sigma
        if self.default_value is None:
            self.default_value = mu

    def sample_column(self, rs, size):
        v = rs.normal(self.mu, self.sigma, size)
        if self.lower is not None and self.upper is not None:
            v = np.clip(v, self.lower, self.upper)
        return v


class CategoricalContextFeature(ContextFeature):
    """Reference alias of ``CategoricalHyperparameter`` (``context_space.py:28``)."""

    def __init__(self, name, choices: Sequence[Any], weights=None, default_value=None, meta=None):
        super().__init__(name, default_value, meta)
        self.choices = list(choices)
        if len(self.choices) == 0:
            raise ValueError(f"{name}: empty choices")
        self.weights = None if weights is None else list(weights)
        if self.default_value is None:
            self.default_value = self.choices[0]
        elif self.default_value not in self.choices:
            raise ValueError(f"{name}: default {default_value!r} not in choices")

    def sample_column(self, rs, size):
        if self.weights is None:
            idx = np.floor(rs.random_sample(size) * len(self.choices)).astype(np.int64)
        else:
            p = np.asarray(self.weights, dtype=np.float64)
            idx = rs.choice(len(self.choices), size=size, p=p / p.sum())
        return np.asarray(self.choices, dtype=object)[idx]


class ContextSpace(object):
    """Reference: ``carl/context/context_space.py:31-229`` (same methods, same results)."""

    def __init__(self, context_space: dict[str, ContextFeature]) -> None:
        self.context_space = context_space

    @property
    def context_feature_names(self) -> list[str]:
        return list(self.context_space.keys())

    def insert_defaults(self, context: Context, context_keys: List[str] | None = None) -> Context:
        """``context_space.py:54-80``: defaults (optionally only ``context_keys``) updated by ``context``."""
        context_with_defaults = self.get_default_context()
        if context_keys:
            context_with_defaults = {key: context_with_defaults[key] for key in context_keys}
        context_with_defaults.update(context)
        return context_with_defaults

    def verify_context(self, context: Context) -> bool:
        """``context_space.py:82-114``: names known and numerical values in bounds."""
        is_valid = True
        cfs = self.context_feature_names
        for cfname, v in context.items():
            if cfname not in cfs:
                is_valid = False
                break
            cf = self.context_space[cfname]
            if isinstance(cf, NumericalContextFeature):
                if not (cf.lower <= v <= cf.upper):
                    is_valid = False
                    break
        return is_valid

    def get_default_context(self) -> Context:
        """``context_space.py:116-125``."""
        return {cf.name: cf.default_value for cf in self.context_space.values()}

    def get_lower_and_upper_bound(self, context_feature_name: str) -> tuple[float, float]:
        """``context_space.py:127-143``."""
        cf = self.context_space[context_feature_name]
        return (cf.lower, cf.upper)

    def to_gymnasium_space(self, context_feature_names: List[str] | None = None, as_dict: bool = False):
        """``context_space.py:145-188``. Uses gymnasium's spaces when importable, else the
        API-compatible stand-ins in ``carl_b200.spaces``."""
        if context_feature_names is None:
            context_feature_names = self.context_feature_names
        if as_dict:
            context_space = {}
            for cf_name in context_feature_names:
                context_feature = self.context_space[cf_name]
                if isinstance(context_feature, NumericalContextFeature):
                    context_space[context_feature.name] = spaces.Box(
                        low=context_feature.lower, high=context_feature.upper
                    )
                else:
                    context_space[context_feature.name] = spaces.Discrete(len(context_feature.choices))
            return spaces.Dict(context_space)
        low = np.array([self.context_space[cf].lower for cf in context_feature_names])
        high = np.array([self.context_space[cf].upper for cf in context_feature_names])
        return spaces.Box(low=low, high=high, dtype=np.float32)

    def sample_contexts(self, context_keys: List[str] | None = None, size: int = 1) -> Context | List[Contexts]:
        """``context_space.py:190-229``: unseeded ``rvs`` per feature; features outside
        ``context_keys`` are dropped *unless* sampled (reference behaviour: the sampled value of
        every feature overrides the inserted defaults)."""
        if context_keys is None:
            context_keys = self.context_space.keys()
        else:
            for key in context_keys:
                if key not in self.context_space.keys():
                    raise ValueError(f"Invalid context feature name: {key}")
        contexts = []
        for _ in range(size):
            context = {}
            for cf in self.context_space.values():
                try:
                    context[cf.name] = cf.rvs()
                except ValueError:
                    # unbounded uniform features cannot be sampled; keep the default
                    context[cf.name] = cf.default_value
            context = self.insert_defaults(context, list(context_keys))
            contexts += [context]
        if size == 1:
            return contexts[0]
        return contexts
