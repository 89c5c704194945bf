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


class ContextSpace:
    """The set of context features of one env class: names, defaults, bounds, validation and the
    observation-space view. Public surface of the reference's ``ContextSpace``
    (``carl/context/context_space.py:31-229``), same results."""

    def __init__(self, context_space: dict[str, ContextFeature]) -> None:
        self.context_space = context_space

    @property
    def context_feature_names(self) -> list[str]:
        return [*self.context_space]

    def get_default_context(self) -> Context:
        return {f.name: f.default_value for f in self.context_space.values()}

    def insert_defaults(self, context: Context, context_keys: List[str] | None = None) -> Context:
        """Defaults (restricted to ``context_keys`` when given) overridden by ``context``."""
        base = self.get_default_context()
        if context_keys:
            base = {k: base[k] for k in context_keys}
        return {**base, **context}

    def verify_context(self, context: Context) -> bool:
        """True iff every name is a known feature and every numerical value lies in its bounds."""
        for name, value in context.items():
            feature = self.context_space.get(name)
            if feature is None:
                return False
            if isinstance(feature, NumericalContextFeature) and not (feature.lower <= value <= feature.upper):
                return False
        return True

    def get_lower_and_upper_bound(self, context_feature_name: str) -> tuple[float, float]:
        f = self.context_space[context_feature_name]
        return f.lower, f.upper

    def to_gymnasium_space(self, context_feature_names: List[str] | None = None, as_dict: bool = False):
        """Observation-space view of (a subset of) the features: a ``Dict`` of per-feature ``Box`` /
        ``Discrete`` spaces, or one float32 ``Box`` over the bounds vector. gymnasium's classes are
        used when importable, else the stand-ins of ``carl_b200.spaces``."""
        names = self.context_feature_names if context_feature_names is None else context_feature_names
        feats = [self.context_space[n] for n in names]
        if not as_dict:
            return spaces.Box(low=np.array([f.lower for f in feats]), high=np.array([f.upper for f in feats]),
                              dtype=np.float32)
        per_feature = {}
        for f in feats:
            numeric = isinstance(f, NumericalContextFeature)
            per_feature[f.name] = spaces.Box(low=f.lower, high=f.upper) if numeric else spaces.Discrete(len(f.choices))
        return spaces.Dict(per_feature)

    def sample_contexts(self, context_keys: List[str] | None = None, size: int = 1) -> Context | List[Contexts]:
        """Unseeded samples of every feature (``rvs``), merged over the defaults of ``context_keys``.
        A feature whose distribution is unbounded keeps its default. ``size == 1`` returns the bare
        context, otherwise a list."""
        if context_keys is None:
            context_keys = list(self.context_space)
        bad = [k for k in context_keys if k not in self.context_space]
        if bad:
            raise ValueError(f"Invalid context feature name: {bad[0]}")

        def draw(f: ContextFeature):
            try:
                return f.rvs()
            except ValueError:
                return f.default_value

        drawn = [self.insert_defaults({f.name: draw(f) for f in self.context_space.values()}, list(context_keys))
                 for _ in range(size)]
        return drawn[0] if size == 1 else drawn
