"""Seeded context sampling.

Reference: ``carl/context/sampler.py:11-61`` — ``ContextSampler`` subclasses ConfigSpace's
``ConfigurationSpace``; here it subclasses :class:`FeatureSpace`, which reproduces the
sampling stream pinned by the reference notebooks (tests/golden/context_sampler.json).
"""
from __future__ import annotations

from typing import Any

import numpy as np

from carl_b200.context.context_space import ContextFeature, ContextSpace
from carl_b200.context.search_space_encoding import FeatureSpace, search_space_to_config_space
from carl_b200.utils.types import Context, Contexts


class ContextSampler(FeatureSpace):
    def __init__(
        self,
        context_distributions: list[ContextFeature] | dict[str, ContextFeature] | str | Any,
        context_space: ContextSpace,
        seed: int,
        name: str | None = None,
    ):
        self.context_distributions = context_distributions
        super().__init__(name=name, seed=seed)

        if isinstance(context_distributions, list):
            self.add_context_features(context_distributions)
        elif isinstance(context_distributions, dict) and all(
            isinstance(v, ContextFeature) for v in context_distributions.values()
        ):
            self.add_context_features(context_distributions.values())
        elif isinstance(context_distributions, str) or (
            hasattr(context_distributions, "items") and "hyperparameters" in context_distributions
        ):
            cs = search_space_to_config_space(context_distributions)
            self.add_context_features(cs.get_hyperparameters())
        else:
            raise ValueError(
                f"Unknown type `{type(context_distributions)}` for `context_distributions`."
            )

        self.context_feature_names = [cf.name for cf in self.get_context_features()]
        self.context_space = context_space

    def add_context_features(self, context_features) -> None:
        self.add_hyperparameters(context_features)

    def get_context_features(self) -> list[ContextFeature]:
        return list(self.values())

    def sample_contexts(self, n_contexts: int) -> Contexts:
        """``sampler.py:45-51``: ``{i: context}`` for ``i in range(n_contexts)``."""
        contexts = self._sample_contexts(size=n_contexts)
        return {i: C for i, C in enumerate(contexts)}

    def _sample_contexts(self, size: int = 1) -> list[Context]:
        """``sampler.py:53-61``: samples merged over the context space's defaults."""
        contexts = self.sample_configuration(size=size)
        default_context = self.context_space.get_default_context()
        if size == 1:
            contexts = [contexts]
        return [dict(default_context | dict(C)) for C in contexts]

    def sample_context_table(self, n_contexts: int, feature_names: list[str] | None = None) -> np.ndarray:
        """Batched fast path: the same draws as :meth:`sample_contexts` as a dense
        ``float64[n_contexts, F]`` table in ``feature_names`` order (defaults filled), without
        building ``n_contexts`` Python dicts. Consumes the sampler's stream identically."""
        default_context = self.context_space.get_default_context()
        if feature_names is None:
            feature_names = list(default_context.keys())
        n = int(n_contexts)
        cols = {f.name: f.sample_column(self.random, n) for f in self.values()}
        table = np.empty((n, len(feature_names)), dtype=np.float64)
        for j, k in enumerate(feature_names):
            if k in cols:
                table[:, j] = np.asarray(cols[k], dtype=np.float64)
            else:
                table[:, j] = float(default_context[k])
        return table
