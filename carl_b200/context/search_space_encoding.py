"""Hydra / JSON search-space → context-feature list.

Reference: ``carl/context/search_space_encoding.py:46-144`` (which converts to a ConfigSpace
``ConfigurationSpace`` through ``ConfigSpace.read_and_write.json``). ConfigSpace and omegaconf
are absent third-party packages; this module reads the same documents (the ConfigSpace JSON
layout shown in the reference docstring at :80-104 and the hydra ``hyperparameters:{name:cfg}``
mapping at :112-128) into this package's feature classes. omegaconf ``DictConfig`` objects are
accepted duck-typed (anything with ``.items()``).
"""
from __future__ import annotations

import json
from typing import Any, Mapping

from carl_b200.context.context_space import (
    CategoricalContextFeature,
    ContextFeature,
    NormalFloatContextFeature,
    UniformFloatContextFeature,
    UniformIntegerContextFeature,
)


class FeatureSpace:
    """Ordered, seedable bag of context features (stand-in for ``ConfigurationSpace``)."""

    def __init__(self, name: str | None = None, seed: int | None = None, space: Any = None):
        import numpy as np

        self.name = name
        self._features: dict[str, ContextFeature] = {}
        self._seed = seed
        self.random = np.random.RandomState(seed)
        if space:
            self.add_hyperparameters(_features_from_shorthand(space))

    # ConfigurationSpace-compatible surface used by the reference (sampler.py:22-43,54)
    def seed(self, seed: int | None = None) -> None:
        import numpy as np

        self._seed = seed
        self.random = np.random.RandomState(seed)

    def add_hyperparameters(self, features) -> None:
        for f in features:
            if not isinstance(f, ContextFeature):
                raise TypeError(f"Expected a ContextFeature, got {type(f)}.")
            if f.name in self._features:
                raise ValueError(f"Context feature {f.name!r} already present.")
            self._features[f.name] = f

    add = add_hyperparameters

    def get_hyperparameters(self) -> list[ContextFeature]:
        return list(self.values())

    def values(self):
        # ConfigSpace keeps hyperparameters sorted by name; the sampler stream depends on it
        return [self._features[k] for k in sorted(self._features)]

    def keys(self):
        return sorted(self._features)

    def items(self):
        return [(k, self._features[k]) for k in sorted(self._features)]

    def __getitem__(self, k):
        return self._features[k]

    def __len__(self):
        return len(self._features)

    def __contains__(self, k):
        return k in self._features

    def sample_configuration(self, size: int | None = None):
        """Column-wise draws in alphabetical feature order from the space's RandomState
        (reproduces the reference notebooks' ``ContextSampler(seed=0)`` output)."""
        n = 1 if size is None else int(size)
        cols = {f.name: f.sample_column(self.random, n) for f in self.values()}
        rows = [{k: _to_py(cols[k][i]) for k in cols} for i in range(n)]
        if size is None or size == 1:
            return rows[0]
        return rows


def _to_py(v):
    import numpy as np

    if isinstance(v, np.generic):
        return v.item()
    return v


def _feature_from_json(cfg: Mapping[str, Any]) -> ContextFeature:
    cfg = dict(cfg)
    typ = cfg.get("type")
    name = cfg["name"]
    default = cfg.get("default", cfg.get("default_value"))
    log = bool(cfg.get("log", False))
    if typ == "uniform_float":
        return UniformFloatContextFeature(name, cfg["lower"], cfg["upper"], default_value=default, log=log)
    if typ == "uniform_int":
        return UniformIntegerContextFeature(name, cfg["lower"], cfg["upper"], default_value=default, log=log)
    if typ == "normal_float":
        return NormalFloatContextFeature(
            name, cfg["mu"], cfg["sigma"], lower=cfg.get("lower"), upper=cfg.get("upper"),
            default_value=default, log=log,
        )
    if typ == "categorical":
        return CategoricalContextFeature(name, list(cfg["choices"]), weights=cfg.get("weights"), default_value=default)
    if typ == "constant":
        return CategoricalContextFeature(name, [cfg["value"]], default_value=cfg["value"])
    raise ValueError(f"Unsupported hyperparameter type {typ!r} for {name!r}.")


def _features_from_shorthand(space: Mapping[str, Any]) -> list[ContextFeature]:
    """ConfigSpace's dict shorthand: ``(int,int)`` → uniform int, ``(float,float)`` → uniform
    float, list → categorical, scalar → constant."""
    out: list[ContextFeature] = []
    for name, v in space.items():
        if isinstance(v, ContextFeature):
            out.append(v)
        elif isinstance(v, tuple) and len(v) == 2:
            if all(isinstance(x, int) and not isinstance(x, bool) for x in v):
                out.append(UniformIntegerContextFeature(name, v[0], v[1]))
            else:
                out.append(UniformFloatContextFeature(name, float(v[0]), float(v[1])))
        elif isinstance(v, list):
            out.append(CategoricalContextFeature(name, v))
        else:
            out.append(CategoricalContextFeature(name, [v], default_value=v))
    return out


def search_space_to_config_space(search_space: Any, seed: int | None = None) -> FeatureSpace:
    """Reference: ``search_space_encoding.py:46-144``. Accepts a path to a ConfigSpace JSON file,
    a hydra-style mapping (``{"hyperparameters": {name: cfg}}``), a ConfigSpace-JSON-style dict
    (``{"hyperparameters": [cfg, ...]}``) or an existing :class:`FeatureSpace`."""
    if isinstance(search_space, FeatureSpace):
        cs = search_space
    else:
        if isinstance(search_space, str):
            with open(search_space, "r") as f:
                doc = json.loads(f.read())
        elif hasattr(search_space, "items"):
            doc = {k: v for k, v in search_space.items()}
        else:
            raise ValueError(
                f"search_space must be of type str or DictConfig. Got {type(search_space)}."
            )
        hps = doc.get("hyperparameters", [])
        if hasattr(hps, "items"):  # hydra layout: name -> cfg (search_space_encoding.py:114-122)
            hp_list = []
            for name, cfg in hps.items():
                cfg = dict(cfg)
                cfg["name"] = name
                cfg.setdefault("default", None)
                cfg.setdefault("log", False)
                hp_list.append(cfg)
            hps = hp_list
        cs = FeatureSpace(name=doc.get("name"))
        cs.add_hyperparameters([_feature_from_json(h) for h in hps])
    if seed is not None:
        cs.seed(seed=seed)
    return cs
