"""Observation / action space descriptors.

The reference builds ``gymnasium.spaces.{Box,Dict,Discrete}`` objects
(``carl/envs/carl_env.py:159-188``, ``carl/context/context_space.py:145-188``).
gymnasium is an optional third-party package: if it is importable its classes are
re-exported unchanged, otherwise API-compatible stand-ins (shape/dtype/low/high,
``contains``, ``sample``) are used so that the host layer has no hard dependency.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any

import numpy as np

try:  # pragma: no cover - exercised only where gymnasium is installed
    from gymnasium.spaces import Box, Dict, Discrete, Space  # type: ignore

    HAVE_GYMNASIUM = True
except Exception:  # gymnasium absent
    HAVE_GYMNASIUM = False

    class Space:
        shape: tuple | None = None
        dtype: Any = None

        def __init__(self, shape=None, dtype=None, seed=None):
            self.shape = None if shape is None else tuple(shape)
            self.dtype = None if dtype is None else np.dtype(dtype)
            self._np_random = np.random.default_rng(seed)

        def seed(self, seed=None):
            self._np_random = np.random.default_rng(seed)
            return [seed]

        def contains(self, x) -> bool:  # pragma: no cover
            raise NotImplementedError

        def __contains__(self, x) -> bool:
            return self.contains(x)

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            low_a = np.asarray(low, dtype=np.float64)
            high_a = np.asarray(high, dtype=np.float64)
            if shape is None:
                shape = np.broadcast(low_a, high_a).shape
                if shape == ():
                    shape = (1,)
            super().__init__(shape, dtype, seed)
            with np.errstate(over="ignore"):
                self.low = np.broadcast_to(low_a, self.shape).astype(self.dtype)
                self.high = np.broadcast_to(high_a, self.shape).astype(self.dtype)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return bool(
                x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high)
            )

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return self._np_random.uniform(lo, hi).astype(self.dtype)

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

        def __eq__(self, other):
            return (
                isinstance(other, Box)
                and self.shape == other.shape
                and np.array_equal(self.low, other.low)
                and np.array_equal(self.high, other.high)
            )

    class Discrete(Space):
        def __init__(self, n: int, seed=None, start: int = 0):
            super().__init__((), np.int64, seed)
            self.n = int(n)
            self.start = int(start)

        def contains(self, x) -> bool:
            try:
                xi = int(x)
            except Exception:
                return False
            return self.start <= xi < self.start + self.n

        def sample(self):
            return int(self.start + self._np_random.integers(self.n))

        def __repr__(self):
            return f"Discrete({self.n})"

        def __eq__(self, other):
            return isinstance(other, Discrete) and self.n == other.n and self.start == other.start

    class Dict(Space):
        def __init__(self, spaces=None, seed=None, **kw):
            super().__init__(None, None, seed)
            self.spaces = OrderedDict(spaces or {})
            self.spaces.update(kw)

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def __len__(self):
            return len(self.spaces)

        def contains(self, x) -> bool:
            return isinstance(x, dict) and all(k in x and self.spaces[k].contains(x[k]) for k in self.spaces)

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def __repr__(self):
            return "Dict(" + ", ".join(f"{k!r}: {s!r}" for k, s in self.spaces.items()) + ")"


def batch_box(space: "Box", n: int) -> "Box":
    """Batched copy of a Box: shape ``(n,) + space.shape`` (reference precedent:
    ``gymnasium.vector.utils.batch_space`` at ``carl/envs/brax/wrappers.py:114,119``)."""
    low = np.broadcast_to(space.low, (n,) + tuple(space.shape)).copy()
    high = np.broadcast_to(space.high, (n,) + tuple(space.shape)).copy()
    return Box(low=low, high=high, dtype=space.dtype)
