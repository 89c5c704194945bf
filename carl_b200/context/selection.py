"""Context selectors of the batched engine.

Same public surface as the reference's ``carl/context/selection.py`` (``AbstractSelector`` with
``contexts / context_ids / contexts_keys / n_calls / context_id / context_key / select()``, and the
``RandomSelector``, ``RoundRobinSelector``, ``StaticSelector``, ``CustomSelector`` policies), built
around one vectorised primitive: ``select_batch(n)`` returns the ids that ``n`` consecutive
``select()`` calls produce. The batched env resets many env instances against one shared selector,
so the scalar ``select()`` is just ``select_batch(1)`` here.
"""
from __future__ import annotations

from typing import Any, Callable, List, Optional, Tuple

import numpy as np

from carl_b200.utils.types import Context, Contexts


class AbstractSelector:
    """Bookkeeping shared by all selection policies (reference: ``selection.py:11-95``)."""

    def __init__(self, contexts: Contexts):
        self.contexts: Contexts = contexts
        self.contexts_keys: List[Any] = list(contexts.keys())
        self.context_ids: List[int] = list(np.arange(len(self.contexts_keys)))
        self.n_calls: int = 0
        self.context_id: Optional[int] = None  # index into contexts_keys; None until the first select

    # -- policy hook ---------------------------------------------------------------------------
    def _next_ids(self, n: int) -> np.ndarray:
        """Ids of the next ``n`` selections (policy specific); may read/update ``context_id``."""
        out = np.empty(n, dtype=np.int64)
        for j in range(n):  # generic fallback: a policy that only implements the scalar `_select`
            _, cid = self._select()
            self.context_id = cid
            out[j] = cid
        return out

    def _select(self) -> Tuple[Context, int]:  # scalar hook kept for subclasses written against the reference
        raise NotImplementedError

    # -- public API ----------------------------------------------------------------------------
    def select_batch(self, n: int) -> np.ndarray:
        ids = np.asarray(self._next_ids(int(n)), dtype=np.int64)
        if ids.size:
            self.context_id = int(ids[-1])
        self.n_calls += int(n)
        return ids

    def select(self) -> Context:
        cid = int(self.select_batch(1)[0])
        return self.contexts[self.contexts_keys[cid]]

    @property
    def context_key(self) -> Any | None:
        # the reference tests truthiness of the id, so id 0 reports None as well (selection.py:91-95)
        return self.contexts_keys[self.context_id] if self.context_id else None


class RoundRobinSelector(AbstractSelector):
    """Cycle through the context set in key order (reference: ``selection.py:110-122``)."""

    def _next_ids(self, n: int) -> np.ndarray:
        start = -1 if self.context_id is None else int(self.context_id)
        return (start + 1 + np.arange(n, dtype=np.int64)) % len(self.contexts_keys)


class StaticSelector(AbstractSelector):
    """Never change the context (reference: ``selection.py:125-136``)."""

    def _next_ids(self, n: int) -> np.ndarray:
        cid = self.context_ids[0] if self.context_id is None else self.context_id
        return np.full(n, int(cid), dtype=np.int64)


class RandomSelector(AbstractSelector):
    """Uniformly random context per selection, drawn one at a time from NumPy's global legacy
    generator exactly as the reference does (``np.random.choice``, ``selection.py:98-107``)."""

    def _next_ids(self, n: int) -> np.ndarray:
        return np.fromiter((np.random.choice(self.context_ids) for _ in range(n)), dtype=np.int64, count=n)


class CustomSelector(AbstractSelector):
    """User policy: ``selector_function(selector) -> (context, context_id)``
    (reference: ``selection.py:139-180``; e.g. ``lambda s: (s.contexts[s.contexts_keys[1]], 1)``)."""

    def __init__(self, contexts: Contexts, selector_function: Callable[[AbstractSelector], Tuple[Context, int]]):
        super().__init__(contexts=contexts)
        self.selector_function = selector_function

    def _next_ids(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.int64)
        for j in range(n):
            _, cid = self.selector_function(self)
            self.context_id = cid
            out[j] = cid
            if j + 1 < n:
                self.n_calls += 1  # the user function may key on n_calls (reference docstring example)
        if n > 1:
            self.n_calls -= n - 1  # select_batch adds n once
        return out
