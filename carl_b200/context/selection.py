"""Context selectors.

Same classes, attributes and selection rules as the reference's
``carl/context/selection.py:11-180``. Added for the batched engine:
``select_batch(n)`` — the ids ``n`` consecutive ``select()`` calls would return,
computed vectorised for the built-in selectors (the batched env resets ``n`` env
instances "in env-index order" against one shared selector).
"""
from __future__ import annotations

from abc import abstractmethod
from typing import Any, Callable, List, Optional, Tuple

import numpy as np

from carl_b200.utils.types import Context, Contexts


class AbstractSelector(object):
    """Reference: ``selection.py:11-95``."""

    def __init__(self, contexts: Contexts):
        self.contexts: Contexts = contexts
        self.context_ids: List[int] = list(np.arange(len(contexts)))
        self.contexts_keys: List[Any] = list(contexts.keys())
        self.n_calls: int = 0
        self.context_id: Optional[int] = None

    @abstractmethod
    def _select(self) -> Tuple[Context, int]:
        ...

    def select(self) -> Context:
        """``selection.py:64-76``."""
        context, context_id = self._select()
        self.context_id = context_id
        self.n_calls += 1
        return context

    def select_batch(self, n: int) -> np.ndarray:
        """Ids of ``n`` consecutive ``select()`` calls (generic fallback: the loop itself)."""
        ids = np.empty(n, dtype=np.int64)
        for i in range(n):
            self.select()
            ids[i] = self.context_id
        return ids

    @property
    def context_key(self) -> Any | None:
        """``selection.py:78-95`` (returns None for id 0 as the reference does)."""
        if self.context_id:
            key = self.contexts_keys[self.context_id]
        else:
            key = None
        return key


class RandomSelector(AbstractSelector):
    """``selection.py:98-107``: ``np.random.choice`` on the global legacy RNG."""

    def _select(self) -> Tuple[Context, int]:
        context_id = np.random.choice(self.context_ids)
        context = self.contexts[self.contexts_keys[context_id]]
        return context, context_id


class RoundRobinSelector(AbstractSelector):
    """``selection.py:110-122``."""

    def _select(self) -> Tuple[Context, int]:
        if self.context_id is None:
            self.context_id = -1
        self.context_id = (self.context_id + 1) % len(self.contexts)
        context = self.contexts[self.contexts_keys[self.context_id]]
        return context, self.context_id

    def select_batch(self, n: int) -> np.ndarray:
        if n == 0:
            return np.empty(0, dtype=np.int64)
        start = -1 if self.context_id is None else int(self.context_id)
        ids = (start + 1 + np.arange(n, dtype=np.int64)) % len(self.contexts)
        self.context_id = int(ids[-1])
        self.n_calls += n
        return ids


class StaticSelector(AbstractSelector):
    """``selection.py:125-136``."""

    def _select(self) -> Tuple[Context, int]:
        if self.context_id is None:
            self.context_id = self.context_ids[0]
        context = self.contexts[self.contexts_keys[self.context_id]]
        return context, self.context_id

    def select_batch(self, n: int) -> np.ndarray:
        if self.context_id is None:
            self.context_id = self.context_ids[0]
        self.n_calls += n
        return np.full(n, int(self.context_id), dtype=np.int64)


class CustomSelector(AbstractSelector):
    """``selection.py:139-180``."""

    def __init__(
        self,
        contexts: Contexts,
        selector_function: Callable[[AbstractSelector], Tuple[Context, int]],
    ):
        super().__init__(contexts=contexts)
        self.selector_function = selector_function

    def _select(self) -> Tuple[Context, int]:
        context, context_id = self.selector_function(self)
        self.context_id = context_id
        return context, context_id
