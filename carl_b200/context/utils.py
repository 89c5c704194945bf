"""Bounds helper of the context plugin surface (reference counterpart: ``carl/context/utils.py``,
``get_context_bounds``)."""
from __future__ import annotations

from typing import Any, Mapping, Sequence

import numpy as np


def get_context_bounds(context_keys: Sequence[str], context_bounds: Mapping[str, tuple[float, float, Any]]):
    """``(lower, upper)`` float arrays for ``context_keys`` out of ``{name: (lower, upper, dtype)}``."""
    picked = [context_bounds[k] for k in context_keys]
    lower = np.fromiter((b[0] for b in picked), dtype=np.float64, count=len(picked))
    upper = np.fromiter((b[1] for b in picked), dtype=np.float64, count=len(picked))
    return lower, upper
