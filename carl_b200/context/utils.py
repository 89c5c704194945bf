"""Reference: ``carl/context/utils.py:6-35``."""
from typing import Any, Dict, List, Tuple, Type

import numpy as np


def get_context_bounds(
    context_keys: List[str], context_bounds: Dict[str, Tuple[float, float, Type[Any]]]
) -> Tuple[np.ndarray, np.ndarray]:
    """Lower / upper bound arrays for ``context_keys`` from ``{name: (lower, upper, dtype)}``."""
    lower_bounds = np.empty(shape=len(context_keys))
    upper_bounds = np.empty(shape=len(context_keys))
    for i, context_key in enumerate(context_keys):
        lower, upper, _dtype = context_bounds[context_key]
        lower_bounds[i] = lower
        upper_bounds[i] = upper
    return lower_bounds, upper_bounds
