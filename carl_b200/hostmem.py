"""Page-locked host arrays for the host-buffer step path.

``CARLEnv.step(numpy_actions)`` has to get the actions into memory the GPU can read. For an ordinary
(pageable) numpy array that costs one staging copy per step; an array from :func:`pinned_empty` is
already page-locked and mapped, so the step kernel reads it in place over PCIe (``carlb_env_step_host``'s
zero-copy path, ``carl_b200/csrc/abi.cu``). A policy that writes its actions straight into such an
array saves the copy.
"""
from __future__ import annotations

import bisect

import numpy as np
import torch

_starts: list[int] = []   # sorted start addresses of the registered blocks
_blocks: dict[int, tuple[int, torch.Tensor]] = {}  # start -> (end, owner tensor kept alive)


def pinned_empty(shape, dtype=np.int32) -> np.ndarray:
    """A C-contiguous numpy array in page-locked, device-mapped host memory."""
    dtype = np.dtype(dtype)
    n_bytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    t = torch.empty(max(n_bytes, 1), dtype=torch.uint8).pin_memory()  # needs the CUDA driver: no CPU fallback
    _register(t)
    return t.numpy()[:n_bytes].view(dtype).reshape(shape)


def _register(t: torch.Tensor) -> None:
    start = t.data_ptr()
    bisect.insort(_starts, start)
    _blocks[start] = (start + t.numel() * t.element_size(), t)


def is_pinned(addr: int, n_bytes: int) -> bool:
    """True when ``[addr, addr + n_bytes)`` lies inside a block handed out by :func:`pinned_empty`."""
    k = bisect.bisect_right(_starts, addr) - 1
    return k >= 0 and addr + n_bytes <= _blocks[_starts[k]][0]


def release(array: np.ndarray) -> None:
    """Forget (and free, once the array is gone) the block ``array`` starts in."""
    k = bisect.bisect_right(_starts, array.ctypes.data) - 1
    if k >= 0 and array.ctypes.data < _blocks[_starts[k]][0]:
        _blocks.pop(_starts.pop(k))
