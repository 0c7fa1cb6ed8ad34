"""Small host helpers shared by the API mirror."""
from __future__ import annotations

import copy

import numpy as np
import torch


def device() -> torch.device:
    """Where scene tensors live: the current CUDA device, or the CPU when none is visible
    (host-side logic stays usable there; rendering raises)."""
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def f32(x, dev=None) -> torch.Tensor:
    """``jnp.asarray(x)`` analogue: float32 tensor on the scene device (keeps autograd leaves as they are)."""
    dev = dev or device()
    if isinstance(x, torch.Tensor):
        if x.dtype == torch.float32 and x.device == dev:
            return x
        return x.to(device=dev, dtype=torch.float32)
    if hasattr(x, "__cuda_array_interface__") or hasattr(x, "__dlpack__") and not isinstance(x, np.ndarray):
        try:
            return torch.as_tensor(x, device=dev).to(torch.float32)
        except Exception:
            pass
    return torch.as_tensor(np.asarray(x, dtype=np.float32), device=dev)


def i32(x, dev=None) -> torch.Tensor:
    dev = dev or device()
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=torch.int32)
    return torch.as_tensor(np.asarray(x, dtype=np.int32), device=dev)


def replace(obj, **fields):
    """Functional update (the reference's ``eqx.tree_at``): shallow copy with some fields swapped.
    Derived caches are dropped so the copy never sees stale device tables."""
    new = copy.copy(obj)
    for k, v in fields.items():
        object.__setattr__(new, k, v)
    if hasattr(new, "_cache"):
        object.__setattr__(new, "_cache", {})
    return new


def contig(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()
