"""Mirror of reference ``iactrace/core/reflection.py:5-19`` (host/torch form of the in-kernel op)."""
from __future__ import annotations

import torch


def reflect(d: torch.Tensor, n: torch.Tensor):
    """Reflect direction ``d`` off normal ``n`` -> (reflected, -cos)."""
    cos_angle = torch.sum(d * n, dim=-1, keepdim=True)
    return d - 2.0 * cos_angle * n, -cos_angle
