"""Euler-angle transforms (mirror of reference ``iactrace/core/transforms.py:72-106``).

The kernels evaluate the same formula on the device (``csrc/iact_common.cuh``); this host copy
exists for API parity and for building scenes.
"""
from __future__ import annotations

import math

import torch

from .._util import f32


def euler_to_matrix(tip_tilt_rotation) -> torch.Tensor:
    """Degrees -> rotation matrix ``Rz(rotation) @ Ry(tilt) @ Rx(tip)`` (local -> world)."""
    a = f32(tip_tilt_rotation) * (math.pi / 180.0)
    rx, ry, rz = a[0], a[1], a[2]
    one, zero = torch.ones_like(rx), torch.zeros_like(rx)
    cx, sx, cy, sy, cz, sz = torch.cos(rx), torch.sin(rx), torch.cos(ry), torch.sin(ry), torch.cos(rz), torch.sin(rz)
    Rx = torch.stack([torch.stack([one, zero, zero]), torch.stack([zero, cx, -sx]), torch.stack([zero, sx, cx])])
    Ry = torch.stack([torch.stack([cy, zero, sy]), torch.stack([zero, one, zero]), torch.stack([-sy, zero, cy])])
    Rz = torch.stack([torch.stack([cz, -sz, zero]), torch.stack([sz, cz, zero]), torch.stack([zero, zero, one])])
    return Rz @ Ry @ Rx


def look_at_rotation(mirror_pos, target_pos=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0)) -> torch.Tensor:
    """Mirror of ``transforms.py:4-24`` (config-authoring helper)."""
    mirror_pos, target_pos, up = f32(mirror_pos), f32(target_pos), f32(up)
    fwd = target_pos - mirror_pos
    fwd = fwd / torch.linalg.norm(fwd)
    right = torch.linalg.cross(fwd, up)
    right = right / torch.linalg.norm(right)
    upc = torch.linalg.cross(right, fwd)
    return torch.stack([right, upc, fwd], dim=1)
