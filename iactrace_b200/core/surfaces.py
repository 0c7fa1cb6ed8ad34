"""Aspheric surface value type (mirror of reference ``iactrace/core/surfaces.py:8-65``).

The device kernels carry the arithmetic (``sag_raw`` / ``dsag_dr2`` in ``csrc/iact_common.cuh``);
the torch methods here are the host-side API equivalents.
"""
from __future__ import annotations

import numpy as np
import torch

from .._util import f32


class AsphericSurface:
    """Surface z(r) = c r^2 / (1 + sqrt(1 - (1+k) c^2 r^2)) + sum a_i (r^2)^(2i+2)."""

    def __init__(self, curvature: float, conic: float, aspheric=()):
        self.curvature = float(curvature)
        self.conic = float(conic)
        a = aspheric.detach().cpu().numpy() if isinstance(aspheric, torch.Tensor) else aspheric
        self.aspheric = np.asarray(a, dtype=np.float32).reshape(-1)

    @staticmethod
    def from_template(tmpl) -> "AsphericSurface":
        """Convert a YAML template dict (``mirror_templates.<name>``) into a surface."""
        s = tmpl["surface"]
        return AsphericSurface(float(s["curvature"]), float(s["conic"]), s.get("aspheric", []))

    def params_key(self):
        return (self.curvature, self.conic, tuple(self.aspheric.tolist()))

    def _sag_raw(self, x, y):
        x, y = f32(x), f32(y)
        r2 = x * x + y * y
        c, k = self.curvature, self.conic
        z = r2 * c / (1 + torch.sqrt(1 - (1 + k) * c * c * r2))
        for i, a in enumerate(self.aspheric.tolist()):
            z = z + a * r2 ** (2 * i + 2)
        return z

    def sag(self, x, y, offset):
        offset = f32(offset)
        return self._sag_raw(f32(x) + offset[0], f32(y) + offset[1]) - self._sag_raw(offset[0], offset[1])

    def point(self, x, y, offset):
        x, y = f32(x), f32(y)
        return torch.stack([x, y, self.sag(x, y, offset)], dim=-1)

    def normal(self, x, y, offset):
        offset = f32(offset)
        xs, ys = f32(x) + offset[0], f32(y) + offset[1]
        r2 = xs * xs + ys * ys
        c, k = self.curvature, self.conic
        g = 0.5 * c / torch.sqrt(1 - (1 + k) * c * c * r2)
        for i, a in enumerate(self.aspheric.tolist()):
            g = g + a * (2 * i + 2) * r2 ** (2 * i + 1)
        n = torch.stack([-g * 2 * xs, -g * 2 * ys, torch.ones_like(xs)], dim=-1)
        return n / torch.linalg.norm(n, dim=-1, keepdim=True)

    def point_and_normal(self, xy, offset):
        xy = f32(xy)
        return self.point(xy[..., 0], xy[..., 1], offset), self.normal(xy[..., 0], xy[..., 1], offset)
