"""Hexagonal-pixel sensors (mirror of reference ``iactrace/sensors/hexagonal.py``).

Grid detection and the axial lookup table are host-side, once per sensor, in float32 like the
reference (``_detect_hex_grid`` is tie-sensitive: SURVEY.md hazard H3 -- pass ``grid=`` to pin the
constants).  Pixel binning runs in the trace kernel (``hex_pixel`` / ``splat_soft_hex``).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native as N
from .._util import f32, i32
from .square import _SensorBase

SQRT3 = 1.7320508075688772
SQRT3_2 = 0.8660254037844386
SQRT3_3 = 0.5773502691896257
_f = np.float32


def _rot(x, y, angle):
    c, s = _f(np.cos(_f(angle))), _f(np.sin(_f(angle)))
    return c * x - s * y, s * x + c * y


def _detect_hex_grid(centers):
    """Hex size, grid rotation (mod 60 deg) and offset from the pixel centres (``hexagonal.py:61-80``)."""
    c = np.asarray(centers, _f)
    n = len(c)
    diff = c[:, None] - c[None, :]
    d2 = np.sum(diff ** 2, axis=2, dtype=_f)
    d2[np.eye(n, dtype=bool)] = np.inf
    idx = int(np.argmin(d2))
    vec = diff[idx // n, idx % n]
    angle = np.mod(np.arctan2(vec[1], vec[0]), _f(np.pi / 3))
    offset = c[int(np.argmin(np.sum(c ** 2, axis=1, dtype=_f)))]
    return _f(np.sqrt(np.min(d2)) / _f(SQRT3)), _f(angle), offset


def _build_lookup_table(centers, hex_size, rotation, offset):
    """Axial-coordinate -> pixel-id table (``hexagonal.py:83-101``)."""
    c = np.asarray(centers, _f)
    xr, yr = _rot(c[:, 0] - offset[0], c[:, 1] - offset[1], -rotation)
    q = (_f(SQRT3_3) * xr - yr / _f(3)) / _f(hex_size)
    r = (_f(2) * yr / _f(3)) / _f(hex_size)
    qi, ri = np.round(q).astype(np.int32), np.round(r).astype(np.int32)
    q_min, r_min = int(qi.min()), int(ri.min())
    table = np.full((int(qi.max()) - q_min + 1, int(ri.max()) - r_min + 1), -1, np.int32)
    table[qi - q_min, ri - r_min] = np.arange(len(c), dtype=np.int32)
    return table, q_min, r_min


class HexagonalSensor(_SensorBase):
    """Hexagonal pixel sensor with hard assignment (``hexagonal.py:104-194``)."""

    kind_code = N.SENSOR_HEX

    def __init__(self, position, rotation, hex_centers, edge_width: float = 0.0, grid: dict | None = None) -> None:
        self.position = f32(position)
        self.rotation = f32(rotation)
        centers = hex_centers.detach().cpu().numpy() if isinstance(hex_centers, torch.Tensor) else hex_centers
        centers = np.asarray(centers, _f)
        self.hex_centers = f32(centers)
        self.n_pixels = len(centers)
        self.edge_width = float(edge_width)
        if grid is None:
            size, rot, offset = _detect_hex_grid(centers)
            self.hex_size = float(size)
            self.hex_inradius = float(_f(size * _f(SQRT3_2)))
            self.grid_rotation = float(rot)
            self.grid_offset = (float(offset[0]), float(offset[1]))
            table, self.q_min, self.r_min = _build_lookup_table(centers, self.hex_size, self.grid_rotation, offset)
        else:
            self.hex_size = float(grid["hex_size"])
            self.hex_inradius = float(grid["hex_inradius"])
            self.grid_rotation = float(grid["grid_rotation"])
            self.grid_offset = (float(grid["grid_offset"][0]), float(grid["grid_offset"][1]))
            table = np.asarray(grid["lookup_table"], np.int32)
            self.q_min, self.r_min = int(grid["q_min"]), int(grid["r_min"])
        self.lookup_table = i32(table)
        # circle about the grid offset that contains every hexagon (circumradius = hex_size), with float32 slack
        d = np.asarray(centers, np.float64) - np.asarray(self.grid_offset, np.float64)
        self.outer_radius = float(np.sqrt((d ** 2).sum(1)).max() + 1.001 * self.hex_size + 1e-5) if len(d) else 0.0

    def grid_constants(self) -> dict:
        """The detected grid statics, in a form accepted back by ``grid=``."""
        return dict(hex_size=self.hex_size, hex_inradius=self.hex_inradius, grid_rotation=self.grid_rotation,
                    grid_offset=self.grid_offset, q_min=self.q_min, r_min=self.r_min,
                    lookup_table=self.lookup_table.cpu().numpy())

    def get_accumulator_shape(self):
        return (self.n_pixels,)

    def _struct(self, keep):
        s = N.IactSensor()
        s.kind = self.kind_code
        self._pose(s)
        s.edge_width = getattr(self, "edge_width", 0.0)
        s.hex_size, s.hex_inradius, s.grid_rotation = self.hex_size, self.hex_inradius, self.grid_rotation
        s.grid_offset[0], s.grid_offset[1] = self.grid_offset
        s.q_min, s.r_min = self.q_min, self.r_min
        s.table_q, s.table_r = int(self.lookup_table.shape[0]), int(self.lookup_table.shape[1])
        s.n_pixels = self.n_pixels
        lut = self.lookup_table.contiguous()
        keep.append(lut)
        s.lookup = N.ptr(lut) if lut.is_cuda else None
        s.sigma = getattr(self, "sigma", 0.0)
        s.kernel_size = getattr(self, "kernel_size", 0)
        s.hex_outer_radius = getattr(self, "outer_radius", 0.0)
        return s


class DifferentiableHexagonalSensor(HexagonalSensor):
    """Hexagonal sensor with Gaussian splatting over ``kernel_size`` neighbour rings
    (``hexagonal.py:197-314``); sigma in units of the hex inradius."""

    kind_code = N.SENSOR_SOFT_HEX

    def __init__(self, position, rotation, hex_centers, sigma: float = 0.5, kernel_size: int = 1,
                 grid: dict | None = None) -> None:
        super().__init__(position, rotation, hex_centers, 0.0, grid)
        self.sigma = float(sigma)
        self.kernel_size = int(kernel_size)
