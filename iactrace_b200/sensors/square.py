"""Square-pixel sensors (mirror of reference ``iactrace/sensors/square.py``).

The binning arithmetic runs in the trace kernel (``square_pixel`` / ``splat_soft_square`` in
``csrc``); these classes hold the statics and expose ``accumulate`` through ``iact_accumulate``.
"""
from __future__ import annotations

import torch

from .. import _native as N
from .._util import f32


class _SensorBase:
    kind_code = -1

    def get_accumulator_shape(self):
        raise NotImplementedError

    def _struct(self, keep) -> "N.IactSensor":
        raise NotImplementedError

    def _pose(self, s):
        # host copy of the pose, cached per tensor version: reading it back on every render would
        # synchronise the stream
        sig = (self.position.data_ptr(), self.position._version, self.rotation.data_ptr(), self.rotation._version)
        cached = self.__dict__.get("_pose_cache")
        if cached is None or cached[0] != sig:
            cached = (sig, self.position.detach().cpu().tolist(), self.rotation.detach().cpu().tolist())
            self.__dict__["_pose_cache"] = cached
        _, pos, rot = cached
        for i in range(3):
            s.position[i] = pos[i]
            s.euler[i] = rot[i]

    def accumulate(self, x, y, values) -> torch.Tensor:
        """Bin free-standing hits ``(x, y, values)`` into pixels on the GPU."""
        N.require_cuda()
        x, y, values = f32(x).reshape(-1).contiguous(), f32(y).reshape(-1).contiguous(), f32(values).reshape(-1).contiguous()
        out = torch.empty(self.get_accumulator_shape(), dtype=torch.float32, device=x.device)
        keep = []
        s = self._struct(keep)
        N.check(N.lib().iact_accumulate(s, N.ptr(x), N.ptr(y), N.ptr(values), x.numel(), N.ptr(out), N.stream_ptr()),
                "accumulate")
        return out


class SquareSensor(_SensorBase):
    """Square pixel sensor with hard assignment (``square.py:28-91``)."""

    kind_code = N.SENSOR_SQUARE

    def __init__(self, position, rotation, width, height, bounds, edge_width: float = 0.0) -> None:
        self.position = f32(position)
        self.rotation = f32(rotation)
        self.width = int(width)
        self.height = int(height)
        self.edge_width = float(edge_width)
        xmin, xmax, ymin, ymax = bounds
        self.x0 = float(xmin)
        self.y0 = float(ymin)
        self.dx = float((xmax - xmin) / width)
        self.dy = float((ymax - ymin) / height)

    def get_accumulator_shape(self):
        return (self.height, self.width)

    def _struct(self, keep):
        s = N.IactSensor()
        s.kind = self.kind_code
        self._pose(s)
        s.width, s.height = self.width, self.height
        s.x0, s.y0, s.dx, s.dy = self.x0, self.y0, self.dx, self.dy
        s.edge_width = getattr(self, "edge_width", 0.0)
        s.sigma = getattr(self, "sigma", 0.0)
        s.kernel_size = getattr(self, "kernel_size", 0)
        return s


class DifferentiableSquareSensor(SquareSensor):
    """Square sensor with Gaussian soft splatting over (2k+1)^2 pixels (``square.py:94-172``)."""

    kind_code = N.SENSOR_SOFT_SQUARE

    def __init__(self, position, rotation, width, height, bounds=(-1, 1, -1, 1), sigma: float = 0.1,
                 kernel_size: int = 2) -> None:
        super().__init__(position, rotation, width, height, bounds, 0.0)
        self.sigma = float(sigma)
        self.kernel_size = int(kernel_size)
