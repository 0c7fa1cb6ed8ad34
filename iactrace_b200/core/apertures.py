"""Aperture value types (mirror of reference ``iactrace/core/apertures.py:8-62``)."""
from __future__ import annotations

import math

import numpy as np


class Aperture:
    """Abstract aperture shape."""

    def area(self):
        raise NotImplementedError

    def check_aperture(self, x, y):
        raise NotImplementedError


class DiskAperture(Aperture):
    """Circular aperture."""

    def __init__(self, radius: float = 1.0):
        self.radius = float(radius)

    def area(self) -> float:
        return math.pi * self.radius ** 2

    def check_aperture(self, x, y):
        return x ** 2 + y ** 2 <= self.radius ** 2


class PolygonAperture(Aperture):
    """Convex polygonal aperture (vertices in order)."""

    def __init__(self, vertices):
        self.vertices = np.asarray(vertices, dtype=np.float32)
        self.n_vertices = len(self.vertices)

    def area(self) -> float:
        x, y = self.vertices[:, 0], self.vertices[:, 1]
        return float(0.5 * abs(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y)))

    def check_aperture(self, x, y):
        inside = None
        n = self.n_vertices
        for i in range(n):
            v1, v2 = self.vertices[i], self.vertices[(i + 1) % n]
            cross = (v2[0] - v1[0]) * (y - v1[1]) - (v2[1] - v1[1]) * (x - v1[0])
            ok = cross >= 0
            inside = ok if inside is None else inside & ok
        return inside
