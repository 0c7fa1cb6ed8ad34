"""Obstruction value types and per-type groups (mirror of reference ``iactrace/core/obstructions.py``).

Groups keep dense float32 tables that the trace kernel stages into shared memory; the ray tests
themselves live in ``csrc/iact_trace.cuh`` (hit_cylinder / hit_box / hit_sphere / hit_obox /
hit_triangle).
"""
from __future__ import annotations

import numpy as np
import torch

from .._util import f32


class Obstruction:
    """Base class for single obstructions."""


class ObstructionGroup:
    """Base class for grouped obstructions."""

    def __len__(self):
        raise NotImplementedError


class Cylinder(Obstruction):
    def __init__(self, p1, p2, radius):
        self.p1 = np.asarray(p1, np.float32)
        self.p2 = np.asarray(p2, np.float32)
        self.radius = float(radius)


class Box(Obstruction):
    def __init__(self, p1, p2):
        self.p1 = np.asarray(p1, np.float32)
        self.p2 = np.asarray(p2, np.float32)


class Sphere(Obstruction):
    def __init__(self, center, radius):
        self.center = np.asarray(center, np.float32)
        self.radius = float(radius)


class OrientedBox(Obstruction):
    def __init__(self, center, half_extents, rotation):
        self.center = np.asarray(center, np.float32)
        self.half_extents = np.asarray(half_extents, np.float32)
        self.rotation = np.asarray(rotation, np.float32).reshape(3, 3)


class Triangle(Obstruction):
    def __init__(self, v0, v1, v2):
        self.v0 = np.asarray(v0, np.float32)
        self.v1 = np.asarray(v1, np.float32)
        self.v2 = np.asarray(v2, np.float32)


def _stack(items, shape):
    return f32(np.stack(items).astype(np.float32)) if len(items) else f32(np.zeros((0,) + shape, np.float32))


class CylinderGroup(ObstructionGroup):
    """p1 (N,3), p2 (N,3), r (N,)."""

    def __init__(self, cylinders=None, p1=None, p2=None, r=None):
        if cylinders is not None:
            p1 = np.stack([c.p1 for c in cylinders])
            p2 = np.stack([c.p2 for c in cylinders])
            r = np.array([c.radius for c in cylinders], np.float32)
        self.p1, self.p2, self.r = f32(p1).reshape(-1, 3), f32(p2).reshape(-1, 3), f32(r).reshape(-1)

    def __len__(self):
        return self.p1.shape[0]


class BoxGroup(ObstructionGroup):
    """p1 (N,3), p2 (N,3) opposite corners of axis-aligned boxes."""

    def __init__(self, boxes=None, p1=None, p2=None):
        if boxes is not None:
            p1 = np.stack([b.p1 for b in boxes])
            p2 = np.stack([b.p2 for b in boxes])
        self.p1, self.p2 = f32(p1).reshape(-1, 3), f32(p2).reshape(-1, 3)

    def __len__(self):
        return self.p1.shape[0]


class SphereGroup(ObstructionGroup):
    """centers (N,3), radii (N,)."""

    def __init__(self, spheres=None, centers=None, radii=None):
        if spheres is not None:
            centers = np.stack([s.center for s in spheres])
            radii = np.array([s.radius for s in spheres], np.float32)
        self.centers, self.radii = f32(centers).reshape(-1, 3), f32(radii).reshape(-1)

    def __len__(self):
        return self.centers.shape[0]


class OrientedBoxGroup(ObstructionGroup):
    """centers (N,3), half_extents (N,3), rotations (N,3,3) local->world."""

    def __init__(self, boxes=None, centers=None, half_extents=None, rotations=None):
        if boxes is not None:
            centers = np.stack([b.center for b in boxes])
            half_extents = np.stack([b.half_extents for b in boxes])
            rotations = np.stack([b.rotation for b in boxes])
        self.centers = f32(centers).reshape(-1, 3)
        self.half_extents = f32(half_extents).reshape(-1, 3)
        self.rotations = f32(rotations).reshape(-1, 3, 3)

    def __len__(self):
        return self.centers.shape[0]


class TriangleGroup(ObstructionGroup):
    """v0, v1, v2 (N,3)."""

    def __init__(self, triangles=None, v0=None, v1=None, v2=None):
        if triangles is not None:
            v0 = np.stack([t.v0 for t in triangles])
            v1 = np.stack([t.v1 for t in triangles])
            v2 = np.stack([t.v2 for t in triangles])
        self.v0, self.v1, self.v2 = f32(v0).reshape(-1, 3), f32(v1).reshape(-1, 3), f32(v2).reshape(-1, 3)

    def __len__(self):
        return self.v0.shape[0]


def group_obstructions(obstructions):
    """List of obstructions -> list of groups in the reference's fixed type order
    (cylinder, box, sphere, oriented box, triangle; ``obstructions.py:258-278``)."""
    groups = []
    for cls, gcls in ((Cylinder, CylinderGroup), (Box, BoxGroup), (Sphere, SphereGroup),
                      (OrientedBox, OrientedBoxGroup), (Triangle, TriangleGroup)):
        items = [o for o in obstructions if type(o) is cls]
        if items:
            groups.append(gcls(items))
    return groups
