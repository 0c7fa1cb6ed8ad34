"""Mirror of ``iactrace.core`` (reference ``iactrace/core/__init__.py``)."""
from .intersections import (intersect_plane, intersect_cylinder, intersect_box, intersect_sphere,
                            intersect_oriented_box, intersect_triangle, intersect_conic)
from .surfaces import AsphericSurface
from .apertures import Aperture, DiskAperture, PolygonAperture
from .integrators import Integrator, MCIntegrator
from .reflection import reflect
from .transforms import euler_to_matrix, look_at_rotation
from .render import render, render_debug, render_response_matrix
from .obstructions import (
    Obstruction, ObstructionGroup, Cylinder, CylinderGroup, Box, BoxGroup, Sphere, SphereGroup,
    OrientedBox, OrientedBoxGroup, Triangle, TriangleGroup, group_obstructions,
)

__all__ = [
    "intersect_plane", "intersect_cylinder", "intersect_box", "intersect_sphere", "intersect_oriented_box",
    "intersect_triangle", "intersect_conic",
    "AsphericSurface", "Aperture", "DiskAperture", "PolygonAperture", "Integrator", "MCIntegrator",
    "reflect", "euler_to_matrix", "look_at_rotation", "render", "render_debug", "render_response_matrix",
    "Obstruction", "ObstructionGroup", "Cylinder", "CylinderGroup", "Box", "BoxGroup", "Sphere", "SphereGroup",
    "OrientedBox", "OrientedBoxGroup", "Triangle", "TriangleGroup", "group_obstructions",
]
