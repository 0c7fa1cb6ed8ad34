"""Mirror value type and grouped structure-of-arrays mirror tables
(mirror of reference ``iactrace/telescope/mirrors.py``).

A group holds per-facet parameters and the Monte-Carlo sample tables as float32 tensors on the
GPU.  ``positions``, ``rotations``, ``perturbation_scale`` and ``weights`` may be autograd leaves:
``render`` differentiates through them with the hand-written VJP kernel.
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np
import torch

from .. import _native as N
from .._util import f32, contig
from ..core.apertures import DiskAperture, PolygonAperture
from ..core.surfaces import AsphericSurface


class Mirror:
    """Single mirror element (``mirrors.py:10-35``)."""

    def __init__(self, position, rotation, surface, aperture, points=None, normals=None, weights=None,
                 optical_stage=0, offset=None):
        self.position = np.asarray(_host(position), np.float32)
        self.rotation = np.asarray(_host(rotation), np.float32)
        self.surface = surface
        self.aperture = aperture
        self.offset = np.asarray(_host(offset), np.float32) if offset is not None else np.zeros(2, np.float32)
        self.points = points if points is not None else np.zeros((0, 3), np.float32)
        self.normals = normals if normals is not None else np.zeros((0, 3), np.float32)
        self.weights = weights if weights is not None else np.zeros((0, 1), np.float32)
        self.optical_stage = int(optical_stage)


def _host(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x


class MirrorGroup:
    """Grouped mirrors sharing surface and aperture type (``mirrors.py:46-99``)."""

    kind = "?"

    def _init_common(self, positions, rotations, surface, optical_stage, offsets):
        self.positions = f32(positions).reshape(-1, 3)
        self.rotations = f32(rotations).reshape(-1, 3)
        self.curvature = surface.curvature
        self.conic = surface.conic
        self.aspheric = surface.aspheric
        self.optical_stage = int(optical_stage)
        n = self.positions.shape[0]
        self.offsets = f32(offsets).reshape(-1, 2) if offsets is not None else f32(np.zeros((n, 2), np.float32))
        dev = self.positions.device
        self.points = torch.zeros((n, 0, 3), device=dev)
        self.normals = torch.zeros((n, 0, 3), device=dev)
        self.weights = torch.zeros((n, 0, 1), device=dev)
        self.perturbation_delta = torch.zeros((n, 0, 3), device=dev)
        self.perturbation_scale = torch.zeros(n, device=dev)
        self.sample_stream = None        # core.integrators.SampleStream when the samples are regenerated on demand

    def __len__(self):
        return self.positions.shape[0]

    def get_surface(self) -> AsphericSurface:
        return AsphericSurface(self.curvature, self.conic, self.aspheric)

    def _facets_struct(self, keep):
        """IactFacets view of this group's tensors (keeps references alive in ``keep``)."""
        t = [contig(x.detach()) for x in (self.positions, self.rotations, self.perturbation_scale, self.points,
                                          self.normals, self.perturbation_delta, self.weights)]
        keep.extend(t)
        return N.IactFacets(len(self), self.points.shape[1], *[N.ptr(x) for x in t])

    def transform_to_world(self):
        """World-space sample points, perturbed unit normals and weights: (N,M,3),(N,M,3),(N,M,1)
        (``mirrors.py:64-79``), evaluated by the ``transform_kernel``."""
        N.require_cuda()
        n, m = len(self), self.points.shape[1]
        world = torch.empty((n, m, 8), dtype=torch.float32, device=self.points.device)
        bounds = torch.empty((n, 4), dtype=torch.float32, device=self.points.device)
        if n * m:
            keep = []
            fa = self._facets_struct(keep)
            N.check(N.lib().iact_transform_to_world(fa, 0, N.ptr(world), N.ptr(bounds), N.stream_ptr()),
                    "transform_to_world")
        return world[..., 0:3], world[..., 4:7], self.weights


class AsphericDiskMirrorGroup(MirrorGroup):
    """Aspheric mirrors with circular apertures (``mirrors.py:101-158``)."""

    kind = "disk"

    def __init__(self, positions, rotations, surface, radii, optical_stage=0, offsets=None):
        self._init_common(positions, rotations, surface, optical_stage, offsets)
        self.radii = f32(radii).reshape(-1)

    def check_aperture(self, x, y, mirror_idx):
        return x ** 2 + y ** 2 <= self.radii[mirror_idx] ** 2

    def get_sampling_params(self):
        return {"type": "disk", "radii": self.radii, "offsets": self.offsets, "surface": self.get_surface()}


class AsphericPolygonMirrorGroup(MirrorGroup):
    """Aspheric mirrors with convex polygon apertures of equal vertex count (``mirrors.py:161-229``)."""

    kind = "polygon"

    def __init__(self, positions, rotations, surface, vertices_list, optical_stage=0, offsets=None):
        self._init_common(positions, rotations, surface, optical_stage, offsets)
        self.vertices = f32(vertices_list)
        self.n_vertices = int(self.vertices.shape[1])

    def check_aperture(self, x, y, mirror_idx):
        verts = self.vertices[mirror_idx]
        n = self.n_vertices
        inside = None
        for i in range(n):
            v1, v2 = verts[i], verts[(i + 1) % n]
            cross = (v2[0] - v1[0]) * (y - v1[1]) - (v2[1] - v1[1]) * (x - v1[0])
            inside = (cross >= 0) if inside is None else inside & (cross >= 0)
        return inside

    def get_sampling_params(self):
        return {"type": "polygon", "vertices": self.vertices, "offsets": self.offsets, "surface": self.get_surface()}


def _group_by_surface_params(mirrors):
    """Insertion-ordered grouping by (curvature, conic, aspheric) (``mirrors.py:316-336``)."""
    grouped = defaultdict(list)
    for m in mirrors:
        if isinstance(m.surface, AsphericSurface):
            grouped[m.surface.params_key()].append(m)
    return grouped


def group_mirrors(mirrors):
    """Mirrors -> groups: stage ascending, then disk groups by surface, then polygon groups by
    vertex count and surface (``mirrors.py:232-313``).  Facet order inside a group is input order,
    which fixes the PRNG key each facet receives."""
    if not mirrors:
        return []
    groups = []
    by_stage = defaultdict(list)
    for m in mirrors:
        by_stage[m.optical_stage].append(m)
    for stage, sm in sorted(by_stage.items()):
        disk = [m for m in sm if isinstance(m.surface, AsphericSurface) and isinstance(m.aperture, DiskAperture)]
        for _, ml in _group_by_surface_params(disk).items():
            groups.append(AsphericDiskMirrorGroup(
                np.stack([m.position for m in ml]), np.stack([m.rotation for m in ml]), ml[0].surface,
                np.array([m.aperture.radius for m in ml], np.float32), optical_stage=stage,
                offsets=np.stack([m.offset for m in ml])))
        poly = [m for m in sm if isinstance(m.surface, AsphericSurface) and isinstance(m.aperture, PolygonAperture)]
        by_nv = defaultdict(list)
        for m in poly:
            by_nv[len(m.aperture.vertices)].append(m)
        for _, ml in by_nv.items():
            for _, mll in _group_by_surface_params(ml).items():
                groups.append(AsphericPolygonMirrorGroup(
                    np.stack([m.position for m in mll]), np.stack([m.rotation for m in mll]), mll[0].surface,
                    np.stack([m.aperture.vertices for m in mll]), optical_stage=stage,
                    offsets=np.stack([m.offset for m in mll])))
    return groups
