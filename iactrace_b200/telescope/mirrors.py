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
from .._util import f32, contig, replace
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

    def with_surface(self, curvature=None, conic=None, offsets=None):
        """This group with other surface parameters (reference ``core/surfaces.py:25-65``); tensors that require grad
        make ``render`` differentiable w.r.t. them.

        Stage 0: the sample tables are rebuilt as torch functions of (curvature, conic, offsets) -- surface point
        ``(x, y, sag(x + x0, y + y0) - sag(x0, y0))``, unit normal from the sag gradient, tangent-space perturbation
        (``reflection.py:22-49``) and weight ``n_z / area * M`` (``integrators.py:125-127``) -- with the drawn aperture
        coordinates (x, y) and tangent angles held fixed: what ``jax.grad`` through ``MCIntegrator.sample_group`` gives.
        Stages >= 1 use their surface inside ``render`` (``surfaces.py:67-107``); there the parameters are handed to
        the kernel, whose VJP differentiates the Newton root implicitly."""
        c = self.curvature if curvature is None else curvature
        k = self.conic if conic is None else conic
        off = self.offsets if offsets is None else f32(offsets).reshape(-1, 2)
        if self.optical_stage != 0 or self.points.shape[1] == 0:
            if getattr(self, "sample_stream", None) is not None and self.optical_stage == 0:
                raise NotImplementedError("with_surface needs materialised samples: MCIntegrator(n, stream=False)")
            return replace(self, curvature=c, conic=k, offsets=off)
        dev = self.points.device
        tt = lambda v: v.to(dev) if isinstance(v, torch.Tensor) else torch.tensor(float(v), dtype=torch.float32, device=dev)
        ct, kt = tt(c), tt(k)
        asph = [float(a) for a in (self.aspheric.tolist() if hasattr(self.aspheric, "tolist") else self.aspheric)]

        def sag(x, y):
            r2 = x * x + y * y
            z = r2 * ct / (1 + torch.sqrt(1 - (1 + kt) * ct * ct * r2))
            for i, a in enumerate(asph):
                z = z + a * r2 ** (2 * i + 2)
            return z

        def dsag_dr2(r2):
            f1 = 0.5 * ct / torch.sqrt(1 - (1 + kt) * ct * ct * r2)
            for i, a in enumerate(asph):
                f1 = f1 + a * (2 * i + 2) * r2 ** (2 * i + 1)
            return f1

        old_n, old_d = self.normals.detach(), self.perturbation_delta.detach()
        x, y = self.points.detach()[..., 0], self.points.detach()[..., 1]
        x0, y0 = off[:, None, 0], off[:, None, 1]
        X, Y = x + x0, y + y0
        pts = torch.stack([x, y, sag(X, Y) - sag(x0, y0)], dim=-1)
        f1 = dsag_dr2(X * X + Y * Y)
        m = torch.stack([-2 * X * f1, -2 * Y * f1, torch.ones_like(X)], dim=-1)
        nrm = m / m.norm(dim=-1, keepdim=True)

        def tangents(n):
            ref = torch.where((n[..., 2:3].abs() > 0.9), torch.tensor([1.0, 0.0, 0.0], device=dev), torch.tensor([0.0, 0.0, 1.0], device=dev))
            t1 = torch.cross(n, ref.expand_as(n), dim=-1)
            t1 = t1 / t1.norm(dim=-1, keepdim=True)
            return t1, torch.cross(n, t1, dim=-1)

        t1o, t2o = tangents(old_n)
        th1, th2 = (old_d * t1o).sum(-1, keepdim=True), (old_d * t2o).sum(-1, keepdim=True)     # the drawn N(0,1) angles
        t1, t2 = tangents(nrm)
        dlt = th1 * t1 + th2 * t2
        wts = self.weights * (nrm[..., 2:3] / old_n[..., 2:3])
        return replace(self, curvature=c, conic=k, offsets=off, points=pts, normals=nrm, perturbation_delta=dlt, weights=wts)

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
