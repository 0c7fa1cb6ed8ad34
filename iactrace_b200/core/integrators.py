"""Monte-Carlo facet sampling (mirror of reference ``iactrace/core/integrators.py``).

``MCIntegrator.sample_group`` launches the CUDA sampler (``csrc/iact_sample.cu``), which evaluates
the reference's JAX threefry key tree with random access: one thread per (facet, sample).
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch

from .. import _native as N
from .. import random as R
from .._util import replace, contig


class Integrator(ABC):
    """Abstract mirror-sampling integrator (``integrators.py:18-55``)."""

    @abstractmethod
    def sample_group(self, group, key):
        ...

    def sample_mirror_groups(self, mirror_groups, key):
        if not mirror_groups:
            return []
        keys = R.split(key, len(mirror_groups) + 1)
        return [self.sample_group(g, k) for g, k in zip(mirror_groups, keys[:-1])]


class MCIntegrator(Integrator):
    """Monte-Carlo integrator: ``n_samples`` uniform aperture points per facet (``integrators.py:58-188``)."""

    def __init__(self, n_samples: int = 128) -> None:
        self.n_samples = int(n_samples)

    def sample_group(self, group, key):
        params = group.get_sampling_params()
        gtype = params["type"]
        if gtype not in ("disk", "polygon"):
            raise TypeError(f"Unknown MirrorGroup type: {gtype}")
        N.require_cuda()
        key = R.as_key(key)
        F, M = len(group), self.n_samples
        dev = group.positions.device
        pts = torch.empty((F, M, 3), dtype=torch.float32, device=dev)
        nrm = torch.empty((F, M, 3), dtype=torch.float32, device=dev)
        dlt = torch.empty((F, M, 3), dtype=torch.float32, device=dev)
        wts = torch.empty((F, M, 1), dtype=torch.float32, device=dev)
        surf = N.IactSurface(group.curvature, group.conic, len(group.aspheric))
        if len(group.aspheric) > N.MAX_ASPH:
            raise ValueError(f"at most {N.MAX_ASPH} aspheric terms are supported")
        for i, a in enumerate(group.aspheric.tolist()):
            surf.aspheric[i] = a
        offs = contig(group.offsets.detach())
        if gtype == "disk":
            radii = contig(group.radii.detach())
            rc = N.lib().iact_sample_disk_group(N.key_arg(key), R.mode_code(), F, M, surf, N.ptr(radii), N.ptr(offs),
                                                N.ptr(pts), N.ptr(nrm), N.ptr(dlt), N.ptr(wts), N.stream_ptr())
        else:
            verts = contig(group.vertices.detach())
            rc = N.lib().iact_sample_polygon_group(N.key_arg(key), R.mode_code(), F, M, surf, group.n_vertices,
                                                   N.ptr(verts), N.ptr(offs), N.ptr(pts), N.ptr(nrm), N.ptr(dlt),
                                                   N.ptr(wts), N.stream_ptr())
        N.check(rc, "sample_group")
        return replace(group, points=pts, normals=nrm, perturbation_delta=dlt, weights=wts)
