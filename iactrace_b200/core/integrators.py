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


class SampleStream:
    """A facet-sample stream that is NOT held in memory: the key, key-derivation mode and length of the
    ``MCIntegrator(n_samples)`` draw of one mirror group.  Threefry is counter based, so any window of the stream is
    regenerated bit-identically on demand (``MCIntegrator.sample_rows``); ``render`` walks such a group in L2-sized
    windows (``core/streaming.py``)."""

    def __init__(self, key, mode: str, n_samples: int) -> None:
        self.key, self.mode, self.n_samples = R.as_key(key).copy(), mode, int(n_samples)


class MCIntegrator(Integrator):
    """Monte-Carlo integrator: ``n_samples`` uniform aperture points per facet (``integrators.py:58-188``).

    ``stream``: None (default) keeps the sample tables in device memory unless they would exceed
    ``config.stream_samples_bytes`` (then only the key is kept and ``render`` regenerates the samples window by
    window); True / False force one or the other.  The random numbers are the same either way."""

    def __init__(self, n_samples: int = 128, stream: bool | None = None) -> None:
        self.n_samples = int(n_samples)
        self.stream = stream

    def sample_group(self, group, key):
        from .. import config
        params = group.get_sampling_params()
        if params["type"] not in ("disk", "polygon"):
            raise TypeError(f"Unknown MirrorGroup type: {params['type']}")
        stream = self.stream
        if stream is None:
            stream = len(group) * self.n_samples * 72 > config.stream_samples_bytes     # 40 B local + 32 B world per sample
        if stream:
            import torch as _t
            dev, n = group.positions.device, len(group)
            empty = lambda k: _t.zeros((n, 0, k), dtype=_t.float32, device=dev)
            return replace(group, points=empty(3), normals=empty(3), perturbation_delta=empty(3), weights=empty(1),
                           sample_stream=SampleStream(key, R.get_rng_mode(), self.n_samples))
        return replace(self.sample_rows(group, key, 0, self.n_samples, self.n_samples), sample_stream=None)

    @staticmethod
    def sample_rows(group, key, first: int, n_rows: int, n_total: int, mode: str | None = None):
        """Samples ``first .. first + n_rows - 1`` of the ``n_total``-sample stream of ``key`` as a group with
        materialised (F, n_rows, .) tables (``iact_sample_*_group_rows``)."""
        params = group.get_sampling_params()
        gtype = params["type"]
        if gtype not in ("disk", "polygon"):
            raise TypeError(f"Unknown MirrorGroup type: {gtype}")
        N.require_cuda()
        key = R.as_key(key)
        F, M = len(group), int(n_rows)
        dev = group.positions.device
        pts = torch.empty((F, M, 3), dtype=torch.float32, device=dev)
        nrm = torch.empty((F, M, 3), dtype=torch.float32, device=dev)
        dlt = torch.empty((F, M, 3), dtype=torch.float32, device=dev)
        wts = torch.empty((F, M, 1), dtype=torch.float32, device=dev)
        surf = N.IactSurface(float(group.curvature), float(group.conic), len(group.aspheric))
        if len(group.aspheric) > N.MAX_ASPH:
            raise ValueError(f"at most {N.MAX_ASPH} aspheric terms are supported")
        for i, a in enumerate(group.aspheric.tolist()):
            surf.aspheric[i] = a
        offs = contig(group.offsets.detach())
        with torch.cuda.device(dev):
            if gtype == "disk":
                radii = contig(group.radii.detach())
                rc = N.lib().iact_sample_disk_group_rows(N.key_arg(key), R.mode_code(mode), F, M, int(first), int(n_total), surf,
                                                         N.ptr(radii), N.ptr(offs), N.ptr(pts), N.ptr(nrm), N.ptr(dlt),
                                                         N.ptr(wts), N.stream_ptr())
            else:
                verts = contig(group.vertices.detach())
                rc = N.lib().iact_sample_polygon_group_rows(N.key_arg(key), R.mode_code(mode), F, M, int(first), int(n_total),
                                                            surf, group.n_vertices, N.ptr(verts), N.ptr(offs), N.ptr(pts),
                                                            N.ptr(nrm), N.ptr(dlt), N.ptr(wts), N.stream_ptr())
        N.check(rc, "sample_group")
        return replace(group, points=pts, normals=nrm, perturbation_delta=dlt, weights=wts)
