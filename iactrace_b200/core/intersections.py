"""Ray-primitive intersections as batched torch functions (API mirror of reference
``iactrace/core/intersections.py``: same names, argument order and eps conventions).

Inside ``render`` these tests run in the CUDA trace kernel (``csrc/iact_trace.cuh``); the functions
here are the stand-alone equivalents the reference exports from ``iactrace.core``.  They broadcast
over leading dimensions and run on whatever device their inputs live on.
"""
from __future__ import annotations

import torch

from .._util import f32

INF = float("inf")


def _dot(a, b):
    return (a * b).sum(-1)


def intersect_plane(ray_origin, ray_direction, plane_center, plane_rotation):
    """2-D coordinates of the hit on a plane given by centre and rotation matrix (z-axis = normal);
    parallel rays and hits behind the origin give (1e10, 1e10)."""
    o, d, c, R = f32(ray_origin), f32(ray_direction), f32(plane_center), f32(plane_rotation)
    u1, u2, n = R[:, 0], R[:, 1], R[:, 2]
    ndotd, ndoto, ndotp = _dot(d, n), _dot(o, n), (n * c).sum()
    parallel = ndotd.abs() < 1e-10
    t = (ndotp - ndoto) / torch.where(parallel, torch.ones_like(ndotd), ndotd)
    op = o + t[..., None] * d - c
    invalid = parallel | (t <= 0)
    big = torch.full_like(t, 1e10)
    return torch.stack([torch.where(invalid, big, _dot(op, u1)), torch.where(invalid, big, _dot(op, u2))], -1)


def intersect_cylinder(ray_origin, ray_direction, p1, p2, radius):
    """Nearest positive ray parameter on a capped cylinder, inf if none."""
    o, d, p1, p2 = f32(ray_origin), f32(ray_direction), f32(p1), f32(p2)
    radius = f32(radius)
    axis = p2 - p1
    height = torch.linalg.norm(axis, dim=-1)
    axis = axis / height[..., None]
    oc = o - p1
    oc_ax, rd_ax = _dot(oc, axis), _dot(d, axis)
    ocp, rdp = oc - oc_ax[..., None] * axis, d - rd_ax[..., None] * axis
    a, b, c = _dot(rdp, rdp), 2 * _dot(ocp, rdp), _dot(ocp, ocp) - radius * radius
    disc = b * b - 4 * a * c
    eps = 1e-8
    sq = torch.sqrt(torch.clamp(disc, min=0.0))
    t1, t2 = (-b - sq) / (2 * a + eps), (-b + sq) / (2 * a + eps)
    y1, y2 = oc_ax + t1 * rd_ax, oc_ax + t2 * rd_ax
    inf = torch.full_like(t1, INF)
    t1 = torch.where((t1 > eps) & (y1 >= 0) & (y1 <= height) & (disc >= 0), t1, inf)
    t2 = torch.where((t2 > eps) & (y2 >= 0) & (y2 <= height) & (disc >= 0), t2, inf)
    tb, tt = -oc_ax / (rd_ax + eps), (height - oc_ax) / (rd_ax + eps)
    pb, pt = ocp + tb[..., None] * rdp, ocp + tt[..., None] * rdp
    tb = torch.where((tb > eps) & (_dot(pb, pb) <= radius ** 2), tb, inf)
    tt = torch.where((tt > eps) & (_dot(pt, pt) <= radius ** 2), tt, inf)
    return torch.minimum(torch.minimum(t1, t2), torch.minimum(tb, tt))


def _slab(t1, t2, eps):
    tmin = torch.minimum(t1, t2).max(-1).values
    tmax = torch.maximum(t1, t2).min(-1).values
    hit = (tmax >= tmin) & (tmax > eps)
    tr = torch.where(tmin > eps, tmin, tmax)
    return hit, tr


def intersect_box(ray_origin, ray_direction, p1, p2):
    """Axis-aligned box given by two opposite corners."""
    o, d, p1, p2 = f32(ray_origin), f32(ray_direction), f32(p1), f32(p2)
    eps = 1e-8
    inv = 1.0 / (d + eps)
    hit, tr = _slab((torch.minimum(p1, p2) - o) * inv, (torch.maximum(p1, p2) - o) * inv, eps)
    return torch.where(hit, tr, torch.full_like(tr, INF))


def intersect_oriented_box(ray_origin, ray_direction, center, half_extents, rotation):
    """Oriented box: centre, half sizes and local->world rotation matrix."""
    o, d, c, h, R = f32(ray_origin), f32(ray_direction), f32(center), f32(half_extents), f32(rotation)
    eps = 1e-8
    lo = torch.einsum("...ji,...j->...i", R, o - c)
    ld = torch.einsum("...ji,...j->...i", R, d)
    inv = 1.0 / (ld + eps * torch.sign(ld + eps))
    hit, tr = _slab((-h - lo) * inv, (h - lo) * inv, eps)
    return torch.where(hit & (tr > eps), tr, torch.full_like(tr, INF))


def intersect_triangle(ray_origin, ray_direction, v0, v1, v2):
    """Moeller-Trumbore."""
    o, d, v0, v1, v2 = f32(ray_origin), f32(ray_direction), f32(v0), f32(v1), f32(v2)
    eps = 1e-8
    e1, e2 = v1 - v0, v2 - v0
    d, e2b = torch.broadcast_tensors(d, e2)
    h = torch.linalg.cross(d, e2b)
    a = _dot(e1, h)
    f = 1.0 / (a + eps * torch.sign(a + eps))
    s = o - v0
    u = f * _dot(s, h)
    q = torch.linalg.cross(*torch.broadcast_tensors(s, e1))
    v = f * _dot(d, q)
    t = f * _dot(e2, q)
    valid = (a.abs() >= eps) & (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t > eps)
    return torch.where(valid, t, torch.full_like(t, INF))


def intersect_sphere(ray_origin, ray_direction, center, radius):
    o, d, c, r = f32(ray_origin), f32(ray_direction), f32(center), f32(radius)
    eps = 1e-8
    oc = o - c
    a, b, cc = _dot(d, d), 2.0 * _dot(oc, d), _dot(oc, oc) - r * r
    disc = b * b - 4.0 * a * cc
    sq = torch.sqrt(torch.clamp(disc, min=0.0))
    t1, t2 = (-b - sq) / (2.0 * a + eps), (-b + sq) / (2.0 * a + eps)
    inf = torch.full_like(t1, INF)
    return torch.minimum(torch.where((t1 > eps) & (disc >= 0), t1, inf), torch.where((t2 > eps) & (disc >= 0), t2, inf))


def intersect_conic(ray_origin, ray_direction, curvature, conic):
    """Closed-form root of c (x^2 + y^2) + (1+k) c z^2 - 2 z = 0 along the ray (smallest t > 1e-8)."""
    o, d = f32(ray_origin), f32(ray_direction)
    ox, oy, oz, dx, dy, dz = o[..., 0], o[..., 1], o[..., 2], d[..., 0], d[..., 1], d[..., 2]
    c, k1 = float(curvature), 1.0 + float(conic)
    inf = torch.full_like(ox, INF)
    if abs(c) < 1e-12:
        return torch.where(dz.abs() > 1e-10, -oz / dz, inf)
    A = c * (dx * dx + dy * dy + k1 * dz * dz)
    B = 2 * (c * (ox * dx + oy * dy + k1 * oz * dz) - dz)
    C = c * (ox * ox + oy * oy + k1 * oz * oz) - 2 * oz
    disc = B * B - 4 * A * C
    sq = torch.sqrt(torch.clamp(disc, min=0.0))
    t1, t2 = (-B - sq) / (2 * A + 1e-30), (-B + sq) / (2 * A + 1e-30)
    v1, v2 = t1 > 1e-8, t2 > 1e-8
    tc = torch.where(v1 & v2, torch.minimum(t1, t2), torch.where(v1, t1, torch.where(v2, t2, inf)))
    return torch.where(disc < 0, inf, tc)
