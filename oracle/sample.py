"""Oracle restatement of the reference's Monte-Carlo facet sampling.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Follows, function by function:
* ``iactrace/utils/sampling.py:8-27``   sample_disk
* ``iactrace/utils/sampling.py:30-67``  sample_polygon
* ``iactrace/core/surfaces.py:25-65``   AsphericSurface._sag_raw / sag / normal / point_and_normal
* ``iactrace/core/reflection.py:22-49`` compute_perturbation_delta
* ``iactrace/core/integrators.py:97-188`` MCIntegrator._sample_disk_group / _sample_polygon_group

``dt`` selects the arithmetic type of the geometry (np.float32 = reference
semantics; np.float64 = "exact" evaluation from the same f32 random numbers).
"""
from __future__ import annotations

import numpy as np

from . import prng


def sag_raw(x, y, c, k, asph, dt=np.float32):
    """surfaces.py:25-39.  ``c``/``k`` are Python floats at sampling time, so the
    constant ``(1+k)*c*c`` folds in double precision before meeting the array."""
    x = np.asarray(x, dtype=dt)
    y = np.asarray(y, dtype=dt)
    r2 = x * x + y * y
    with np.errstate(invalid="ignore"):
        denom = dt(1) + np.sqrt(dt(1) - dt((1 + k) * c * c) * r2)
    z = r2 * dt(c) / denom
    asph = np.asarray(asph, dtype=dt)
    if asph.size > 0:
        powers = np.arange(2, 2 + 2 * len(asph), 2)
        z = z + np.sum(asph * r2[..., None] ** powers.astype(dt), axis=-1)
    return z.astype(dt)


def dsag_raw(x, y, c, k, asph, dt=np.float32):
    """(dz/dx, dz/dy) of ``sag_raw`` in the form reverse-mode autodiff produces
    (surfaces.py:51-58 uses ``jax.grad`` of ``_sag_raw``)."""
    x = np.asarray(x, dtype=dt)
    y = np.asarray(y, dtype=dt)
    r2 = x * x + y * y
    kc2 = dt((1 + k) * c * c)
    with np.errstate(invalid="ignore", divide="ignore"):
        s = np.sqrt(dt(1) - kc2 * r2)
        d = dt(1) + s
        # d(n/d) = dn/d - n/d^2 * dd ;  dd/dr2 = (0.5/s) * (-kc2)
        dz_dr2 = dt(c) / d - (r2 * dt(c)) / (d * d) * ((dt(0.5) / s) * (-kc2))
    asph = np.asarray(asph, dtype=dt)
    if asph.size > 0:
        powers = np.arange(2, 2 + 2 * len(asph), 2).astype(dt)
        dz_dr2 = dz_dr2 + np.sum(asph * powers * r2[..., None] ** (powers - dt(1)), axis=-1)
    return (dz_dr2 * (x + x)).astype(dt), (dz_dr2 * (y + y)).astype(dt)


def sag(x, y, offset, c, k, asph, dt=np.float32):
    """surfaces.py:41-45."""
    x0, y0 = dt(offset[0]), dt(offset[1])
    z0 = sag_raw(x0, y0, c, k, asph, dt)
    return sag_raw(np.asarray(x, dt) + x0, np.asarray(y, dt) + y0, c, k, asph, dt) - z0


def point_and_normal(xy, offset, c, k, asph, dt=np.float32):
    """surfaces.py:47-65."""
    x = np.asarray(xy[..., 0], dtype=dt)
    y = np.asarray(xy[..., 1], dtype=dt)
    z = sag(x, y, offset, c, k, asph, dt)
    pts = np.stack([x, y, z], axis=-1)
    dzdx, dzdy = dsag_raw(x + dt(offset[0]), y + dt(offset[1]), c, k, asph, dt)
    n = np.stack([-dzdx, -dzdy, np.ones_like(dzdx)], axis=-1)
    n = n / np.sqrt(np.sum(n * n, axis=-1, keepdims=True))
    return pts.astype(dt), n.astype(dt)


def sample_disk(k, n, mode=prng.PARTITIONABLE):
    """sampling.py:8-27 -> (n, 2) float32 in the unit disk."""
    k1, k2 = prng.split(k, 2, mode)
    r = np.sqrt(prng.uniform(k1, n, mode=mode))
    theta = prng.uniform(k2, n, mode=mode) * np.float32(2) * np.float32(np.pi)
    return np.stack([r * np.cos(theta), r * np.sin(theta)], axis=-1).astype(np.float32)


def sample_polygon(k, verts, n, mode=prng.PARTITIONABLE):
    """sampling.py:30-67 -> (n, 2) float32 (fan triangulation from vertex 0)."""
    verts = np.asarray(verts, dtype=np.float32)
    nv = len(verts)
    tris = np.stack([np.stack([verts[0], verts[i], verts[i + 1]]) for i in range(1, nv - 1)])
    v0, v1, v2 = tris[:, 0], tris[:, 1], tris[:, 2]
    areas = np.abs((v1[:, 0] - v0[:, 0]) * (v2[:, 1] - v0[:, 1])
                   - (v2[:, 0] - v0[:, 0]) * (v1[:, 1] - v0[:, 1])) / np.float32(2)
    probs = areas / areas.sum(dtype=np.float32)
    k1, k2, k3, _k4 = prng.split(k, 4, mode)
    tri_idx = prng.choice_p(k1, probs, n, mode)
    u = np.sqrt(prng.uniform(k2, n, mode=mode))
    v = prng.uniform(k3, n, mode=mode)
    a = np.float32(1) - u
    b = u * (np.float32(1) - v)
    c = u * v
    tv = tris[tri_idx]
    pts = a[:, None] * tv[:, 0] + b[:, None] * tv[:, 1] + c[:, None] * tv[:, 2]
    return pts.astype(np.float32)


def perturbation_delta(normals, k, mode=prng.PARTITIONABLE, dt=np.float32):
    """reflection.py:22-49."""
    normals = np.asarray(normals, dtype=dt)
    n = normals.shape[0]
    k1, k2 = prng.split(k, 2, mode)
    th1 = prng.normal(k1, n, mode).astype(dt)
    th2 = prng.normal(k2, n, mode).astype(dt)
    use_x = np.abs(normals[:, 2:3]) > dt(0.9)
    ref = np.where(use_x, np.array([1, 0, 0], dt), np.array([0, 0, 1], dt))
    t1 = np.cross(normals, ref)
    t1 = t1 / np.sqrt(np.sum(t1 * t1, axis=-1, keepdims=True))
    t2 = np.cross(normals, t1)
    return (th1[:, None] * t1 + th2[:, None] * t2).astype(dt)


def polygon_area(verts, dt=np.float32):
    """integrators.py:172-174 (shoelace)."""
    x = np.asarray(verts, dt)[:, 0]
    y = np.asarray(verts, dt)[:, 1]
    return dt(0.5) * np.abs(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


def sample_group(group, k, n_samples, mode=prng.PARTITIONABLE, dt=np.float32):
    """integrators.py:68-188.  ``group`` is an oracle scene group dict; returns a
    copy with points/normals/delta/weights filled: (F,M,3),(F,M,3),(F,M,3),(F,M,1)."""
    F = group["positions"].shape[0]
    c, kk, asph = group["curvature"], group["conic"], group["aspheric"]
    mkeys = prng.split(k, F, mode)
    P, N, D, W = [], [], [], []
    for f in range(F):
        ks, kp = prng.split(mkeys[f], 2, mode)
        off = group["offsets"][f]
        if group["kind"] == "disk":
            radius = np.float32(group["radii"][f])
            xy = sample_disk(ks, n_samples, mode) * radius
            area = dt(np.float32(np.pi)) * dt(radius) ** 2
        else:
            verts = group["vertices"][f]
            xy = sample_polygon(ks, verts, n_samples, mode)
            area = polygon_area(verts, dt)
        pts, nrm = point_and_normal(xy, off, c, kk, asph, dt)
        # the deltas are drawn from the f32 normals' tangent frame
        delta = perturbation_delta(nrm, kp, mode, dt)
        w = nrm[:, 2:3] / area * dt(n_samples)
        P.append(pts); N.append(nrm); D.append(delta); W.append(w.astype(dt))
    out = dict(group)
    out.update(points=np.stack(P), normals=np.stack(N), delta=np.stack(D), weights=np.stack(W))
    return out
