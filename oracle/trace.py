"""Oracle restatement of the reference's render hot path (array-at-a-time NumPy).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Follows, op for op:
* ``iactrace/core/transforms.py:72-106``    euler_to_matrix
* ``iactrace/telescope/mirrors.py:64-79``   MirrorGroup.transform_to_world
* ``iactrace/core/intersections.py:6-285``  intersect_plane/cylinder/box/oriented_box/triangle/sphere/conic
* ``iactrace/core/intersections.py:290-367`` newton_raphson_intersect
* ``iactrace/core/surfaces.py:67-107``      AsphericSurface.intersect
* ``iactrace/core/reflection.py:5-19``      reflect
* ``iactrace/core/render.py:21-324``        _check_occlusions, _reflect_at_stage, _intersect_group,
                                            _trace_single_mirror, render, render_debug, render_response_matrix
* ``iactrace/sensors/square.py:66-91,144-172``  SquareSensor / DifferentiableSquareSensor.accumulate
* ``iactrace/sensors/hexagonal.py:174-194,264-314`` HexagonalSensor / DifferentiableHexagonalSensor.accumulate

``dt=np.float32`` reproduces the reference's arithmetic type; ``dt=np.float64``
evaluates the same formulas "exactly" so that f32 summation-order noise of
either side is visible (SURVEY.md hazard H6).
"""
from __future__ import annotations

import numpy as np

from .scene import SQRT3, SQRT3_2, SQRT3_3, rotate2d, cartesian_to_axial

INF = np.inf


def _dot(a, b):
    return np.sum(a * b, axis=-1)


def _norm(a):
    return np.sqrt(np.sum(a * a, axis=-1, keepdims=True))


# --------------------------------------------------------------------------- transforms
def euler_to_matrix(ttr, dt=np.float32):
    """transforms.py:72-106: degrees, R = Rz(rotation) @ Ry(tilt) @ Rx(tip)."""
    ttr = np.asarray(ttr, dtype=dt)
    rx, ry, rz = (ttr * dt(np.pi / 180.0)).astype(dt)
    cx, sx, cy, sy, cz, sz = (dt(np.cos(rx)), dt(np.sin(rx)), dt(np.cos(ry)), dt(np.sin(ry)),
                              dt(np.cos(rz)), dt(np.sin(rz)))
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dt)
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dt)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dt)
    return (Rz @ Ry @ Rx).astype(dt)


def transform_to_world(g, dt=np.float32):
    """mirrors.py:64-79 -> tp (F,M,3), tn (F,M,3), tw (F,M,1)."""
    F = g["positions"].shape[0]
    tp, tn = [], []
    for f in range(F):
        R = euler_to_matrix(g["rotations"][f], dt)
        p = np.einsum("ij,nj->ni", R, g["points"][f].astype(dt)) + g["positions"][f].astype(dt)
        n = np.einsum("ij,nj->ni", R, g["normals"][f].astype(dt))
        d = np.einsum("ij,nj->ni", R, g["delta"][f].astype(dt))
        pert = n + dt(g["scale"][f]) * d
        pert = pert / _norm(pert)
        tp.append(p.astype(dt))
        tn.append(pert.astype(dt))
    return np.stack(tp), np.stack(tn), g["weights"].astype(dt)


# --------------------------------------------------------------------------- primitives
def intersect_plane(o, d, center, R, dt=np.float32):
    """intersections.py:6-41 -> (...,2)."""
    u1, u2, n = R[:, 0], R[:, 1], R[:, 2]
    ndotd = _dot(d, n)
    ndoto = _dot(o, n)
    ndotp = np.sum(n * center)
    parallel = np.abs(ndotd) < dt(1e-10)
    safe = np.where(parallel, dt(1.0), ndotd)
    with np.errstate(all="ignore"):
        t = (ndotp - ndoto) / safe
        hit = o + t[..., None] * d
        op = hit - center
        x = _dot(op, u1)
        y = _dot(op, u2)
    invalid = parallel | (t <= 0)
    x = np.where(invalid, dt(1e10), x)
    y = np.where(invalid, dt(1e10), y)
    return np.stack([x, y], axis=-1).astype(dt)


def intersect_cylinder(o, d, p1, p2, radius, dt=np.float32):
    """intersections.py:44-87.  o,d: (...,3); p1,p2: (K,3); radius: (K,) -> t (...,K)."""
    o = o[..., None, :]
    d = d[..., None, :]
    axis = p2 - p1
    height = np.sqrt(np.sum(axis * axis, axis=-1))
    axis = axis / height[:, None]
    oc = o - p1
    oc_ax = _dot(oc, axis)
    rd_ax = _dot(d, axis)
    oc_perp = oc - oc_ax[..., None] * axis
    rd_perp = d - rd_ax[..., None] * axis
    a = _dot(rd_perp, rd_perp)
    b = dt(2) * _dot(oc_perp, rd_perp)
    c = _dot(oc_perp, oc_perp) - radius * radius
    disc = b * b - dt(4) * a * c
    eps = dt(1e-8)
    with np.errstate(all="ignore"):
        sq = np.sqrt(np.maximum(disc, dt(0)))
        t1 = (-b - sq) / (dt(2) * a + eps)
        t2 = (-b + sq) / (dt(2) * a + eps)
        y1 = oc_ax + t1 * rd_ax
        y2 = oc_ax + t2 * rd_ax
        t1 = np.where((t1 > eps) & (y1 >= 0) & (y1 <= height) & (disc >= 0), t1, INF)
        t2 = np.where((t2 > eps) & (y2 >= 0) & (y2 <= height) & (disc >= 0), t2, INF)
        tb = -oc_ax / (rd_ax + eps)
        tt = (height - oc_ax) / (rd_ax + eps)
        pb = oc_perp + tb[..., None] * rd_perp
        pt = oc_perp + tt[..., None] * rd_perp
        tb = np.where((tb > eps) & (_dot(pb, pb) <= radius ** 2), tb, INF)
        tt = np.where((tt > eps) & (_dot(pt, pt) <= radius ** 2), tt, INF)
    return np.minimum(np.minimum(t1, t2), np.minimum(tb, tt))


def intersect_box(o, d, p1, p2, dt=np.float32):
    """intersections.py:90-110 -> (...,K)."""
    eps = dt(1e-8)
    o = o[..., None, :]
    d = d[..., None, :]
    bmin = np.minimum(p1, p2)
    bmax = np.maximum(p1, p2)
    with np.errstate(all="ignore"):
        inv = dt(1.0) / (d + eps)
        t1 = (bmin - o) * inv
        t2 = (bmax - o) * inv
    tn = np.minimum(t1, t2)
    tf = np.maximum(t1, t2)
    tmin = np.max(tn, axis=-1)
    tmax = np.min(tf, axis=-1)
    hit = (tmax >= tmin) & (tmax > eps)
    tr = np.where(tmin > eps, tmin, tmax)
    return np.where(hit, tr, INF)


def intersect_oriented_box(o, d, center, half, rot, dt=np.float32):
    """intersections.py:113-149 -> (...,K)."""
    eps = dt(1e-8)
    oc = o[..., None, :] - center                       # (...,K,3)
    lo = np.einsum("kji,...kj->...ki", rot, oc)         # rot.T @ (o - c)
    ld = np.einsum("kji,...j->...ki", rot, d)
    with np.errstate(all="ignore"):
        inv = dt(1.0) / (ld + eps * np.sign(ld + eps))
        t1 = (-half - lo) * inv
        t2 = (half - lo) * inv
    tn = np.minimum(t1, t2)
    tf = np.maximum(t1, t2)
    tmin = np.max(tn, axis=-1)
    tmax = np.min(tf, axis=-1)
    hit = (tmax >= tmin) & (tmax > eps)
    tr = np.where(tmin > eps, tmin, tmax)
    return np.where(hit & (tr > eps), tr, INF)


def intersect_triangle(o, d, v0, v1, v2, dt=np.float32):
    """intersections.py:152-192 (Moeller-Trumbore) -> (...,K)."""
    eps = dt(1e-8)
    e1 = v1 - v0
    e2 = v2 - v0
    dd = np.broadcast_to(d[..., None, :], d.shape[:-1] + e2.shape)
    h = np.cross(dd, e2)
    a = _dot(e1, h)
    parallel = np.abs(a) < eps
    with np.errstate(all="ignore"):
        f = dt(1.0) / (a + eps * np.sign(a + eps))
        s = o[..., None, :] - v0
        u = f * _dot(s, h)
        q = np.cross(s, e1)
        v = f * _dot(dd, q)
        t = f * _dot(e2, q)
    valid = (~parallel) & (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t > eps)
    return np.where(valid, t, INF)


def intersect_sphere(o, d, center, radius, dt=np.float32):
    """intersections.py:195-226 -> (...,K)."""
    eps = dt(1e-8)
    oc = o[..., None, :] - center
    dd = d[..., None, :]
    a = _dot(dd, dd)
    b = dt(2.0) * _dot(oc, dd)
    c = _dot(oc, oc) - radius * radius
    disc = b * b - dt(4.0) * a * c
    with np.errstate(all="ignore"):
        sq = np.sqrt(np.maximum(disc, dt(0)))
        t1 = (-b - sq) / (dt(2.0) * a + eps)
        t2 = (-b + sq) / (dt(2.0) * a + eps)
    t1 = np.where((t1 > eps) & (disc >= 0), t1, INF)
    t2 = np.where((t2 > eps) & (disc >= 0), t2, INF)
    return np.minimum(t1, t2)


def group_intersect(o, d, g, dt=np.float32):
    """obstructions.py:74-79,115-120,156-161,204-209,250-255: min t over the group's primitives."""
    c = lambda a: np.asarray(a, dtype=dt)
    t = g["type"]
    if t == "cylinder":
        ts = intersect_cylinder(o, d, c(g["p1"]), c(g["p2"]), c(g["r"]), dt)
    elif t == "box":
        ts = intersect_box(o, d, c(g["p1"]), c(g["p2"]), dt)
    elif t == "sphere":
        ts = intersect_sphere(o, d, c(g["centers"]), c(g["radii"]), dt)
    elif t == "oriented_box":
        ts = intersect_oriented_box(o, d, c(g["centers"]), c(g["half_extents"]), c(g["rotations"]), dt)
    elif t == "triangle":
        ts = intersect_triangle(o, d, c(g["v0"]), c(g["v1"]), c(g["v2"]), dt)
    else:
        raise ValueError(t)
    return np.min(ts, axis=-1)


def check_occlusions(o, d, obstruction_groups, dt=np.float32):
    """render.py:21-41 -> shadow mask (1 = lit)."""
    mask = np.ones(o.shape[:-1], dtype=dt)
    for g in obstruction_groups or []:
        t = group_intersect(o, d, g, dt)
        mask = mask * np.where(t < dt(1e10), dt(0.0), dt(1.0))
    return mask


def reflect(d, n, dt=np.float32):
    """reflection.py:5-19 -> (reflected, -cos)."""
    c = np.sum(d * n, axis=-1, keepdims=True)
    return (d - dt(2.0) * c * n).astype(dt), -c


# --------------------------------------------------------------------------- stage >= 1
def _sag_raw_jit(x, y, c, k, asph, dt):
    """surfaces.py:25-39 evaluated inside jit: c,k are weak f32 scalars, so
    ``(1+k)*c*c*r2`` is evaluated left to right in ``dt``."""
    r2 = x * x + y * y
    with np.errstate(all="ignore"):
        denom = dt(1) + np.sqrt(dt(1) - (dt(1) + dt(k)) * dt(c) * dt(c) * r2)
        z = r2 * dt(c) / denom
    if len(asph) > 0:
        powers = np.arange(2, 2 + 2 * len(asph), 2).astype(dt)
        z = z + np.sum(np.asarray(asph, dt) * r2[..., None] ** powers, axis=-1)
    return z


def _dsag_raw_jit(x, y, c, k, asph, dt):
    r2 = x * x + y * y
    kc2 = (dt(1) + dt(k)) * dt(c) * dt(c)
    with np.errstate(all="ignore"):
        s = np.sqrt(dt(1) - kc2 * r2)
        dd = dt(1) + s
        dz = dt(c) / dd - (r2 * dt(c)) / (dd * dd) * ((dt(0.5) / s) * (-kc2))
    if len(asph) > 0:
        powers = np.arange(2, 2 + 2 * len(asph), 2).astype(dt)
        dz = dz + np.sum(np.asarray(asph, dt) * powers * r2[..., None] ** (powers - dt(1)), axis=-1)
    return dz * (x + x), dz * (y + y)


def intersect_conic(o, d, c, k, dt=np.float32):
    """intersections.py:229-285."""
    ox, oy, oz = o[..., 0], o[..., 1], o[..., 2]
    dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
    c = dt(c)
    k1 = dt(1) + dt(k)
    A = c * (dx * dx + dy * dy + k1 * dz * dz)
    B = dt(2) * (c * (ox * dx + oy * dy + k1 * oz * dz) - dz)
    C = c * (ox * ox + oy * oy + k1 * oz * oz) - dt(2) * oz
    is_plane = np.abs(c) < dt(1e-12)
    with np.errstate(all="ignore"):
        t_plane = np.where(np.abs(dz) > dt(1e-10), -oz / dz, INF)
        disc = B * B - dt(4) * A * C
        none = disc < 0
        sq = np.sqrt(np.maximum(disc, dt(0)))
        t1 = (-B - sq) / (dt(2) * A + dt(1e-30))
        t2 = (-B + sq) / (dt(2) * A + dt(1e-30))
    v1 = t1 > dt(1e-8)
    v2 = t2 > dt(1e-8)
    tc = np.where(v1 & v2, np.minimum(t1, t2), np.where(v1, t1, np.where(v2, t2, INF)))
    tc = np.where(none, INF, tc)
    return np.where(is_plane, t_plane, tc)


def surface_intersect(o, d, offset, c, k, asph, dt=np.float32, max_iter=10, tol=1e-8):
    """surfaces.py:67-107 + intersections.py:290-367 -> (t, point, normal)."""
    x0, y0 = dt(offset[0]), dt(offset[1])
    z0 = _sag_raw_jit(x0, y0, c, k, asph, dt)
    o_raw = np.stack([o[..., 0] + x0, o[..., 1] + y0, o[..., 2] + z0], axis=-1)
    t = intersect_conic(o_raw, d, c, k, dt)
    ox, oy, oz = o[..., 0], o[..., 1], o[..., 2]
    dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]

    def g(tt):
        with np.errstate(all="ignore"):
            x = ox + tt * dx
            y = oy + tt * dy
            z = oz + tt * dz
            return z - (_sag_raw_jit(x + x0, y + y0, c, k, asph, dt) - z0)

    def gp(tt):
        with np.errstate(all="ignore"):
            x = ox + tt * dx
            y = oy + tt * dy
            sx, sy = _dsag_raw_jit(x + x0, y + y0, c, k, asph, dt)
            return dz - (sx * dx + sy * dy)

    conv = np.zeros(t.shape, bool)
    tol = dt(tol)
    with np.errstate(all="ignore"):
        for _ in range(max_iter):
            gv = g(t)
            gd = gp(t)
            gd = np.where(np.abs(gd) > dt(1e-12), gd, dt(1e-12))
            tn = t - gv / gd
            new_conv = conv | (np.abs(gv) < tol)
            t = np.where(conv, t, tn)
            conv = new_conv
        xh = ox + t * dx
        yh = oy + t * dy
        resid = np.abs(g(t))
        valid = (t > dt(1e-8)) & (resid < tol * dt(100))
        t_out = np.where(valid, t, INF)
        zh = _sag_raw_jit(xh + x0, yh + y0, c, k, asph, dt) - z0
        pt = np.stack([xh, yh, zh], axis=-1)
        sx, sy = _dsag_raw_jit(xh + x0, yh + y0, c, k, asph, dt)
        n = np.stack([-sx, -sy, np.ones_like(sx)], axis=-1)
        n = n / _norm(n)
    return t_out, pt, n


def check_aperture(g, x, y, mi, dt=np.float32):
    """mirrors.py:147-149 (disk), 209-220 (convex polygon, CCW)."""
    with np.errstate(all="ignore"):
        if g["kind"] == "disk":
            return x ** 2 + y ** 2 <= dt(g["radii"][mi]) ** 2
        verts = np.asarray(g["vertices"][mi], dt)
        n = len(verts)
        inside = np.ones(x.shape, bool)
        for i in range(n):
            v1, v2 = verts[i], verts[(i + 1) % n]
            cross = (v2[0] - v1[0]) * (y - v1[1]) - (v2[1] - v1[1]) * (x - v1[0])
            inside &= cross >= 0
        return inside


def intersect_group(o, d, g, dt=np.float32):
    """render.py:82-115 -> best_t, best_pts, best_norms over the mirrors of one group."""
    all_t, all_p, all_n = [], [], []
    for mi in range(g["positions"].shape[0]):
        pos = g["positions"][mi].astype(dt)
        R = euler_to_matrix(g["rotations"][mi], dt)
        with np.errstate(all="ignore"):
            ol = np.einsum("ij,...j->...i", R.T, o - pos)
            dl = np.einsum("ij,...j->...i", R.T, d)
            ts, pl, nl = surface_intersect(ol, dl, g["offsets"][mi], g["curvature"], g["conic"],
                                           g["aspheric"], dt)
            ok = check_aperture(g, pl[..., 0], pl[..., 1], mi, dt)
            ts = np.where(ok, ts, INF)
            pw = np.einsum("ij,...j->...i", R, pl) + pos
            nw = np.einsum("ij,...j->...i", R, nl)
        all_t.append(ts); all_p.append(pw); all_n.append(nw)
    all_t = np.stack(all_t)
    closest = np.argmin(all_t, axis=0)
    best_t = np.min(all_t, axis=0)
    best_p = np.take_along_axis(np.stack(all_p), closest[None, ..., None], axis=0)[0]
    best_n = np.take_along_axis(np.stack(all_n), closest[None, ..., None], axis=0)[0]
    return best_t, best_p, best_n


def reflect_at_stage(o, d, v, stage_groups, obstruction_groups, dt=np.float32):
    """render.py:44-79."""
    best_t = np.full(o.shape[:-1], INF, dt)
    best_p = np.zeros(o.shape, dt)
    best_n = np.zeros(o.shape, dt)
    for g in stage_groups:
        t, p, n = intersect_group(o, d, g, dt)
        closer = t < best_t
        best_t = np.where(closer, t, best_t)
        best_p = np.where(closer[..., None], p, best_p)
        best_n = np.where(closer[..., None], n, best_n)
    refl, cos = reflect(d, best_n, dt)
    hit = best_t < dt(1e10)
    shadow = check_occlusions(o, d, obstruction_groups, dt)
    with np.errstate(all="ignore"):
        nv = v * hit * shadow * np.abs(cos[..., 0])
    return best_p.astype(dt), refl.astype(dt), nv.astype(dt)


# --------------------------------------------------------------------------- sensors
def _to_i32(a):
    with np.errstate(all="ignore"):
        return np.clip(np.nan_to_num(a, nan=-2.0e9), -2.0e9, 2.0e9).astype(np.int64).astype(np.int32)


def square_index(s, x, y, dt=np.float32):
    """square.py:68-84 -> (flat_idx, valid, edge_distance)."""
    with np.errstate(all="ignore"):
        xc = (x - dt(s["x0"])) / dt(s["dx"])
        yc = (y - dt(s["y0"])) / dt(s["dy"])
        xi = _to_i32(np.floor(xc))
        yi = _to_i32(np.floor(yc))
        valid = (xi >= 0) & (xi < s["width"]) & (yi >= 0) & (yi < s["height"])
        xf = xc - xi.astype(dt)
        yf = yc - yi.astype(dt)
        dist = np.minimum(np.minimum(xf, dt(1) - xf) * dt(s["dx"]), np.minimum(yf, dt(1) - yf) * dt(s["dy"]))
        valid = valid & ~(dist < dt(s["edge_width"]))
    xi = np.clip(xi, 0, s["width"] - 1)
    yi = np.clip(yi, 0, s["height"] - 1)
    return yi * s["width"] + xi, valid, dist


def axial_round(q, r):
    """hexagonal.py:32-39 (round half to even)."""
    s = -q - r
    qi, ri, si = np.round(q), np.round(r), np.round(s)
    dq, dr, ds = np.abs(qi - q), np.abs(ri - r), np.abs(si - s)
    qi = np.where((dq > dr) & (dq > ds), -ri - si, qi)
    ri = np.where((dr > dq) & (dr > ds), -qi - si, ri)
    return qi, ri


def hex_norm(x, y, inradius, dt=np.float32):
    """hexagonal.py:42-47."""
    return np.maximum(np.abs(x), dt(0.5) * np.abs(x) + dt(SQRT3_2) * np.abs(y)) / dt(inradius)


def _hex_lookup(s, qi, ri):
    """hexagonal.py:155-172."""
    T = s["lookup_table"]
    qx = qi - s["q_min"]
    rx = ri - s["r_min"]
    inb = (qx >= 0) & (qx < T.shape[0]) & (rx >= 0) & (rx < T.shape[1])
    pix = T[np.clip(qx, 0, T.shape[0] - 1), np.clip(rx, 0, T.shape[1] - 1)]
    valid = inb & (pix >= 0)
    return np.where(valid, pix, 0), valid


def hex_index(s, x, y, dt=np.float32):
    """hexagonal.py:174-191 -> (pixel_idx, valid, hex_dist)."""
    with np.errstate(all="ignore"):
        xg, yg = rotate2d(x - dt(s["grid_offset"][0]), y - dt(s["grid_offset"][1]), -s["grid_rotation"], dt)
        q, r = cartesian_to_axial(xg, yg, s["hex_size"], dt)
        qi, ri = axial_round(q, r)
        pix, valid = _hex_lookup(s, _to_i32(qi), _to_i32(ri))
        cx = dt(s["hex_size"] * SQRT3) * (qi + ri / dt(2))
        cy = dt(s["hex_size"] * 1.5) * ri
        hd = hex_norm(xg - cx, yg - cy, s["hex_inradius"], dt)
        thr = dt(1.0 - s["edge_width"] / s["hex_inradius"])
        valid = valid & ~(hd > thr)
    return pix, valid, hd


def _segment_sum(vals, idx, n, dt):
    return np.bincount(idx.ravel(), weights=vals.ravel().astype(np.float64), minlength=n).astype(dt) \
        if dt == np.float64 else _segment_sum_f32(vals.ravel(), idx.ravel(), n)


def _segment_sum_f32(vals, idx, n):
    # sequential f32 scatter-add, as jax.ops.segment_sum does on XLA:CPU
    out = np.zeros(n, np.float32)
    np.add.at(out, idx, vals.astype(np.float32))
    return out


def accumulator_shape(s):
    return (s["height"], s["width"]) if "square" in s["type"] else (s["n_pixels"],)


def accumulate(s, x, y, v, dt=np.float32):
    """The four ``accumulate`` methods; returns an array of accumulator_shape(s)."""
    x = np.asarray(x, dt).ravel()
    y = np.asarray(y, dt).ravel()
    v = np.asarray(v, dt).ravel()
    t = s["type"]
    if t == "square":
        idx, valid, _ = square_index(s, x, y, dt)
        img = _segment_sum(np.where(valid, v, dt(0)), idx, s["height"] * s["width"], dt)
        return img.reshape(s["height"], s["width"])
    if t == "hexagonal":
        idx, valid, _ = hex_index(s, x, y, dt)
        return _segment_sum(np.where(valid, v, dt(0)), idx, s["n_pixels"], dt)
    if t == "soft_square":
        # square.py:144-172
        with np.errstate(all="ignore"):
            xp = (x - dt(s["x0"])) / dt(s["dx"])
            yp = (y - dt(s["y0"])) / dt(s["dy"])
            xb = _to_i32(np.floor(xp))
            yb = _to_i32(np.floor(yp))
            xf = xp - xb.astype(dt)
            yf = yp - yb.astype(dt)
            ox = s["offset_x"][None, :]
            oy = s["offset_y"][None, :]
            xi = xb[:, None].astype(np.int64) + ox
            yi = yb[:, None].astype(np.int64) + oy
            ddx = xf[:, None] - ox.astype(dt)
            ddy = yf[:, None] - oy.astype(dt)
            w = np.exp(dt(-0.5) * (ddx ** 2 + ddy ** 2) / dt(s["sigma"] ** 2))
            w = w / np.sum(w, axis=1, keepdims=True)
            valid = (xi >= 0) & (xi < s["width"]) & (yi >= 0) & (yi < s["height"])
            xi = np.clip(xi, 0, s["width"] - 1)
            yi = np.clip(yi, 0, s["height"] - 1)
            spl = v[:, None] * w * valid
        img = _segment_sum(np.nan_to_num(spl), (yi * s["width"] + xi), s["height"] * s["width"], dt)
        return img.reshape(s["height"], s["width"])
    if t == "soft_hexagonal":
        # hexagonal.py:264-314
        with np.errstate(all="ignore"):
            xg, yg = rotate2d(x - dt(s["grid_offset"][0]), y - dt(s["grid_offset"][1]), -s["grid_rotation"], dt)
            q, r = cartesian_to_axial(xg, yg, s["hex_size"], dt)
            qb, rb = axial_round(q, r)
            bx = dt(s["hex_size"] * SQRT3) * (qb + rb / dt(2))
            by = dt(s["hex_size"] * 1.5) * rb
            ddx = xg - bx
            ddy = yg - by
            qi = _to_i32(qb)[:, None].astype(np.int64) + s["nb_q"][None, :]
            ri = _to_i32(rb)[:, None].astype(np.int64) + s["nb_r"][None, :]
            nq = s["nb_q"].astype(dt)
            nr = s["nb_r"].astype(dt)
            nbx = dt(s["hex_size"] * SQRT3) * (nq + nr / dt(2))
            nby = dt(s["hex_size"] * 1.5) * nr
            hx = ddx[:, None] - nbx[None, :]
            hy = ddy[:, None] - nby[None, :]
            hd = hex_norm(hx, hy, s["hex_inradius"], dt)
            w = np.exp(dt(-0.5) * (hd / dt(s["sigma"])) ** 2)
            w = w / np.sum(w, axis=1, keepdims=True)
            pix, valid = _hex_lookup(s, qi, ri)
            spl = v[:, None] * w * valid
        return _segment_sum(np.nan_to_num(spl), pix, s["n_pixels"], dt)
    raise ValueError(t)


# --------------------------------------------------------------------------- render drivers
def _stages(groups):
    """render.py:12-18."""
    by = {}
    for g in groups:
        by.setdefault(g["stage"], []).append(g)
    return dict(sorted(by.items()))


def _primary_tables(stages, dt):
    data = [transform_to_world(g, dt) for g in stages[0]]
    return (np.concatenate([d[0] for d in data]), np.concatenate([d[1] for d in data]),
            np.concatenate([d[2] for d in data]))


def trace_single_mirror(f, tp, tn, tw, sources, values, source_type, stage_idx, stages, obs,
                        spos, srot, dt=np.float32):
    """render.py:118-157 -> pts (S,M,2), vals (S,M)."""
    S = sources.shape[0]
    M = tp.shape[1]
    with np.errstate(all="ignore"):
        if source_type == "point":
            dirs = tp[f][None, :, :] - sources[:, None, :]
            dirs = dirs / _norm(dirs)
        else:
            dirs = np.broadcast_to(sources[:, None, :], (S, M, 3)).astype(dt)
        origins = np.broadcast_to(tp[f][None], dirs.shape)
        normals = np.broadcast_to(tn[f][None], dirs.shape)
        shadow = check_occlusions(origins, -dirs, obs, dt)
        refl, cos = reflect(dirs, normals, dt)
        vals = values[:, None] * cos[..., 0] / tw[f][None, :, 0] * shadow
        o_cur, d_cur, v_cur = origins, refl, vals.astype(dt)
        for si in stage_idx[1:]:
            o_cur, d_cur, v_cur = reflect_at_stage(o_cur, d_cur, v_cur, stages[si], obs, dt)
        pts = intersect_plane(o_cur, d_cur, spos, srot, dt)
    return pts, v_cur


def _setup(scene, sources, values, sensor_idx, dt):
    s = scene["sensors"][sensor_idx]
    spos = s["position"].astype(dt)
    srot = euler_to_matrix(s["rotation"], dt)
    stages = _stages(scene["groups"])
    return s, spos, srot, stages, sorted(stages.keys()), np.asarray(sources, dt), np.asarray(values, dt)


def render(scene, sources, values, source_type="point", sensor_idx=0, dt=np.float32):
    """render.py:174-220."""
    s, spos, srot, stages, sidx, sources, values = _setup(scene, sources, values, sensor_idx, dt)
    acc = np.zeros(accumulator_shape(s), dt)
    if not sidx or 0 not in stages:
        return acc
    tp, tn, tw = _primary_tables(stages, dt)
    for f in range(tp.shape[0]):
        pts, v = trace_single_mirror(f, tp, tn, tw, sources, values, source_type, sidx, stages,
                                     scene["obstructions"], spos, srot, dt)
        acc = acc + accumulate(s, pts[..., 0], pts[..., 1], v, dt)
    return acc


def render_debug(scene, sources, values, source_type="point", sensor_idx=0, dt=np.float32):
    """render.py:223-268 -> pts (F*S*M,2), vals (F*S*M,), facet-major then source then sample."""
    s, spos, srot, stages, sidx, sources, values = _setup(scene, sources, values, sensor_idx, dt)
    if not sidx or 0 not in stages:
        return np.zeros((0, 2), dt), np.zeros((0,), dt)
    tp, tn, tw = _primary_tables(stages, dt)
    P, V = [], []
    for f in range(tp.shape[0]):
        pts, v = trace_single_mirror(f, tp, tn, tw, sources, values, source_type, sidx, stages,
                                     scene["obstructions"], spos, srot, dt)
        P.append(pts.reshape(-1, 2)); V.append(v.reshape(-1))
    return np.concatenate(P), np.concatenate(V)


def render_response_matrix(scene, sources, values, source_type="point", sensor_idx=0, dt=np.float32):
    """render.py:271-324 -> (S, n_pixels)."""
    s, spos, srot, stages, sidx, sources, values = _setup(scene, sources, values, sensor_idx, dt)
    S = sources.shape[0]
    npx = int(np.prod(accumulator_shape(s)))
    acc = np.zeros((S, npx), dt)
    if not sidx or 0 not in stages:
        return acc
    tp, tn, tw = _primary_tables(stages, dt)
    for f in range(tp.shape[0]):
        pts, v = trace_single_mirror(f, tp, tn, tw, sources, values, source_type, sidx, stages,
                                     scene["obstructions"], spos, srot, dt)
        rows = np.stack([accumulate(s, pts[i, :, 0], pts[i, :, 1], v[i], dt).reshape(-1) for i in range(S)])
        acc = acc + rows
    return acc


def pixel_index(s, x, y, dt=np.float32):
    """Per-ray (index, valid, edge-distance measure) for hard sensors; used by the parity tests
    to separate rays that sit within rounding noise of a pixel edge."""
    if s["type"] == "square":
        return square_index(s, np.asarray(x, dt), np.asarray(y, dt), dt)
    if s["type"] == "hexagonal":
        return hex_index(s, np.asarray(x, dt), np.asarray(y, dt), dt)
    raise ValueError(s["type"])
