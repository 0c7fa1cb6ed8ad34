"""Oracle for gradients: the single-mirror render restated in differentiable torch (float64, CPU).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Reverse-mode autodiff of this restatement is what
``jax.grad`` of the reference's ``render`` (``iactrace/core/render.py:174-220``) produces for
single-stage telescopes: the same chain ``transform_to_world`` (``telescope/mirrors.py:64-79``) ->
directions (``render.py:129-133``) -> ``reflect`` (``core/reflection.py:5-19``) -> value
(``render.py:141``) -> ``intersect_plane`` (``core/intersections.py:6-41``) -> ``accumulate``
(``sensors/square.py:66-91,144-172``, ``sensors/hexagonal.py:174-194,264-314``), with the shadow mask
and every index / rounding decision held constant (zero gradient), exactly as JAX treats
``where`` / ``floor`` / ``round``.  The shadow mask itself comes from the NumPy oracle.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import trace as otrace
from .scene import SQRT3, SQRT3_2, SQRT3_3

DT = torch.float64


def euler_to_matrix(e):
    a = e * (math.pi / 180.0)
    cx, sx, cy, sy, cz, sz = torch.cos(a[0]), torch.sin(a[0]), torch.cos(a[1]), torch.sin(a[1]), torch.cos(a[2]), torch.sin(a[2])
    one, zero = torch.ones_like(cx), torch.zeros_like(cx)
    Rx = torch.stack([torch.stack([one, zero, zero]), torch.stack([zero, cx, -sx]), torch.stack([zero, sx, cx])])
    Ry = torch.stack([torch.stack([cy, zero, sy]), torch.stack([zero, one, zero]), torch.stack([-sy, zero, cy])])
    Rz = torch.stack([torch.stack([cz, -sz, zero]), torch.stack([sz, cz, zero]), torch.stack([zero, zero, one])])
    return Rz @ Ry @ Rx


def _accumulate(s, x, y, v):
    t = s["type"]
    xn, yn = x.detach().numpy(), y.detach().numpy()
    if t in ("square", "hexagonal"):
        idx, valid, _ = otrace.pixel_index(s, xn, yn, np.float64)
        n = int(np.prod(otrace.accumulator_shape(s)))
        out = torch.zeros(n, dtype=DT)
        out = out.index_add(0, torch.from_numpy(idx.astype(np.int64)), v * torch.from_numpy(valid.astype(np.float64)))
        return out.reshape(otrace.accumulator_shape(s))
    if t == "soft_square":
        xp = (x - s["x0"]) / s["dx"]
        yp = (y - s["y0"]) / s["dy"]
        xb, yb = torch.floor(xp).detach(), torch.floor(yp).detach()
        fx, fy = xp - xb, yp - yb
        ox = torch.from_numpy(s["offset_x"].astype(np.float64))[None, :]
        oy = torch.from_numpy(s["offset_y"].astype(np.float64))[None, :]
        w = torch.exp(-0.5 * ((fx[:, None] - ox) ** 2 + (fy[:, None] - oy) ** 2) / s["sigma"] ** 2)
        w = w / w.sum(1, keepdim=True)
        xi = xb[:, None] + ox
        yi = yb[:, None] + oy
        valid = (xi >= 0) & (xi < s["width"]) & (yi >= 0) & (yi < s["height"])
        idx = (yi.clamp(0, s["height"] - 1) * s["width"] + xi.clamp(0, s["width"] - 1)).long()
        out = torch.zeros(s["height"] * s["width"], dtype=DT)
        out = out.index_add(0, idx.reshape(-1), (v[:, None] * w * valid).reshape(-1))
        return out.reshape(s["height"], s["width"])
    if t == "soft_hexagonal":
        ca, sa = math.cos(-s["grid_rotation"]), math.sin(-s["grid_rotation"])
        tx, ty = x - s["grid_offset"][0], y - s["grid_offset"][1]
        xg, yg = ca * tx - sa * ty, sa * tx + ca * ty
        q = (SQRT3_3 * xg - yg / 3) / s["hex_size"]
        r = (2 * yg / 3) / s["hex_size"]
        qb, rb = otrace.axial_round(q.detach().numpy(), r.detach().numpy())
        qb, rb = torch.from_numpy(qb), torch.from_numpy(rb)
        dx = xg - s["hex_size"] * SQRT3 * (qb + rb / 2)
        dy = yg - s["hex_size"] * 1.5 * rb
        nq = torch.from_numpy(s["nb_q"].astype(np.float64))
        nr = torch.from_numpy(s["nb_r"].astype(np.float64))
        hx = dx[:, None] - (s["hex_size"] * SQRT3 * (nq + nr / 2))[None, :]
        hy = dy[:, None] - (s["hex_size"] * 1.5 * nr)[None, :]
        hd = torch.maximum(hx.abs(), 0.5 * hx.abs() + SQRT3_2 * hy.abs()) / s["hex_inradius"]
        w = torch.exp(-0.5 * (hd / s["sigma"]) ** 2)
        w = w / w.sum(1, keepdim=True)
        qi = qb[:, None].numpy().astype(np.int64) + s["nb_q"][None, :]
        ri = rb[:, None].numpy().astype(np.int64) + s["nb_r"][None, :]
        pix, valid = otrace._hex_lookup(s, qi, ri)
        out = torch.zeros(s["n_pixels"], dtype=DT)
        return out.index_add(0, torch.from_numpy(pix.astype(np.int64)).reshape(-1),
                             (v[:, None] * w * torch.from_numpy(valid.astype(np.float64))).reshape(-1))
    raise ValueError(t)


def _sag_t(x, y, c, k, asph):
    r2 = x * x + y * y
    z = r2 * c / (1 + torch.sqrt(torch.as_tensor(1 - (1 + k) * c * c * r2, dtype=DT)))
    for i, a in enumerate(np.asarray(asph, np.float64).tolist()):
        z = z + a * r2 ** (2 * i + 2)
    return z


def _dsag_t(x, y, c, k, asph):
    r2 = x * x + y * y
    f1 = 0.5 * c / torch.sqrt(torch.as_tensor(1 - (1 + k) * c * c * r2, dtype=DT))
    for i, a in enumerate(np.asarray(asph, np.float64).tolist()):
        f1 = f1 + a * (2 * i + 2) * r2 ** (2 * i + 1)
    return 2 * x * f1, 2 * y * f1


def _reflect_at_stage(o, d, val, stage_groups, obstructions, stage_leaves=None):
    """render.py:44-115 for one optical stage >= 1, differentiable in (o, d, val).
    The mirror choice, the hit mask and the shadow mask come from the NumPy oracle (constants); the
    ray parameter is the NumPy oracle's converged Newton root refined by one differentiable Newton
    step, whose derivative is the implicit-function derivative that autodiff of the reference's
    10-step scan converges to."""
    on, dn = o.detach().numpy(), d.detach().numpy()
    best_t = np.full(on.shape[:-1], np.inf)
    best_key = np.full(on.shape[:-1], -1)
    cands = []
    for g in stage_groups:
        for mi in range(g["positions"].shape[0]):
            pos = g["positions"][mi].astype(np.float64)
            R = otrace.euler_to_matrix(g["rotations"][mi], np.float64)
            with np.errstate(all="ignore"):
                ol = np.einsum("ij,...j->...i", R.T, on - pos)
                dl = np.einsum("ij,...j->...i", R.T, dn)
                ts, pl, _ = otrace.surface_intersect(ol, dl, g["offsets"][mi], g["curvature"], g["conic"], g["aspheric"], np.float64)
                ok = otrace.check_aperture(g, pl[..., 0], pl[..., 1], mi, np.float64)
            ts = np.where(ok, ts, np.inf)
            closer = ts < best_t
            best_t = np.where(closer, ts, best_t)
            best_key = np.where(closer, len(cands), best_key)
            cands.append((g, mi))
    hit = best_t < 1e10
    shadow = otrace.check_occlusions(on, dn, obstructions, np.float64)
    p_out = torch.zeros_like(o)
    n_out = torch.zeros_like(o)
    for key, (g, mi) in enumerate(cands):
        m = torch.from_numpy((best_key == key) & hit)
        if not m.any():
            continue
        c, k, asph = g["curvature"], g["conic"], g["aspheric"]
        lv = stage_leaves.get(id(g)) if stage_leaves else None
        pos = lv["positions"][mi] if lv else torch.tensor(g["positions"][mi].astype(np.float64))
        R = euler_to_matrix(lv["rotations"][mi] if lv else torch.tensor(g["rotations"][mi].astype(np.float64)))
        x0, y0 = torch.tensor(float(g["offsets"][mi][0]), dtype=DT), torch.tensor(float(g["offsets"][mi][1]), dtype=DT)
        if lv:                                                  # surface parameters as leaves (surfaces.py:25-45)
            c, k = lv.get("curvature", c), lv.get("conic", k)
            if "offsets" in lv:
                x0, y0 = lv["offsets"][mi][0], lv["offsets"][mi][1]
        z0 = _sag_t(x0, y0, c, k, asph)
        ol = (o - pos) @ R          # R^T (o - pos)
        dl = d @ R
        t0 = torch.from_numpy(np.where(np.isfinite(best_t), best_t, 1.0))
        x, y = ol[..., 0] + t0 * dl[..., 0], ol[..., 1] + t0 * dl[..., 1]
        gval = (ol[..., 2] + t0 * dl[..., 2]) - (_sag_t(x + x0, y + y0, c, k, asph) - z0)
        sx, sy = _dsag_t(x + x0, y + y0, c, k, asph)
        t = t0 - gval / (dl[..., 2] - (sx * dl[..., 0] + sy * dl[..., 1]))
        x, y = ol[..., 0] + t * dl[..., 0], ol[..., 1] + t * dl[..., 1]
        pl = torch.stack([x, y, _sag_t(x + x0, y + y0, c, k, asph) - z0], -1)
        sx, sy = _dsag_t(x + x0, y + y0, c, k, asph)
        nl = torch.stack([-sx, -sy, torch.ones_like(sx)], -1)
        nl = nl / nl.norm(dim=-1, keepdim=True)
        pw = pl @ R.T + pos
        nw = nl @ R.T
        p_out = torch.where(m[..., None], pw, p_out)
        n_out = torch.where(m[..., None], nw, n_out)
    cos = (d * n_out).sum(-1)
    refl = d - 2 * cos[..., None] * n_out
    new_val = val * torch.from_numpy(hit.astype(np.float64)) * torch.from_numpy(shadow) * cos.abs()
    return p_out, refl, new_val


def render(scene, leaves, sources, values, source_type="point", sensor_idx=0):
    """Differentiable render.  ``leaves`` = dict of float64 torch tensors (may require grad):
    positions (F,3), rotations (F,3), scale (F,), weights (F,M,1), sensor_position (3,), sensor_rotation (3,);
    ``sources`` (S,3) and ``values`` (S,) float64 torch tensors.  One stage-0 group; groups of later
    optical stages are traversed with ``_reflect_at_stage``."""
    g = scene["groups"][0]
    assert g["stage"] == 0 and all(gr["stage"] > 0 for gr in scene["groups"][1:]), "one stage-0 group, then later stages"
    later = {}
    stage_leaves = {}
    for i, gr in enumerate(scene["groups"][1:]):
        later.setdefault(gr["stage"], []).append(gr)
        if leaves.get("stage"):
            stage_leaves[id(gr)] = leaves["stage"][i]     # {"positions": (N,3), "rotations": (N,3)} torch leaves
    s = scene["sensors"][sensor_idx]
    pts = leaves["points"] if "points" in leaves else torch.from_numpy(g["points"].astype(np.float64))
    nrm = leaves["normals"] if "normals" in leaves else torch.from_numpy(g["normals"].astype(np.float64))
    dlt = leaves["delta"] if "delta" in leaves else torch.from_numpy(g["delta"].astype(np.float64))
    F = pts.shape[0]
    Rs = euler_to_matrix(leaves["sensor_rotation"])
    u1, u2, ns = Rs[:, 0], Rs[:, 1], Rs[:, 2]
    ps = leaves["sensor_position"]
    img = torch.zeros(otrace.accumulator_shape(s), dtype=DT)
    # constant shadow mask from the NumPy oracle, evaluated at the current parameter values
    sc_now = dict(scene)
    gnow = dict(g, positions=leaves["positions"].detach().numpy(), rotations=leaves["rotations"].detach().numpy(),
                scale=leaves["scale"].detach().numpy())
    sc_now["groups"] = [gnow]
    tp, tn, _ = otrace.transform_to_world(gnow, np.float64)
    for f in range(F):
        R = euler_to_matrix(leaves["rotations"][f])
        p = pts[f] @ R.T + leaves["positions"][f]
        nw = nrm[f] @ R.T + leaves["scale"][f] * (dlt[f] @ R.T)
        n = nw / nw.norm(dim=-1, keepdim=True)
        if source_type == "point":
            d = p[None] - sources[:, None, :]
            d = d / d.norm(dim=-1, keepdim=True)
        else:
            d = sources[:, None, :].expand(-1, p.shape[0], -1)
        dnp = d.detach().numpy()
        shadow = torch.from_numpy(otrace.check_occlusions(np.broadcast_to(tp[f][None], dnp.shape), -dnp,
                                                          scene["obstructions"], np.float64))
        c = (d * n[None]).sum(-1)
        r = d - 2 * c[..., None] * n[None]
        val = values[:, None] * (-c) / leaves["weights"][f][None, :, 0] * shadow
        o_cur = p[None].expand_as(r)
        for stage in sorted(later):
            o_cur, r, val = _reflect_at_stage(o_cur, r, val, later[stage], scene["obstructions"], stage_leaves)
        ndotd = (r * ns).sum(-1)
        t = ((ns * ps).sum() - (o_cur * ns).sum(-1)) / ndotd
        h = o_cur + t[..., None] * r - ps
        x, y = (h * u1).sum(-1), (h * u2).sum(-1)
        ok = (t > 0) & (ndotd.abs() >= 1e-10)
        x = torch.where(ok, x, torch.full_like(x, 1e10))
        y = torch.where(ok, y, torch.full_like(y, 1e10))
        img = img + _accumulate(s, x.reshape(-1), y.reshape(-1), val.reshape(-1))
    return img
