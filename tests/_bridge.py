"""Test helpers: product Telescope -> oracle scene dict, synthetic scenes, comparison utilities."""
from __future__ import annotations

import numpy as np

from oracle import scene as oscene


def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)


def sensor_to_oracle(s):
    from iactrace_b200.sensors import (DifferentiableHexagonalSensor, DifferentiableSquareSensor, HexagonalSensor,
                                       SquareSensor)
    if isinstance(s, HexagonalSensor):
        hard = oscene.make_hex_sensor(_np(s.position), _np(s.rotation), _np(s.hex_centers), s.edge_width,
                                      grid=s.grid_constants())
        if isinstance(s, DifferentiableHexagonalSensor):
            return oscene.make_soft_hex_sensor(hard, s.sigma, s.kernel_size)
        return hard
    if isinstance(s, SquareSensor):
        bounds = (s.x0, s.x0 + s.dx * s.width, s.y0, s.y0 + s.dy * s.height)
        if isinstance(s, DifferentiableSquareSensor):
            o = oscene.make_soft_square_sensor(_np(s.position), _np(s.rotation), s.width, s.height, bounds,
                                               s.sigma, s.kernel_size)
        else:
            o = oscene.make_square_sensor(_np(s.position), _np(s.rotation), s.width, s.height, bounds, s.edge_width)
        o.update(x0=s.x0, y0=s.y0, dx=s.dx, dy=s.dy)
        return o
    raise TypeError(type(s))


def to_oracle_scene(tel):
    """Oracle scene fed with the PRODUCT's sample tables (isolates tracing from sampling)."""
    from iactrace_b200.core import obstructions as O
    groups = []
    for g in tel.mirror_groups:
        d = dict(kind=g.kind, stage=g.optical_stage, positions=_np(g.positions), rotations=_np(g.rotations),
                 offsets=_np(g.offsets), curvature=g.curvature, conic=g.conic, aspheric=np.asarray(g.aspheric),
                 points=_np(g.points), normals=_np(g.normals), weights=_np(g.weights),
                 delta=_np(g.perturbation_delta), scale=_np(g.perturbation_scale))
        if g.kind == "disk":
            d["radii"] = _np(g.radii)
        else:
            d["vertices"] = _np(g.vertices)
        groups.append(d)
    obs = []
    for g in tel.obstruction_groups or []:
        if isinstance(g, O.CylinderGroup):
            obs.append(dict(type="cylinder", p1=_np(g.p1), p2=_np(g.p2), r=_np(g.r)))
        elif isinstance(g, O.BoxGroup):
            obs.append(dict(type="box", p1=_np(g.p1), p2=_np(g.p2)))
        elif isinstance(g, O.SphereGroup):
            obs.append(dict(type="sphere", centers=_np(g.centers), radii=_np(g.radii)))
        elif isinstance(g, O.OrientedBoxGroup):
            obs.append(dict(type="oriented_box", centers=_np(g.centers), half_extents=_np(g.half_extents),
                            rotations=_np(g.rotations)))
        elif isinstance(g, O.TriangleGroup):
            obs.append(dict(type="triangle", v0=_np(g.v0), v1=_np(g.v1), v2=_np(g.v2)))
    return dict(name=tel.name, groups=groups, obstructions=obs, sensors=[sensor_to_oracle(s) for s in tel.sensors])


def subset_config(cfg, n_mirrors=None, mirror_step=1):
    """A smaller telescope: every ``mirror_step``-th mirror, at most ``n_mirrors``."""
    c = dict(cfg)
    m = cfg["mirrors"][::mirror_step]
    c["mirrors"] = m[:n_mirrors] if n_mirrors else m
    return c


from iactrace_b200.workloads import cassegrain_config, point_grid, parallel_grid  # noqa: E402,F401
