"""Golden-vector cases shared by the generator (make_golden.py, runs the REFERENCE) and the tests
(which run the oracle and the CUDA path on the same inputs).  Everything here is deterministic and
derives from the packaged CT3/CT5 scenes, so the tests need neither the reference nor YAML files."""
from __future__ import annotations

import copy

import numpy as np

from iactrace_b200.io import load_packed_config


def _subset(cfg, idx):
    c = copy.deepcopy(cfg)
    c["mirrors"] = [c["mirrors"][i] for i in idx]
    return c


def cfg_ct3_small():
    c = _subset(load_packed_config("CT3"), [0, 57, 123, 200, 301, 379])
    c["obstructions"] = c["obstructions"][:6] + c["obstructions"][-3:]
    return c


def cfg_ct5_small():
    c = _subset(load_packed_config("CT5"), [3, 150, 420, 640, 875])
    obs = c["obstructions"]
    c["obstructions"] = obs[:3] + obs[100:103] + obs[-1:] + [
        dict(type="sphere", center=[2.0, -3.0, 20.0], r=1.5),
        dict(type="oriented_box", center=[-4.0, 2.0, 18.0], half_extents=[1.5, 0.5, 2.0],
             rotation=[[0.8660254, -0.5, 0.0], [0.5, 0.8660254, 0.0], [0.0, 0.0, 1.0]]),
        dict(type="triangle", v0=[-9.0, -9.0, 25.0], v1=[-2.0, -8.0, 26.0], v2=[-6.0, -1.0, 24.0]),
    ]
    return c


def cfg_cassegrain():
    mirrors = []
    for ang in (0, 60, 120, 180, 240, 300):
        a = np.radians(ang)
        x, y = float(2.0 * np.cos(a)), float(2.0 * np.sin(a))
        mirrors.append(dict(id=f"P{ang}", template="primary", position=[x, y, 0.0], orientation=[0.0, 0.0, 0.0],
                            aperture=dict(type="circular", radius=1.0), offset=[x, y], stage=0))
    mirrors.append(dict(id="S", template="secondary", position=[0.0, 0.0, 6.0], orientation=[180.0, 0.0, 0.0],
                        aperture=dict(type="circular", radius=1.0), offset=[0.0, 0.0], stage=1))
    obs = [dict(type="cylinder", p1=[0.9, 0.0, 6.2], p2=[3.2, 0.0, 0.3], r=0.03),
           dict(type="cylinder", p1=[0.0, -0.9, 6.2], p2=[0.0, -3.2, 0.3], r=0.03),
           dict(type="box", p1=[3.3, -0.3, 0.0], p2=[3.9, 0.3, 0.8]),
           dict(type="sphere", center=[-3.6, 0.0, 0.5], r=0.3)]
    return dict(telescope=dict(name="test_cassegrain", units="m"),
                mirror_templates=dict(primary=dict(surface=dict(curvature=0.05, conic=-1.0, aspheric=[])),
                                      secondary=dict(surface=dict(curvature=-0.05, conic=-1.0, aspheric=[]))),
                mirrors=mirrors, obstructions=obs,
                sensors=[dict(type="square", position=[0.0, 0.0, -0.45], orientation=[0.0, 0.0, 0.0], width=64,
                              height=64, bounds=[-0.5, 0.5, -0.5, 0.5])])


def cfg_polygon_secondary():
    """Two-stage telescope whose secondary has a CCW square polygon aperture and a tilted sensor."""
    c = cfg_cassegrain()
    c["mirrors"][-1]["aperture"] = dict(type="polygon", vertices=[[-0.8, -0.8], [0.8, -0.8], [0.8, 0.8], [-0.8, 0.8]])
    c["sensors"][0]["orientation"] = [2.0, -1.0, 10.0]
    c["sensors"][0]["edge_width"] = 0.002
    return c


def _parallel(n, seed, fov_deg):
    rng = np.random.default_rng(seed)
    f = np.deg2rad(fov_deg)
    d = np.stack([rng.uniform(-f / 2, f / 2, n), rng.uniform(-f / 2, f / 2, n), -np.ones(n)], 1)
    return (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)


def _points(n, seed, half_deg, dist=1e10):
    rng = np.random.default_rng(seed)
    t = np.deg2rad(rng.uniform(-half_deg, half_deg, (n, 2)))
    return np.stack([dist * np.tan(t[:, 0]), dist * np.tan(t[:, 1]), np.full(n, dist)], 1).astype(np.float32)


CASES = {
    # name: config builder, n_samples, seed, rng mode, roughness arcsec, sources, source_type, sensor indices
    "ct3_point": dict(cfg=cfg_ct3_small, M=9, seed=42, mode="partitionable", rough=24.0, src=_points(3, 1, 1.0), stype="point", sensors=(0, 1)),
    "ct3_parallel_legacy": dict(cfg=cfg_ct3_small, M=8, seed=7, mode="legacy", rough=0.0, src=_parallel(3, 2, 3.0), stype="parallel", sensors=(0,)),
    "ct5_point": dict(cfg=cfg_ct5_small, M=7, seed=0, mode="partitionable", rough=10.0, src=_points(3, 3, 1.2), stype="point", sensors=(0, 1, 2)),
    "ct5_near_source": dict(cfg=cfg_ct5_small, M=6, seed=5, mode="partitionable", rough=0.0,
                            src=np.array([[2.0, -1.0, 72.0], [-6.0, 4.0, 150.0]], np.float32), stype="point", sensors=(0,)),
    "cassegrain": dict(cfg=cfg_cassegrain, M=10, seed=0, mode="partitionable", rough=0.0, src=_parallel(4, 4, 1.0), stype="parallel", sensors=(0,)),
    "polygon_secondary": dict(cfg=cfg_polygon_secondary, M=8, seed=3, mode="partitionable", rough=5.0, src=_parallel(3, 5, 0.6), stype="parallel", sensors=(0,)),
}


def cfg_ct3_full():
    return copy.deepcopy(load_packed_config("CT3"))


def cfg_ct5_full():
    return copy.deepcopy(load_packed_config("CT5"))


def cfg_pentagons():
    """Irregular pentagon facets: the fan triangles from vertex 0 have very uneven areas (about 1 : 5 : 2), which is
    what `jax.random.choice(p=areas)` in sample_polygon (utils/sampling.py:53) has to get right; two facet shapes in
    one group would be split by the reference's grouping, so all facets share one vertex list."""
    verts = [[-0.45, -0.30], [0.10, -0.42], [0.48, 0.05], [0.05, 0.40], [-0.35, 0.22]]       # counter-clockwise
    mirrors = [dict(id=f"F{i}", template="sphere", position=[float(x), float(y), 0.02 * i], orientation=[0.3 * i, -0.2 * i, 10.0 * i],
                    aperture=dict(type="polygon", vertices=verts), stage=0)
               for i, (x, y) in enumerate([(-1.2, -0.6), (0.0, -0.9), (1.1, -0.4), (-0.8, 0.7), (0.5, 0.9)])]
    return dict(telescope=dict(name="pentagons", units="m"),
                mirror_templates=dict(sphere=dict(surface=dict(curvature=0.0333, conic=0.0, aspheric=[]))),
                mirrors=mirrors,
                obstructions=[dict(type="cylinder", p1=[-2.0, 0.1, 6.0], p2=[2.0, -0.1, 6.5], r=0.08)],
                sensors=[dict(type="square", position=[0.0, 0.0, 15.0], orientation=[0.0, 0.0, 0.0], width=40, height=40,
                              bounds=[-0.4, 0.4, -0.4, 0.4])])


# Full-size scenes (all 380 / 876 facets, all 33 / 271 obstructions) and the uneven-polygon sampler case.  Generated
# separately (make_golden.py --large: the reference runs on a Python-loop vmap) into reference_golden_large.npz.
LARGE_CASES = {
    "ct3_full": dict(cfg=cfg_ct3_full, M=8, seed=0, mode="partitionable", rough=0.0, src=_points(2, 11, 1.0), stype="point", sensors=(0,)),
    "ct5_full": dict(cfg=cfg_ct5_full, M=4, seed=0, mode="partitionable", rough=0.0, src=_points(1, 12, 1.2), stype="point", sensors=(0,)),
    "pentagons": dict(cfg=cfg_pentagons, M=64, seed=21, mode="partitionable", rough=0.0, src=_points(2, 13, 0.5), stype="point", sensors=(0,)),
    "pentagons_legacy": dict(cfg=cfg_pentagons, M=33, seed=22, mode="legacy", rough=0.0, src=_points(1, 14, 0.5), stype="point", sensors=(0,)),
}
ALL_CASES = {**CASES, **LARGE_CASES}


def case_values(name):
    n = len(ALL_CASES[name]["src"])
    return np.linspace(0.5, 1.5, n).astype(np.float32)
