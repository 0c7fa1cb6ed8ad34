"""Synthetic benchmark inputs frozen in BASELINE.md (sources for configs 1-5, the Cassegrain scene).

Pure NumPy helpers shared by ``bench.py`` and the tests; they define workloads, not algorithms.
"""
from __future__ import annotations

import numpy as np


def point_grid(n_side: int, half_deg: float, dist: float = 1e10) -> np.ndarray:
    """BASELINE config 2 sources: n_side^2 point sources on a grid of field angles
    theta_x, theta_y in linspace(-half_deg, +half_deg), src = (d tan tx, d tan ty, d)."""
    th = np.deg2rad(np.linspace(-half_deg, half_deg, n_side))
    tx, ty = np.meshgrid(th, th, indexing="xy")
    return np.stack([dist * np.tan(tx).ravel(), dist * np.tan(ty).ravel(), np.full(tx.size, dist)], 1).astype(np.float32)


def parallel_grid(n_side: int, fov_deg: float) -> np.ndarray:
    """examples/ResponseMatrix.ipynb cell 9: n_side^2 normalised directions over a square field of view."""
    fov = np.float32(fov_deg * np.pi / 180)
    x1 = np.linspace(-fov / 2, fov / 2, n_side, dtype=np.float32)
    X, Y = np.meshgrid(x1, x1, indexing="xy")
    d = np.stack([X.ravel(), Y.ravel(), -np.ones(n_side * n_side, np.float32)], 1).astype(np.float32)
    return (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)


def star_field(n: int, fov_deg: float = 3.0, seed: int = 42):
    """examples/Cassegrain.ipynb cell 8 in spirit: n directions uniform in a fov_deg box around -z and
    fluxes 10^(-10 U) (NumPy generators stand in for the notebook's jax.random draws)."""
    rng = np.random.default_rng(seed)
    f = np.deg2rad(fov_deg)
    d = np.stack([rng.uniform(-f / 2, f / 2, n), rng.uniform(-f / 2, f / 2, n), -np.ones(n)], 1)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    flux = (10 ** (-10 * np.random.default_rng(4242).uniform(size=n))).astype(np.float32)
    return d, flux


def cassegrain_config(with_obstructions: bool = True) -> dict:
    """The two-mirror telescope of examples/Cassegrain.ipynb cell 3 as a YAML-schema dict (six
    paraboloid segments r = 1 at radius 2 with parent-surface offsets, hyperbolic secondary at z = 6,
    1024^2 square sensor at z = -0.45) plus BASELINE.md's synthetic obstructions for config 3."""
    mirrors = []
    for ang in (0, 60, 120, 180, 240, 300):
        a = np.radians(ang)
        x, y = float(2.0 * np.cos(a)), float(2.0 * np.sin(a))
        mirrors.append(dict(id=f"P{ang}", template="primary", position=[x, y, 0.0], orientation=[0.0, 0.0, 0.0],
                            aperture=dict(type="circular", radius=1.0), offset=[x, y], stage=0))
    mirrors.append(dict(id="S", template="secondary", position=[0.0, 0.0, 6.0], orientation=[180.0, 0.0, 0.0],
                        aperture=dict(type="circular", radius=1.0), offset=[0.0, 0.0], stage=1))
    obs = []
    if with_obstructions:
        for (a, b) in (((0.9, 0, 6.2), (3.2, 0, 0.3)), ((-0.9, 0, 6.2), (-3.2, 0, 0.3)),
                       ((0, 0.9, 6.2), (0, 3.2, 0.3)), ((0, -0.9, 6.2), (0, -3.2, 0.3))):
            obs.append(dict(type="cylinder", p1=list(map(float, a)), p2=list(map(float, b)), r=0.03))
        obs.append(dict(type="box", p1=[3.3, -0.3, 0.0], p2=[3.9, 0.3, 0.8]))
        obs.append(dict(type="sphere", center=[-3.6, 0.0, 0.5], r=0.3))
    return dict(telescope=dict(name="test_cassegrain", units="m"),
                mirror_templates=dict(primary=dict(surface=dict(curvature=0.05, conic=-1.0, aspheric=[])),
                                      secondary=dict(surface=dict(curvature=-0.05, conic=-1.0, aspheric=[]))),
                mirrors=mirrors, obstructions=obs,
                sensors=[dict(type="square", position=[0.0, 0.0, -0.45], orientation=[0.0, 0.0, 0.0], width=1024,
                              height=1024, bounds=[-0.5, 0.5, -0.5, 0.5])])
