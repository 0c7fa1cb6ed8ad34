"""The bound behind the per-run culling of the leg towards optical stage 1 (csrc/iact_cull.cuh, leg_masks): for rows whose
normals lie within e of a mean normal nb and incoming directions within dd of d0, the reflected direction
r(d, n) = d - 2 (d.n) n (reflection.py:17-19) stays within
    dd (1 + 2 (|nb| + e)^2) + 2 |d0| e (2 |nb| + e)
of r(d0, nb).  Checked here in float64 on random inputs, including non-unit mean normals (the mean of unit normals is
shorter than 1 before normalisation) and non-unit directions (parallel sources are not normalised by the library)."""
import numpy as np


def _reflect(d, n):
    return d - 2.0 * np.sum(d * n, axis=-1, keepdims=True) * n


def test_reflected_direction_stays_inside_the_cone():
    rng = np.random.default_rng(0)
    n_cases, n_rows = 4000, 32
    nb = rng.normal(size=(n_cases, 1, 3))
    nb *= rng.uniform(0.7, 1.05, (n_cases, 1, 1)) / np.linalg.norm(nb, axis=-1, keepdims=True)
    d0 = rng.normal(size=(n_cases, 1, 3))
    d0 *= rng.uniform(0.8, 1.2, (n_cases, 1, 1)) / np.linalg.norm(d0, axis=-1, keepdims=True)
    e = 10 ** rng.uniform(-5, -0.7, (n_cases, 1, 1))
    dd = np.where(rng.uniform(size=(n_cases, 1, 1)) < 0.5, 0.0, 10 ** rng.uniform(-6, -1, (n_cases, 1, 1)))
    # rows anywhere inside the two balls, many of them on the surface (the worst case)
    def ball(radius):
        v = rng.normal(size=(n_cases, n_rows, 3))
        v /= np.linalg.norm(v, axis=-1, keepdims=True)
        return v * radius * np.where(rng.uniform(size=(n_cases, n_rows, 1)) < 0.5, 1.0, rng.uniform(size=(n_cases, n_rows, 1)))
    n = nb + ball(e)
    d = d0 + ball(dd)
    dev = np.linalg.norm(_reflect(d, n) - _reflect(d0, nb), axis=-1)
    nl, dl = np.linalg.norm(nb, axis=-1), np.linalg.norm(d0, axis=-1)
    bound = dd[..., 0] * (1 + 2 * (nl + e[..., 0]) ** 2) + 2 * dl * e[..., 0] * (2 * nl + e[..., 0])
    assert np.all(dev <= bound * (1 + 1e-12))
    # the bound is not vacuous: the worst row of a case comes within a factor of three of it on average
    ratio = dev.max(axis=1) / bound[:, 0]
    assert 0.25 < np.median(ratio) < 1.0
