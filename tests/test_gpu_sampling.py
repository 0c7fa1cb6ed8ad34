"""CUDA sampler (K1) vs the oracle restatement of MCIntegrator.sample_group, same keys."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import iactrace_b200 as I
from iactrace_b200 import random as R
from iactrace_b200.io import build_telescope, load_packed_config
from oracle import prng, scene as oscene
from _bridge import subset_config, cassegrain_config


def _cmp_groups(tel, osc, atol=3e-6):
    for g, og in zip(tel.mirror_groups, osc["groups"]):
        if g.optical_stage != 0:
            assert g.points.shape[1] == 0
            continue
        p, n = g.points.cpu().numpy(), g.normals.cpu().numpy()
        d, w = g.perturbation_delta.cpu().numpy(), g.weights.cpu().numpy()
        assert p.shape == og["points"].shape and w.shape == og["weights"].shape
        np.testing.assert_allclose(p, og["points"], rtol=0, atol=atol)
        np.testing.assert_allclose(n, og["normals"], rtol=0, atol=atol)
        # deltas are O(1)..O(4) standard normals in the tangent plane
        np.testing.assert_allclose(d, og["delta"], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(w, og["weights"], rtol=2e-6)


@pytest.mark.parametrize("mode", [R.PARTITIONABLE, R.LEGACY])
@pytest.mark.parametrize("n_samples", [1, 33, 64])
def test_disk_group_matches_oracle(mode, n_samples):
    cfg = subset_config(load_packed_config("CT3"), n_mirrors=12, mirror_step=29)
    R.set_rng_mode(mode)
    try:
        tel = build_telescope(cfg, I.MCIntegrator(n_samples), R.key(42))
    finally:
        R.set_rng_mode(R.PARTITIONABLE)
    osc = oscene.build_scene(cfg, n_samples, prng.key(42), mode)
    _cmp_groups(tel, osc)


@pytest.mark.parametrize("mode", [R.PARTITIONABLE, R.LEGACY])
def test_polygon_group_matches_oracle(mode):
    cfg = subset_config(load_packed_config("CT5"), n_mirrors=9, mirror_step=97)
    R.set_rng_mode(mode)
    try:
        tel = build_telescope(cfg, I.MCIntegrator(57), None)
    finally:
        R.set_rng_mode(R.PARTITIONABLE)
    osc = oscene.build_scene(cfg, 57, None, mode)
    _cmp_groups(tel, osc)
    # every sample lies inside the hexagon (flat-to-flat 0.9 m) and on the sphere
    p = tel.mirror_groups[0].points.cpu().numpy()
    assert np.all(np.hypot(p[..., 0], p[..., 1]) <= 0.9 / np.sqrt(3) + 1e-5)


def test_offset_paraboloid_group_and_key_chain():
    """Cassegrain primary: per-facet offsets on a parent paraboloid; stage 1 stays unsampled."""
    cfg = cassegrain_config()
    tel = build_telescope(cfg, I.MCIntegrator(40), R.key(0))
    osc = oscene.build_scene(cfg, 40, prng.key(0))
    assert [g.optical_stage for g in tel.mirror_groups] == [0, 1]
    _cmp_groups(tel, osc)


def test_random_normal_uniform_match_oracle():
    for mode, omode in ((R.PARTITIONABLE, prng.PARTITIONABLE), (R.LEGACY, prng.LEGACY)):
        for n in (1, 2, 7, 380):
            got = R.normal(R.key(4242), n, mode).cpu().numpy()
            np.testing.assert_allclose(got, prng.normal(prng.key(4242), n, omode), rtol=3e-6, atol=1e-7)
            got = R.uniform(R.key(7), n, -0.5, 0.25, mode).cpu().numpy()
            np.testing.assert_array_equal(got, prng.uniform(prng.key(7), n, -0.5, 0.25, omode))
    # the public known-answer values straight from the GPU
    assert abs(R.normal(R.key(42), 1, R.PARTITIONABLE).item() - (-0.028304616)) < 1e-7
    assert abs(R.normal(R.key(42), 1, R.LEGACY).item() - (-0.18471177)) < 1e-7


def test_resample_mirrors_key_chain():
    cfg = subset_config(load_packed_config("CT3"), n_mirrors=5)
    tel = build_telescope(cfg, I.MCIntegrator(8), R.key(1))
    tel2 = tel.resample_mirrors(I.MCIntegrator(16), R.key(9))
    from oracle import sample as osample
    og = oscene.build_scene(cfg, 8, prng.key(1))["groups"][0]
    k = prng.split(prng.key(9), 1)[0]          # operations.py:36 split(key, n_groups)
    ref = osample.sample_group(og, k, 16)
    np.testing.assert_allclose(tel2.mirror_groups[0].points.cpu().numpy(), ref["points"], atol=3e-6)
    assert tel.mirror_groups[0].points.shape[1] == 8     # original untouched
