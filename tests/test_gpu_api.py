"""The reference-facing Python API on the GPU: YAML loading from a file, several stage-0 groups,
functional edits feeding the kernels, output options."""
import copy

import numpy as np
import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu

import iactrace_b200 as I
from iactrace_b200 import config
from iactrace_b200.core import render, render_debug, render_response_matrix
from iactrace_b200.io import build_telescope, load_packed_config
from oracle import prng, scene as oscene, trace as otrace
from _bridge import to_oracle_scene, subset_config, point_grid


def test_load_telescope_from_yaml_file_matches_oracle_end_to_end(tmp_path):
    """load_telescope(path, MCIntegrator(n), key) -> render, against the oracle built from the same YAML and seed
    (both sides sample independently from the same key tree)."""
    cfg = subset_config(load_packed_config("CT3"), mirror_step=12)
    path = tmp_path / "ct3_small.yaml"
    path.write_text(yaml.safe_dump(cfg))
    tel = I.load_telescope(str(path), I.MCIntegrator(50), key=I.random.key(42))
    tel_b = I.Telescope.from_yaml(path, I.MCIntegrator(50), key=42)
    assert tel.name == "CT1" and len(tel.mirror_groups) == 1 and len(tel.mirror_groups[0]) == 32
    assert torch.equal(tel.mirror_groups[0].points, tel_b.mirror_groups[0].points)
    osc = oscene.load_yaml(str(path), 50, prng.key(42))
    src = np.array([[0.0, 0.0, 1e10], [1e8, -5e7, 1e10]], np.float32)
    val = np.array([1.0, 2.0], np.float32)
    xy, v = tel(src, val, "point", debug=True)
    oxy, ov = otrace.render_debug(osc, src, val, "point", 0, np.float32)
    xy, v = xy.cpu().numpy(), v.cpu().numpy()
    flips = (v != 0) != (ov != 0)
    assert flips.mean() < 1e-3
    np.testing.assert_allclose(v[~flips], ov[~flips], rtol=2e-5)
    ok = ~flips & (ov != 0)
    assert np.abs(xy[ok] - oxy[ok]).max() < 1e-4
    img = tel(src, val).cpu().numpy()
    oimg = otrace.render(osc, src, val, "point", 0, np.float64)
    assert abs(img.sum() - oimg.sum()) < 2e-3 * oimg.sum()


def test_two_stage0_groups_are_concatenated_in_group_order():
    """Two surface templates -> two disk groups, each with its own key (yaml_loader.py:73-80); the packed
    world table concatenates them (render.py:202-205) and render_debug keeps facet-major order."""
    cfg = copy.deepcopy(subset_config(load_packed_config("CT3"), mirror_step=40))
    cfg["mirror_templates"]["other"] = dict(surface=dict(curvature=0.0331, conic=-0.2, aspheric=[]))
    for m in cfg["mirrors"][1::2]:
        m["template"] = "other"
    tel = build_telescope(cfg, I.MCIntegrator(20), I.random.key(3))
    assert [len(g) for g in tel.mirror_groups] == [5, 5]
    osc = oscene.build_scene(cfg, 20, prng.key(3))
    for g, og in zip(tel.mirror_groups, osc["groups"]):
        np.testing.assert_allclose(g.points.cpu().numpy(), og["points"], atol=3e-6)
    osc2 = to_oracle_scene(tel)
    src = point_grid(2, 0.5)
    val = np.ones(4, np.float32)
    xy, v = render_debug(tel, src, val, "point", 0)
    oxy, ov = otrace.render_debug(osc2, src, val, "point", 0, np.float64)
    assert xy.shape == (10 * 4 * 20, 2)
    flips = (v.cpu().numpy() != 0) != (ov != 0)
    assert flips.mean() < 1e-3
    ok = ~flips & (ov != 0)
    np.testing.assert_allclose(v.cpu().numpy()[ok], ov[ok], rtol=1e-5)
    assert np.abs(xy.cpu().numpy()[ok] - oxy[ok]).max() < 3e-5


def test_functional_edits_reach_the_kernels():
    tel = build_telescope(subset_config(load_packed_config("CT3"), mirror_step=10), I.MCIntegrator(64), I.random.key(0))
    src = np.array([[0.0, 0.0, 1e10]], np.float32)
    val = np.ones(1, np.float32)
    base = render(tel, src, val, "point", 1)
    # weights scale the Monte-Carlo estimate inversely (value = v cos / w)
    half = render(tel.scale_mirror_weights(0, 2.0), src, val, "point", 1)
    torch.testing.assert_close(half, 0.5 * base, rtol=1e-6, atol=0)
    per = torch.linspace(1.0, 3.0, len(tel.mirror_groups[0]))
    scaled = render(tel.scale_mirror_weights(0, per), src, val, "point", 1)
    assert float(scaled.sum()) < float(base.sum())
    # edits agree with the oracle applied to the same edit
    t2 = tel.apply_roughness(60).focus(0.05, 1).set_mirror_rotations(0, tel.mirror_groups[0].rotations * 1.01)
    got = render(t2, src, val, "point", 1).cpu().numpy()
    want = otrace.render(to_oracle_scene(t2), src, val, "point", 1, np.float64)
    assert abs(got.sum() - want.sum()) < 2e-3 * want.sum()
    assert abs(got.sum() - float(base.sum())) > 1e-3 or np.abs(got - base.cpu().numpy()).max() > 0
    # the original is untouched (cache keyed per Telescope)
    torch.testing.assert_close(render(tel, src, val, "point", 1), base, rtol=2e-6, atol=0)   # atomic order varies
    # removing obstructions only adds light
    assert float(render(tel.clear_obstructions(), src, val, "point", 1).sum()) > float(base.sum())


def test_numpy_output_option():
    tel = build_telescope(subset_config(load_packed_config("CT3"), n_mirrors=3), I.MCIntegrator(8), None)
    src = np.array([[0.0, 0.0, 1e10]], np.float32)
    val = np.ones(1, np.float32)
    ref = render(tel, src, val).cpu().numpy()
    config.return_numpy = True
    try:
        img = render(tel, src, val)
        pts, vals = render_debug(tel, src, val)
        mat = render_response_matrix(tel, src, val)
    finally:
        config.return_numpy = False
    assert isinstance(img, np.ndarray) and isinstance(pts, np.ndarray) and isinstance(mat, np.ndarray)
    np.testing.assert_allclose(img, ref, rtol=1e-6)          # atomic order varies run to run
    assert mat.shape == (1, 960) and pts.shape == (3 * 8, 2) and vals.shape == (24,)
