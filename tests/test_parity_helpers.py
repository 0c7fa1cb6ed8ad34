"""CPU check of the parity machinery itself (tests/_parity.py): the float32 oracle plays the kernel, the float64
oracle the reference.  Also shows that the ambiguity accounting is not vacuous: a deliberately corrupted image or
pixel index is caught."""
import numpy as np
import pytest

from iactrace_b200.io import load_packed_config
from oracle import prng, scene as oscene, trace as otrace
from _bridge import subset_config, point_grid
from _parity import ray_parity, compare_image, subset_rays


def _rays(sensor_idx):
    cfg = subset_config(load_packed_config("CT3"), mirror_step=12)
    sc = oscene.build_scene(cfg, 24, prng.key(0))
    src = point_grid(2, 0.6)
    val = np.linspace(0.5, 1.5, len(src)).astype(np.float32)
    xy, v = otrace.render_debug(sc, src, val, "point", sensor_idx, np.float32)
    oxy, ov = otrace.render_debug(sc, src, val, "point", sensor_idx, np.float64)
    s = sc["sensors"][sensor_idx]
    idx, valid, _ = otrace.pixel_index(s, xy[:, 0], xy[:, 1], np.float32)
    pix = np.where(valid, idx, -1).astype(np.int32)
    img = otrace.accumulate(s, xy[:, 0], xy[:, 1], v, np.float64)
    return sc, s, xy, v, pix, oxy, ov, img


@pytest.mark.parametrize("sensor_idx", [0, 1])
def test_parity_helpers_accept_f32_vs_f64_oracle(sensor_idx):
    sc, s, xy, v, pix, oxy, ov, img = _rays(sensor_idx)
    r = ray_parity(xy, v, pix, oxy, ov, s, xy_tol=5e-5, flip_budget=1e-3)
    st = compare_image(img, r, min_lit=3)
    assert st["lit_pixels_compared"] >= 3 and st["flux_share_compared"] > 0.5
    assert r["stats"]["n_lit"] > 100
    # one source's rays against that source's image (response-matrix rows)
    F, S, M = len(v) // (4 * 24), 4, 24
    src_of_ray = (np.arange(len(v)) // M) % S
    m = src_of_ray == 2
    row = otrace.accumulate(s, xy[m, 0], xy[m, 1], v[m], np.float64)
    compare_image(row, subset_rays(r, m), min_lit=1)


def test_parity_helpers_catch_a_wrong_image_and_a_wrong_pixel():
    sc, s, xy, v, pix, oxy, ov, img = _rays(0)
    r = ray_parity(xy, v, pix, oxy, ov, s, xy_tol=5e-5, flip_budget=1e-3)
    bad = img.copy()
    k = int(np.argmax(bad))
    bad[k] *= 1.0 + 5e-4                                  # 5e-4 on the brightest pixel: above the 1e-4 bar
    with pytest.raises(AssertionError):
        compare_image(bad, r)
    wrong = pix.copy()
    lit = np.flatnonzero((v != 0) & (pix >= 0))
    wrong[lit[0]] = (wrong[lit[0]] + 1) % 960             # a ray binned into a neighbour without sitting on an edge
    with pytest.raises(AssertionError):
        ray_parity(xy, v, wrong, oxy, ov, s, xy_tol=5e-5, flip_budget=1e-3)
    dark = v.copy()
    dark[lit[:50]] = 0.0                                  # 50 spurious shadow decisions: over the flip budget
    with pytest.raises(AssertionError):
        ray_parity(xy, dark, pix, oxy, ov, s, xy_tol=5e-5, flip_budget=1e-3)
