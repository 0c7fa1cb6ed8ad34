"""The oracle against the only result-level evidence the reference offers:
CT3 effective area ~100 m^2 on-axis (ResponseMatrix.ipynb cell 13), the f64 anchors recorded in
BASELINE.md, plus internal consistency of its f32 / f64 / C forms."""
import numpy as np
import pytest

from iactrace_b200.io import load_packed_config
from oracle import cport, prng, scene as oscene, trace as otrace
from _bridge import subset_config, point_grid, parallel_grid


@pytest.fixture(scope="module")
def ct3():
    return oscene.build_scene(load_packed_config("CT3"), 256, prng.key(0))


def test_ct3_effective_area_anchors(ct3):
    src = np.array([[0.0, 0.0, 1e10]], np.float32)
    val = np.ones(1, np.float32)
    pts, v = otrace.render_debug(ct3, src, val, "point", 0, np.float64)
    bare = dict(ct3, obstructions=[])
    _, v0 = otrace.render_debug(bare, src, val, "point", 0, np.float64)
    # 380 * pi * 0.3^2 = 107.4 m^2 of glass; BASELINE.md: 106.08 un-shadowed, 100.70 shadowed (MC noise 0.3 %)
    assert abs(v0.sum() - 106.08) < 0.05
    assert abs(v.sum() - 100.70) < 0.5
    assert abs(1 - v.sum() / v0.sum() - 0.0504) < 0.004
    lit = v > 0
    rms = np.sqrt(((pts[lit] ** 2).sum(1) * v[lit]).sum() / v.sum())
    assert abs(rms - 0.0101) < 0.0008                       # PSF rms ~10.1 mm on-axis
    th = np.deg2rad(1.0)
    pts, v = otrace.render_debug(ct3, np.array([[1e10 * np.tan(th), 0, 1e10]], np.float32), val, "point", 0, np.float64)
    lit = v > 0
    assert abs(v.sum() - 99.57) < 0.5
    assert abs((pts[lit, 0] * v[lit]).sum() / v.sum() - (-0.270)) < 0.002


def test_hex_grid_constants_match_survey(ct3):
    s = ct3["sensors"][0]
    assert s["lookup_table"].shape == (41, 36) and (s["q_min"], s["r_min"]) == (-20, -18)
    assert abs(s["hex_size"] - 0.0242122) < 1e-6 and abs(s["hex_inradius"] - 0.0209684) < 1e-6
    assert (s["lookup_table"] >= 0).sum() == 960
    ct5 = oscene.parse_config(load_packed_config("CT5"))[3]
    assert ct5[0]["lookup_table"].shape == (56, 56) and abs(np.rad2deg(ct5[0]["grid_rotation"]) - 29.9992) < 1e-3
    assert ct5[1]["lookup_table"].shape == (58, 57) and abs(np.rad2deg(ct5[1]["grid_rotation"]) - 59.950) < 1e-2


def test_f32_oracle_tracks_f64_oracle():
    sc = oscene.build_scene(subset_config(load_packed_config("CT5"), mirror_step=25), 40, prng.key(1))
    src = point_grid(3, 1.5)
    val = np.ones(9, np.float32)
    p32, v32 = otrace.render_debug(sc, src, val, "point", 0, np.float32)
    p64, v64 = otrace.render_debug(sc, src, val, "point", 0, np.float64)
    same = (v32 != 0) == (v64 != 0)
    assert same.mean() > 0.9995
    ok = same & (v64 != 0)
    np.testing.assert_allclose(v32[ok], v64[ok], rtol=2e-6)
    assert np.abs(p32[ok] - p64[ok]).max() < 6e-5           # 36 m lever arm in float32


@pytest.mark.parametrize("stype", ["point", "parallel"])
@pytest.mark.parametrize("sensor_idx", [0, 2])
def test_c_port_is_bit_identical_to_numpy_oracle(stype, sensor_idx):
    sc = oscene.build_scene(subset_config(load_packed_config("CT5"), mirror_step=40), 21, prng.key(0))
    src = point_grid(3, 1.5) if stype == "point" else parallel_grid(3, 3.0)
    val = np.linspace(0.5, 1.5, 9).astype(np.float32)
    prep = cport.prepare(sc, sensor_idx)
    xy, v = cport.render(prep, src, val, stype, debug=True)
    oxy, ov = otrace.render_debug(sc, src, val, stype, sensor_idx, np.float32)
    assert np.array_equal(v, ov)
    assert np.array_equal(xy, oxy)
    img, _ = cport.render(prep, src, val, stype, threads=3)
    oimg = otrace.render(sc, src, val, stype, sensor_idx, np.float32)
    np.testing.assert_allclose(img, oimg, rtol=1e-5, atol=1e-6 * oimg.max())
    # the float64 build of the same source (tools/parity_fullsize.py) against the NumPy oracle in float64
    xy64, v64 = cport.render(prep, src, val, stype, debug=True, variant="f64")
    oxy64, ov64 = otrace.render_debug(sc, src, val, stype, sensor_idx, np.float64)
    assert np.array_equal(v64 == 0, ov64 == 0)
    np.testing.assert_allclose(v64, ov64, rtol=2e-7, atol=1e-12)
    hit = np.abs(oxy64) < 1e9
    # (the C port takes float32 world tables, the NumPy oracle transforms in float64: 6e-8 on a normal x 36 m)
    np.testing.assert_allclose(xy64[hit], oxy64[hit], rtol=2e-7, atol=1e-5)


def test_response_matrix_and_render_consistency():
    sc = oscene.apply_roughness(oscene.build_scene(subset_config(load_packed_config("CT3"), mirror_step=19), 16, prng.key(42)), 24)
    src = parallel_grid(3, 5.5)
    val = np.ones(9, np.float32)
    M = otrace.render_response_matrix(sc, src, val, "parallel", 0, np.float64)
    img = otrace.render(sc, src, val, "parallel", 0, np.float64)
    np.testing.assert_allclose(M.sum(0), img, rtol=1e-12, atol=1e-12)
    assert M.shape == (9, 960)
    # roughness 24 arcsec in radians (operations.py:128)
    assert abs(float(sc["groups"][0]["scale"][0]) - 24 * np.pi / 648000) < 1e-10
