"""BASELINE.json's full-size configurations, checked through size-independent properties
(the oracle cannot run 4e8 rays): additivity over source subsets, linearity in the fluxes,
response-matrix rows summing to the render, permutation invariance, shadowing monotonicity."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import iactrace_b200 as I
from iactrace_b200.core import render, render_response_matrix
from iactrace_b200.io import build_telescope, load_packed_config
from _bridge import point_grid, parallel_grid


@pytest.fixture(scope="module")
def ct5():
    return build_telescope(load_packed_config("CT5"), I.MCIntegrator(115), I.random.key(0))


@pytest.mark.parametrize("sensor_idx", [0, 2])
def test_config2_ct5_4096_sources_properties(ct5, sensor_idx):
    """Config 2: CT5, 64x64 point sources, 876 x 115 = 100 740 rays per source (4.13e8 rays)."""
    src = torch.from_numpy(point_grid(64, 1.5)).cuda()
    val = torch.linspace(0.5, 1.5, 4096, device="cuda")
    full = render(ct5, src, val, "point", sensor_idx)
    assert torch.isfinite(full).all() and float(full.min()) >= 0.0
    # additivity over a partition of the sources
    a = render(ct5, src[:1500], val[:1500], "point", sensor_idx)
    b = render(ct5, src[1500:], val[1500:], "point", sensor_idx)
    torch.testing.assert_close(a + b, full, rtol=2e-5, atol=2e-6 * float(full.max()))
    # linearity in the fluxes
    scaled = render(ct5, src, 3.0 * val, "point", sensor_idx)
    torch.testing.assert_close(scaled, 3.0 * full, rtol=2e-5, atol=2e-6 * float(full.max()))
    # permutation invariance
    perm = torch.randperm(4096, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    torch.testing.assert_close(render(ct5, src[perm], val[perm], "point", sensor_idx), full, rtol=2e-5,
                               atol=2e-6 * float(full.max()))
    # shadowing only removes light; CT5's masts + camera body shadow 5-15 % of the dish
    clear = render(ct5.clear_obstructions(), src, val, "point", sensor_idx)
    frac = 1.0 - float(full.sum()) / float(clear.sum())
    assert 0.03 < frac < 0.25
    # effective area per unit flux ~ 876 hexagons of 0.7 m^2 = 614 m^2, minus shadow and camera cut-off
    assert 0.5 * 614 < float(clear.sum()) / float(val.sum()) < 614


def test_config4_ct3_response_matrix_properties():
    """Config 4: CT3 + roughness 24", 64x64 parallel directions over 5.5 deg, M = 64 -> (4096, 960)."""
    tel = build_telescope(load_packed_config("CT3"), I.MCIntegrator(64), I.random.key(42)).apply_roughness(24)
    src = torch.from_numpy(parallel_grid(64, 5.5)).cuda()
    val = torch.ones(4096, device="cuda")
    M = render_response_matrix(tel, src, val, "parallel", 0)
    assert M.shape == (4096, 960) and torch.isfinite(M).all() and float(M.min()) >= 0.0
    img = render(tel, src, val, "parallel", 0)
    torch.testing.assert_close(M.sum(0), img, rtol=5e-5, atol=1e-5)
    # row i depends on source i only
    sub = render_response_matrix(tel, src[1000:1100], val[1000:1100], "parallel", 0)
    torch.testing.assert_close(sub, M[1000:1100], rtol=1e-6, atol=1e-9)
    # effective aperture plateau ~100 m^2 on axis (ResponseMatrix.ipynb cell 13), zero outside the camera
    area = M.sum(1).reshape(64, 64)
    centre = float(area[28:36, 28:36].mean())
    assert 85.0 < centre < 105.0
    assert float(area[0, 0]) == 0.0 and float(area[-1, -1]) == 0.0
    # per-pixel response peaks at a sizeable fraction of the dish (cell 19: ~66 m^2 for pixel 42)
    assert 30.0 < float(M.max()) < 105.0
