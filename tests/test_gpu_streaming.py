"""Large sample counts: the counter-based sampler regenerates any window of a facet's stream bit-identically
(``iact_sample_*_group_rows``), and ``render`` / ``render_response_matrix`` / the gradient give the same result
whether the samples are held in memory, walked in L2-sized windows, or streamed from the key alone
(reference: core/integrators.py:97-188 draws all n_samples at once)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import iactrace_b200 as I
from iactrace_b200 import config as Rm, random as R
from iactrace_b200.core import MCIntegrator, render, render_debug, render_response_matrix
from iactrace_b200.io import build_telescope, load_packed_config
from _bridge import subset_config, point_grid, parallel_grid


@pytest.mark.parametrize("scene,mode", [("CT3", "partitionable"), ("CT5", "partitionable"), ("CT5", "legacy"), ("CT3", "legacy")])
def test_any_window_of_the_stream_is_regenerated_bit_identically(scene, mode):
    cfg = subset_config(load_packed_config(scene), mirror_step=40)
    R.set_rng_mode(mode)
    try:
        full = build_telescope(cfg, MCIntegrator(1000, stream=False), R.key(7)).mirror_groups[0]
        key = R.split(R.key(7))[1]                                  # yaml_loader.py:76: key, subkey = split(key)
        for first, n in ((0, 1000), (0, 37), (481, 300), (999, 1), (64, 936)):
            w = MCIntegrator.sample_rows(full, key, first, n, 1000)
            for name in ("points", "normals", "perturbation_delta", "weights"):
                assert torch.equal(getattr(w, name), getattr(full, name)[:, first:first + n]), (name, first, n)
    finally:
        R.set_rng_mode(R.PARTITIONABLE)


def test_streamed_windowed_and_materialised_renders_agree():
    cfg = subset_config(load_packed_config("CT5"), mirror_step=7)
    src = point_grid(3, 1.5)
    val = np.linspace(0.5, 1.5, len(src)).astype(np.float32)
    held = build_telescope(cfg, MCIntegrator(1500, stream=False), R.key(3))
    streamed = build_telescope(cfg, MCIntegrator(1500, stream=True), R.key(3))
    assert streamed.mirror_groups[0].points.shape[1] == 0 and streamed.mirror_groups[0].sample_stream.n_samples == 1500
    old = Rm.window_table_bytes
    try:
        Rm.window_table_bytes = 1 << 40                              # one pass over the whole table
        want = {si: render(held, src, val, "point", si) for si in (0, 2)}
        want_m = render_response_matrix(held, src, val, "point", 0)
        Rm.window_table_bytes = len(held.mirror_groups[0]) * 32 * 400      # windows of <= 384 samples
        from iactrace_b200.core.streaming import window_plan
        assert len(window_plan(held)) == 4 and sum(n for _, n in window_plan(streamed)) == 1500
        for tel in (held, streamed):
            for si in (0, 2):
                got = render(tel, src, val, "point", si)
                torch.testing.assert_close(got, want[si], rtol=2e-5, atol=2e-6 * float(want[si].max()))
            torch.testing.assert_close(render_response_matrix(tel, src, val, "point", 0), want_m, rtol=2e-5,
                                       atol=2e-6 * float(want_m.max()))
        # every ray exactly once: total flux identical to float32 summation noise
        assert abs(float(render(streamed, src, val, "point", 0).sum()) - float(want[0].sum())) < 1e-5 * float(want[0].sum())
        with pytest.raises(NotImplementedError):
            render_debug(streamed, src, val, "point", 0)
    finally:
        Rm.window_table_bytes = old


def test_gradient_through_sample_windows():
    from iactrace_b200.sensors import DifferentiableHexagonalSensor
    cfg = subset_config(load_packed_config("CT5"), mirror_step=60)
    src, val = point_grid(2, 1.0), np.ones(4, np.float32)
    grads = []
    old = Rm.window_table_bytes
    try:
        for stream, wbytes in ((False, 1 << 40), (False, 15 * 32 * 300), (True, 15 * 32 * 300)):
            Rm.window_table_bytes = wbytes
            tel = build_telescope(cfg, MCIntegrator(700, stream=stream), R.key(0)).apply_roughness(30)
            hard = tel.sensors[0]
            tel = tel.replace_sensor(DifferentiableHexagonalSensor(hard.position, hard.rotation, hard.hex_centers, 0.5, 1,
                                                                   grid=hard.grid_constants()), 0)
            g = tel.mirror_groups[0]
            g.rotations.requires_grad_(True)
            g.positions.requires_grad_(True)
            img = render(tel, src, val, "point", 0)
            (img * torch.linspace(-1, 1, img.numel(), device="cuda")).sum().backward()
            grads.append((g.rotations.grad.clone(), g.positions.grad.clone()))
        for a, b in zip(grads[0], grads[1]):
            torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5 * float(b.abs().max()))
        for a, b in zip(grads[0], grads[2]):
            torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5 * float(b.abs().max()))
    finally:
        Rm.window_table_bytes = old


def test_window_rays_are_bit_identical_to_the_table_path():
    """Per ray: the rays traced from a regenerated sample window are bit-identical to the same rows of the render_debug
    output of the fully materialised draw (the windows are what `render` sums when the samples are streamed)."""
    from iactrace_b200.core.streaming import window_plan, iter_windows
    cfg = subset_config(load_packed_config("CT5"), mirror_step=25)
    src = point_grid(2, 1.5)
    val = np.linspace(0.5, 1.5, len(src)).astype(np.float32)
    M = 700
    held = build_telescope(cfg, MCIntegrator(M, stream=False), R.key(11))
    streamed = build_telescope(cfg, MCIntegrator(M, stream=True), R.key(11))
    F, S = len(held.mirror_groups[0]), len(src)
    old = Rm.window_table_bytes
    try:
        Rm.window_table_bytes = 1 << 40
        xy, v, pix = render_debug(held, src, val, "point", 0, return_pixels=True)
        xy, v, pix = xy.reshape(F, S, M, 2), v.reshape(F, S, M), pix.reshape(F, S, M)
        Rm.window_table_bytes = F * 32 * 260                                  # windows of 256 samples
        plan = window_plan(streamed)
        assert len(plan) == 3
        seen = 0
        for (first, n), tw in zip(plan, iter_windows(streamed, plan)):
            Rm.window_table_bytes = 1 << 40                                   # the window itself is traced in one go
            wxy, wv, wpix = render_debug(tw, src, val, "point", 0, return_pixels=True)
            Rm.window_table_bytes = F * 32 * 260
            assert torch.equal(wxy.reshape(F, S, n, 2), xy[:, :, first:first + n])
            assert torch.equal(wv.reshape(F, S, n), v[:, :, first:first + n])
            assert torch.equal(wpix.reshape(F, S, n), pix[:, :, first:first + n])
            seen += n
        assert seen == M and float((v != 0).float().mean()) > 0.5
    finally:
        Rm.window_table_bytes = old
