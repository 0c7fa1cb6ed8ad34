"""Two real GPUs over NCCL: the source-sharded render + all-reduce equals the single-GPU render, the
row-sharded response matrix tiles the full matrix, and the sharded gradient (per-rank VJP + one all-reduce of
the packed leaf gradients) equals the single-GPU gradient.  Skipped with fewer than two devices."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import iactrace_b200 as I
        from iactrace_b200.io import build_telescope, load_packed_config
        from iactrace_b200.parallel import render_sharded, response_matrix_sharded
        from iactrace_b200.workloads import point_grid
        cfg = load_packed_config("CT3")
        cfg = dict(cfg, mirrors=cfg["mirrors"][::6])
        tel = build_telescope(cfg, I.MCIntegrator(32), I.random.key(0))      # same key on every rank -> same samples
        src = torch.from_numpy(point_grid(7, 1.0)).cuda()
        val = torch.linspace(0.5, 1.5, len(src), device="cuda")
        img = render_sharded(tel, src, val, "point", 0)
        full, _ = response_matrix_sharded(tel, src, val, "point", 0, gather=True)
        rows, (a, b) = response_matrix_sharded(tel, src, val, "point", 0)
        # sharded gradient: every rank ends up with the full gradient of a replicated loss
        g = tel.mirror_groups[0]
        g.rotations.requires_grad_(True)
        g.perturbation_scale.requires_grad_(True)
        vg = val.clone().requires_grad_(True)
        G = torch.linspace(-1.0, 1.0, 960, device="cuda")
        gi = render_sharded(tel, src, vg, "point", 0)
        (gi * G).sum().backward()
        torch.cuda.synchronize()
        q.put((rank, img.cpu().numpy(), full.cpu().numpy(), rows.cpu().numpy(), a, b,
               tel.mirror_groups[0].points[:2, :3].cpu().numpy(), gi.detach().cpu().numpy(),
               g.rotations.grad.cpu().numpy(), g.perturbation_scale.grad.cpu().numpy(), vg.grad.cpu().numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_sharded_render_matches_single_gpu():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    import iactrace_b200 as I
    from iactrace_b200.core import render, render_response_matrix
    from iactrace_b200.io import build_telescope, load_packed_config
    from iactrace_b200.workloads import point_grid
    cfg = load_packed_config("CT3")
    cfg = dict(cfg, mirrors=cfg["mirrors"][::6])
    tel = build_telescope(cfg, I.MCIntegrator(32), I.random.key(0))
    src = torch.from_numpy(point_grid(7, 1.0)).cuda()
    val = torch.linspace(0.5, 1.5, len(src), device="cuda")
    want = render(tel, src, val, "point", 0).cpu().numpy()
    want_m = render_response_matrix(tel, src, val, "point", 0).cpu().numpy()
    g = tel.mirror_groups[0]
    g.rotations.requires_grad_(True)
    g.perturbation_scale.requires_grad_(True)
    vg = val.clone().requires_grad_(True)
    (render(tel, src, vg, "point", 0) * torch.linspace(-1.0, 1.0, 960, device="cuda")).sum().backward()
    want_g = [t.grad.cpu().numpy() for t in (g.rotations, g.perturbation_scale, vg)]
    assert np.abs(want_g[0]).max() > 0 and np.abs(want_g[2]).max() > 0
    assert np.array_equal(res[0][6], res[1][6])                   # identical samples on both ranks
    for rank, img, full, rows, a, b, _, gimg, g_rot, g_scale, g_val in res:
        np.testing.assert_allclose(img, want, rtol=2e-5, atol=1e-7)
        np.testing.assert_allclose(gimg, want, rtol=2e-5, atol=1e-7)
        np.testing.assert_allclose(full, want_m, rtol=2e-6, atol=1e-9)
        np.testing.assert_allclose(rows, want_m[a:b], rtol=2e-6, atol=1e-9)
        for got, w in zip((g_rot, g_scale, g_val), want_g):
            np.testing.assert_allclose(got, w, rtol=1e-4, atol=2e-5 * np.abs(w).max())
    assert (res[0][4], res[0][5], res[1][4], res[1][5]) == (0, 25, 25, 49)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_render_follows_the_telescope_device():
    """Multi-GPU in ONE process: the kernels dereference the telescope's tensors, so inputs, outputs and the launch
    follow the device the telescope lives on, whatever the current device is (core/render.py::scene_device)."""
    import iactrace_b200 as I
    from iactrace_b200.core import render
    from iactrace_b200.io import build_telescope, load_packed_config
    from iactrace_b200.workloads import point_grid
    cfg = load_packed_config("CT3")
    cfg = dict(cfg, mirrors=cfg["mirrors"][::12])
    src, val = point_grid(3, 1.0), np.ones(9, np.float32)
    torch.cuda.set_device(0)
    tel0 = build_telescope(cfg, I.MCIntegrator(16), I.random.key(0))
    want = render(tel0, src, val, "point", 0)
    with torch.cuda.device(1):
        tel1 = build_telescope(cfg, I.MCIntegrator(16), I.random.key(0))
    assert tel1.mirror_groups[0].points.device.index == 1
    assert torch.cuda.current_device() == 0
    got = render(tel1, src, val, "point", 0)                       # current device 0, telescope on device 1
    assert got.device.index == 1
    torch.testing.assert_close(got.cpu(), want.cpu(), rtol=2e-5, atol=1e-7)
    got2 = render(tel1, torch.from_numpy(src).cuda(0), val, "point", 0)    # sources on the other device are moved
    torch.testing.assert_close(got2.cpu(), want.cpu(), rtol=2e-5, atol=1e-7)
    # a telescope spread over two devices is refused with a clear error
    from iactrace_b200._util import replace
    bad = replace(tel1, sensors=tel0.sensors)
    with pytest.raises(ValueError, match="several devices"):
        render(bad, src, val, "point", 0)
