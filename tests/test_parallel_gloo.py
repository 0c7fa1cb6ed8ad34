"""Multi-process host logic of the sharded render (world_size 2, gloo, CPU): the shard bounds
partition the sources, the all-reduce sums partial images, the gathered matrix has every row once."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from iactrace_b200.parallel import shard_bounds, render_sharded, response_matrix_sharded


def test_shard_bounds_partition():
    for n in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _fake_render(tel, sources, values, source_type, sensor_idx):
    # linear in the sources, like the real render: pixel p receives sum_i values_i * (p+1) * sources_i[0]
    p = torch.arange(1, 6, dtype=torch.float32)
    return (values[:, None] * sources[:, :1] * p[None, :]).sum(0)


def _fake_matrix(tel, sources, values, source_type, sensor_idx):
    p = torch.arange(1, 6, dtype=torch.float32)
    return values[:, None] * sources[:, :1] * p[None, :]


def _worker(rank, world, port, n_src, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        src = torch.rand((n_src, 3), generator=g)
        val = torch.rand((n_src,), generator=g)
        img = render_sharded(None, src, val, _render=_fake_render)
        full, span = response_matrix_sharded(None, src, val, gather=True, _render=_fake_matrix)
        rows, (a, b) = response_matrix_sharded(None, src, val, _render=_fake_matrix)
        q.put((rank, img.numpy(), full.numpy(), rows.numpy(), a, b))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_src", [7, 64])
def test_two_rank_gloo_matches_single_process(n_src):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_src, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    src = torch.rand((n_src, 3), generator=g)
    val = torch.rand((n_src,), generator=g)
    want_img = _fake_render(None, src, val, "point", 0).numpy()
    want_mat = _fake_matrix(None, src, val, "point", 0).numpy()
    for rank, img, full, rows, a, b in res:
        np.testing.assert_allclose(img, want_img, rtol=1e-5)
        np.testing.assert_allclose(full, want_mat, rtol=1e-6)
        np.testing.assert_allclose(rows, want_mat[a:b], rtol=1e-6)
    assert res[0][5] == res[1][4]        # the row blocks tile the matrix


# ---------------------------------------------------------------- sharded gradient
def _cpu_telescope():
    import iactrace_b200 as I
    from iactrace_b200.core import AsphericSurface
    from iactrace_b200.telescope.mirrors import AsphericDiskMirrorGroup
    g = torch.Generator().manual_seed(1)
    grp = AsphericDiskMirrorGroup(torch.rand((4, 3), generator=g), torch.rand((4, 3), generator=g),
                                  AsphericSurface(0.03, 0.0, []), torch.full((4,), 0.3))
    sens = I.SquareSensor([0.0, 0.0, 15.0], [0.0, 0.0, 0.0], 4, 2, (-1, 1, -1, 1))
    return I.Telescope([grp], [], [sens])


def _fake_diff_render(tel, sources, values, source_type, sensor_idx):
    # differentiable stand-in with the real render's structure: a sum over sources of a nonlinear function of the
    # facet leaves, the sensor pose and the source
    g = tel.mirror_groups[0]
    s = tel.sensors[sensor_idx]
    p = torch.arange(1, 6, dtype=torch.float32)
    per_src = torch.sin(sources @ g.rotations.T + g.positions.sum(1)[None, :]).sum(1) * values      # (S,)
    return (per_src[:, None] * p[None, :]).sum(0) * (1.0 + s.position[2] * 0.01) + 0.0 * g.perturbation_scale.sum()


def _leaf_setup(n_src):
    tel = _cpu_telescope()
    g = tel.mirror_groups[0]
    g.rotations.requires_grad_(True)
    g.positions.requires_grad_(True)
    tel.sensors[0].position.requires_grad_(True)
    gen = torch.Generator().manual_seed(0)
    src = torch.rand((n_src, 3), generator=gen).requires_grad_(True)
    val = torch.rand((n_src,), generator=gen).requires_grad_(True)
    return tel, src, val


def _grad_worker(rank, world, port, n_src, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tel, src, val = _leaf_setup(n_src)
        img = render_sharded(tel, src, val, _render=_fake_diff_render)
        target = torch.linspace(0.0, 1.0, 5)
        (0.5 * ((img - target) ** 2).sum()).backward()
        g = tel.mirror_groups[0]
        q.put((rank, img.detach().numpy(), g.rotations.grad.numpy(), g.positions.grad.numpy(),
               tel.sensors[0].position.grad.numpy(), src.grad.numpy(), val.grad.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_src", [5, 16])
def test_two_rank_sharded_gradient_matches_single_process(n_src):
    """render_sharded under autograd: per-rank VJP of the rank's source slice + one all-reduce of the packed leaf
    gradients gives every rank the gradient a single process computes."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, n_src, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tel, src, val = _leaf_setup(n_src)
    img = _fake_diff_render(tel, src, val, "point", 0)
    (0.5 * ((img - torch.linspace(0.0, 1.0, 5)) ** 2).sum()).backward()
    g = tel.mirror_groups[0]
    want = (img.detach().numpy(), g.rotations.grad.numpy(), g.positions.grad.numpy(),
            tel.sensors[0].position.grad.numpy(), src.grad.numpy(), val.grad.numpy())
    for r in res:
        for got, w in zip(r[1:], want):
            np.testing.assert_allclose(got, w, rtol=2e-5, atol=1e-6)


def test_sharded_helpers_respect_return_numpy():
    from iactrace_b200 import config
    config.return_numpy = True
    try:
        src, val = torch.rand((6, 3)), torch.rand((6,))
        img = render_sharded(None, src, val, _render=_fake_render)
        rows, _ = response_matrix_sharded(None, src, val, _render=_fake_matrix)
        assert isinstance(img, np.ndarray) and isinstance(rows, np.ndarray)
    finally:
        config.return_numpy = False
