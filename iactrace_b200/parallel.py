"""Multi-GPU execution: one process per GPU (``torchrun``), sources sharded across ranks.

Rays are independent, so the path shards without any data-path exchange except the final pixel
sum of a single image: each rank renders its contiguous slice of the sources into a partial
image and one NCCL all-reduce (sum, float32, ``n_pixels`` floats) over NVLink combines them.
Response-matrix rows are source-owned, so ranks just write their own row block (optional
all-gather).  Sample tables are regenerated identically on every rank from the same key (the
sampler is counter-based), so nothing is broadcast.

Gradients: ``render_sharded`` is differentiable.  Every rank back-propagates its
own source slice through the VJP kernel; the per-rank partial gradients of the telescope leaves
(positions / rotations / perturbation_scale / weights, sensor pose, stage >= 1 poses -- 876 x 6
floats on CT5) and of ``sources`` / ``values`` are packed into one flat buffer and summed with ONE
all-reduce, so every rank ends up with the full gradient, as it would on a single GPU
(reference entry of the gradient: ``telescope/mirrors.py:64-79``, ``telescope/operations.py:161-198``).
The loss is assumed replicated: every rank computes it from the same all-reduced image.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced slice [start, stop) of ``n`` sources owned by ``rank``."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(int(n), world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class _AllReduceImage(torch.autograd.Function):
    """Sum of the per-rank partial images.  Backward: the loss is replicated, so each rank's cotangent of the
    full image is also the cotangent of its partial image (d full / d partial = identity)."""

    @staticmethod
    def forward(ctx, img, group):
        out = img.detach().clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        return g, None


class _SyncGrads(torch.autograd.Function):
    """Identity on a list of tensors; backward packs their (per-rank partial) gradients into one flat buffer,
    all-reduces it once and hands the sums back -- every rank receives the full gradient."""

    @staticmethod
    def forward(ctx, group, *ts):
        ctx.group = group
        ctx.shapes = [t.shape for t in ts]
        return tuple(t.view_as(t) for t in ts)

    @staticmethod
    def backward(ctx, *gs):
        like = next(g for g in gs if g is not None)
        flat = torch.cat([(g if g is not None else torch.zeros(s, dtype=like.dtype, device=like.device)).reshape(-1).to(like.dtype)
                          for g, s in zip(gs, ctx.shapes)])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=ctx.group)
        out, off = [], 0
        for s in ctx.shapes:
            n = int(torch.Size(s).numel())
            out.append(flat[off:off + n].reshape(s))
            off += n
        return (None, *out)


def _tensor_render(tel, sources, values, source_type, sensor_idx):
    """``render`` returning a tensor whatever ``config.return_numpy`` says (collectives need tensors)."""
    from . import config
    from .core.render import render
    saved, config.return_numpy = config.return_numpy, False
    try:
        return render(tel, sources, values, source_type, sensor_idx)
    finally:
        config.return_numpy = saved


def _finish(t):
    from . import config
    if isinstance(t, torch.Tensor) and config.return_numpy and not t.requires_grad:
        return t.cpu().numpy()
    return t


def render_sharded(tel, sources, values, source_type="point", sensor_idx: int = 0, group=None, _render=None):
    """``render`` over all ranks of ``group``: every rank passes the FULL source list and gets the
    full image back (partial image + all-reduce).  Differentiable w.r.t. the telescope leaves, ``sources``
    and ``values`` (module docstring): after ``loss.backward()`` every rank holds the full gradient."""
    rank, world = _world(group)
    a, b = shard_bounds(len(sources), rank, world)
    fn = _render or _tensor_render
    if world == 1:
        return _finish(fn(tel, sources[a:b], values[a:b], source_type, sensor_idx))
    from .core.autograd import _leaves, with_leaves
    leaves = _leaves(tel, sensor_idx) if tel is not None else []
    extra = [t for t in (sources, values) if isinstance(t, torch.Tensor)]
    want = torch.is_grad_enabled() and any(t.requires_grad for t in leaves + extra)
    if not want:
        img = fn(tel, sources[a:b], values[a:b], source_type, sensor_idx)
        img = img.detach().clone()
        dist.all_reduce(img, op=dist.ReduceOp.SUM, group=group)
        return _finish(img)
    idx = [i for i, t in enumerate(leaves) if t.requires_grad]
    sync_in = [leaves[i] for i in idx] + [t for t in extra if t.requires_grad]
    synced = list(_SyncGrads.apply(group, *sync_in))
    new_leaves = list(leaves)
    for i, t in zip(idx, synced):
        new_leaves[i] = t
    rest = synced[len(idx):]
    if isinstance(sources, torch.Tensor) and sources.requires_grad:
        sources = rest.pop(0)
    if isinstance(values, torch.Tensor) and values.requires_grad:
        values = rest.pop(0)
    tel2 = with_leaves(tel, sensor_idx, new_leaves)
    img = fn(tel2, sources[a:b], values[a:b], source_type, sensor_idx)
    return _AllReduceImage.apply(img, group)


def response_matrix_sharded(tel, sources, values, source_type="point", sensor_idx: int = 0, group=None,
                            gather: bool = False, _render=None):
    """Row-sharded response matrix (forward only).  Returns this rank's row block and its (start, stop); with
    ``gather=True`` every rank receives the full (S, n_pixels) matrix instead."""
    from . import config
    from .core.render import render_response_matrix
    rank, world = _world(group)
    a, b = shard_bounds(len(sources), rank, world)
    saved, config.return_numpy = config.return_numpy, False
    try:
        rows = (_render or render_response_matrix)(tel, sources[a:b], values[a:b], source_type, sensor_idx)
    finally:
        config.return_numpy = saved
    if not gather or world == 1:
        return _finish(rows), (a, b)
    sizes = [shard_bounds(len(sources), r, world) for r in range(world)]
    width = max(e - s for s, e in sizes)
    pad = torch.zeros((width, rows.shape[1]), dtype=rows.dtype, device=rows.device)
    pad[: rows.shape[0]] = rows
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    full = torch.cat([p[: e - s] for p, (s, e) in zip(parts, sizes)], dim=0)
    return _finish(full), (0, len(sources))
