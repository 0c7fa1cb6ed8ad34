"""Multi-GPU execution: one process per GPU (``torchrun``), sources sharded across ranks.

Rays are independent, so the path shards without any data-path exchange except the final pixel
sum of a single image: each rank renders its contiguous slice of the sources into a partial
image and one NCCL all-reduce (sum, float32, ``n_pixels`` floats) over NVLink combines them.
Response-matrix rows are source-owned, so ranks just write their own row block (optional
all-gather).  Sample tables are regenerated identically on every rank from the same key (the
sampler is counter-based), so nothing is broadcast.
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


def render_sharded(tel, sources, values, source_type="point", sensor_idx: int = 0, group=None, _render=None):
    """``render`` over all ranks of ``group``: every rank passes the FULL source list and gets the
    full image back (partial image + all-reduce)."""
    from .core.render import render
    rank, world = _world(group)
    a, b = shard_bounds(len(sources), rank, world)
    img = (_render or render)(tel, sources[a:b], values[a:b], source_type, sensor_idx)
    if world > 1:
        dist.all_reduce(img, op=dist.ReduceOp.SUM, group=group)
    return img


def response_matrix_sharded(tel, sources, values, source_type="point", sensor_idx: int = 0, group=None,
                            gather: bool = False, _render=None):
    """Row-sharded response matrix.  Returns this rank's row block and its (start, stop); with
    ``gather=True`` every rank receives the full (S, n_pixels) matrix instead."""
    from .core.render import render_response_matrix
    rank, world = _world(group)
    a, b = shard_bounds(len(sources), rank, world)
    rows = (_render or render_response_matrix)(tel, sources[a:b], values[a:b], source_type, sensor_idx)
    if not gather or world == 1:
        return rows, (a, b)
    sizes = [shard_bounds(len(sources), r, world) for r in range(world)]
    width = max(e - s for s, e in sizes)
    pad = torch.zeros((width, rows.shape[1]), dtype=rows.dtype, device=rows.device)
    pad[: rows.shape[0]] = rows
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    full = torch.cat([p[: e - s] for p, (s, e) in zip(parts, sizes)], dim=0)
    return full, (0, len(sources))
