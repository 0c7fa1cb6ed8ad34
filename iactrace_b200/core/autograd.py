"""Autograd bridge: ``render`` as a ``torch.autograd.Function`` backed by the hand-written VJP
kernel (the analogue of a ``jax.custom_vjp`` around the reference's ``render``)."""
from __future__ import annotations

import torch

from .. import _native as N
from .._util import f32, contig


def _leaves(tel, sensor_idx):
    from .render import _get_stages
    stages = _get_stages(tel.mirror_groups)
    out = []
    for g in stages.get(0, []):
        out += [g.positions, g.rotations, g.perturbation_scale, g.weights]
    s = tel.sensors[sensor_idx]
    out += [s.position, s.rotation]
    for k in stages:                       # stage >= 1 mirror poses, in the flat order of IactScene.stages
        if k != 0:
            for g in stages[k]:
                out += [g.positions, g.rotations]
    return out


def with_leaves(tel, sensor_idx, new):
    """Functional copy of ``tel`` whose differentiable leaves (in ``_leaves`` order) are replaced by ``new``."""
    from .render import _get_stages
    from .._util import replace
    it = iter(new)
    swap = {}
    stages = _get_stages(tel.mirror_groups)
    for g in stages.get(0, []):
        swap[id(g)] = replace(g, positions=next(it), rotations=next(it), perturbation_scale=next(it), weights=next(it))
    s = tel.sensors[sensor_idx]
    sensors = list(tel.sensors)
    sensors[sensor_idx] = replace(s, position=next(it), rotation=next(it))
    for k in stages:
        if k != 0:
            for g in stages[k]:
                swap[id(g)] = replace(g, positions=next(it), rotations=next(it))
    return replace(tel, mirror_groups=[swap.get(id(g), g) for g in tel.mirror_groups], sensors=sensors)


def needs_grad(tel, sources, values, sensor_idx) -> bool:
    if not torch.is_grad_enabled():
        return False
    ts = _leaves(tel, sensor_idx) + [t for t in (sources, values) if isinstance(t, torch.Tensor)]
    return any(t.requires_grad for t in ts)


class _Render(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tel, source_type, sensor_idx, src, val, *leaves):
        from .render import build_scene, _stype
        ctx.tel, ctx.source_type, ctx.sensor_idx = tel, source_type, sensor_idx
        # the leaves are saved too: autograd then refuses a backward after one of them was edited in place
        # (the backward pass rebuilds the scene from the telescope's current tensors)
        ctx.save_for_backward(src, val, *leaves)
        with torch.no_grad():
            keep = []
            sc, sensor = build_scene(tel, sensor_idx, keep)
            out = torch.empty(sensor.get_accumulator_shape(), dtype=torch.float32, device=src.device)
            if sc is None:
                return out.zero_()
            N.check(N.lib().iact_render(sc, N.ptr(src), N.ptr(val), src.shape[0], _stype(source_type), N.ptr(out),
                                        N.stream_ptr()), "render")
        return out

    @staticmethod
    def backward(ctx, g_img):
        from .render import build_scene, _stype, _get_stages
        tel, sensor_idx = ctx.tel, ctx.sensor_idx
        src, val = ctx.saved_tensors[:2]
        stages = _get_stages(tel.mirror_groups)
        g_img = contig(g_img.to(torch.float32))
        dev = src.device
        need = ctx.needs_input_grad          # (tel, source_type, sensor_idx, src, val, *leaves)
        zeros = lambda *shape: torch.zeros(shape, device=dev)
        g_src = torch.zeros_like(src) if need[3] else None
        g_val = torch.zeros_like(val) if need[4] else None
        stage0 = stages.get(0, [])
        i_sens = 5 + 4 * len(stage0)         # index of the sensor position in `need`
        # only the gradients somebody asked for are computed: the kernel has a lean instantiation for "facet poses
        # only" (config 5: an alignment fit w.r.t. rotations) that drops the sensor / source / scale / weight adjoints
        g_spos = zeros(3) if need[i_sens] else None
        g_srot = zeros(3) if need[i_sens + 1] else None
        later = [g for k in stages if k != 0 for g in stages[k]]
        n2 = sum(len(g) for g in later)
        want_stage = n2 > 0 and any(need[i_sens + 2:])
        g_mpos = zeros(n2, 3) if want_stage else None
        g_mrot = zeros(n2, 3) if want_stage else None
        grads = []
        keep = []
        sc, _ = build_scene(tel, sensor_idx, keep)
        off = 0
        li = 5                               # index of this group's first leaf in `need`
        for g in stage0:
            F, M = len(g), g.points.shape[1]
            # pose gradients share their accumulators in the kernel: both or neither
            pose = need[li] or need[li + 1]
            gp = zeros(F, 3) if pose else None
            gr = zeros(F, 3) if pose else None
            gs = zeros(F) if need[li + 2] else None
            # per-sample weight gradients cost one global atomic per ray: only when asked for
            gw = zeros(F, M, 1) if need[li + 3] else None
            li += 4
            if sc is not None and F * M:
                fa = g._facets_struct(keep)
                sub = N.IactScene.from_buffer_copy(sc)
                # the VJP kernel walks one group's facets: point the scene at that slice of the tables
                sub.n_facets = F
                sub.world = sc.world + off * M * 8 * 4
                sub.bounds = sc.bounds + off * 4 * 4
                if sc.chunk_bounds:
                    sub.chunk_bounds = sc.chunk_bounds + off * ((M + 31) // 32) * 4 * 4
                gr_struct = N.IactGrads(N.ptr(gp), N.ptr(gr), N.ptr(gs), N.ptr(gw), N.ptr(g_val), N.ptr(g_src),
                                        N.ptr(g_spos), N.ptr(g_srot), N.ptr(g_mpos), N.ptr(g_mrot))
                N.check(N.lib().iact_render_vjp(sub, fa, N.ptr(src), N.ptr(val), src.shape[0],
                                                _stype(ctx.source_type), N.ptr(g_img), gr_struct, N.stream_ptr()),
                        "render_vjp")
            grads += [gp, gr, gs, gw]
            off += F
        stage_grads, off2 = [], 0
        for g in later:
            stage_grads += ([g_mpos[off2:off2 + len(g)], g_mrot[off2:off2 + len(g)]] if want_stage else [None, None])
            off2 += len(g)
        return (None, None, None, g_src, g_val, *grads, g_spos, g_srot, *stage_grads)


def render_with_grad(tel, sources, values, source_type, sensor_idx):
    from .render import scene_device
    dev = scene_device(tel)
    src = f32(sources, dev).reshape(-1, 3).contiguous()
    val = f32(values, dev).reshape(-1).contiguous()
    return _Render.apply(tel, source_type, sensor_idx, src, val, *_leaves(tel, sensor_idx))
