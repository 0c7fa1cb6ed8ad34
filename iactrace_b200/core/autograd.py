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
        out += [g.positions, g.rotations, g.perturbation_scale, g.weights, g.points, g.normals, g.perturbation_delta]
    s = tel.sensors[sensor_idx]
    out += [s.position, s.rotation]
    for k in stages:                       # stage >= 1 mirrors, in the flat order of IactScene.stages: poses, then the
        if k != 0:                         # surface parameters (floats unless the caller made them tensors)
            for g in stages[k]:
                out += [g.positions, g.rotations, g.offsets, g.curvature, g.conic]
    return out


N_LEAVES_STAGE0 = 7
N_LEAVES_LATER = 5


def with_leaves(tel, sensor_idx, new):
    """Functional copy of ``tel`` whose differentiable leaves (in ``_leaves`` order) are replaced by ``new``."""
    from .render import _get_stages
    from .._util import replace
    it = iter(new)
    swap = {}
    stages = _get_stages(tel.mirror_groups)
    for g in stages.get(0, []):
        swap[id(g)] = replace(g, positions=next(it), rotations=next(it), perturbation_scale=next(it), weights=next(it),
                              points=next(it), normals=next(it), perturbation_delta=next(it))
    s = tel.sensors[sensor_idx]
    sensors = list(tel.sensors)
    sensors[sensor_idx] = replace(s, position=next(it), rotation=next(it))
    for k in stages:
        if k != 0:
            for g in stages[k]:
                swap[id(g)] = replace(g, positions=next(it), rotations=next(it), offsets=next(it), curvature=next(it),
                                      conic=next(it))
    return replace(tel, mirror_groups=[swap.get(id(g), g) for g in tel.mirror_groups], sensors=sensors)


def needs_grad(tel, sources, values, sensor_idx) -> bool:
    if not torch.is_grad_enabled():
        return False
    ts = _leaves(tel, sensor_idx) + [sources, values]
    return any(isinstance(t, torch.Tensor) and t.requires_grad for t in ts)


class _Render(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tel, source_type, sensor_idx, src, val, *leaves):
        from .render import build_scene, _stype
        ctx.tel, ctx.source_type, ctx.sensor_idx = tel, source_type, sensor_idx
        # the leaves are saved too: autograd then refuses a backward after one of them was edited in place
        # (the backward pass rebuilds the scene from the telescope's current tensors)
        ctx.save_for_backward(src, val, *[t for t in leaves if isinstance(t, torch.Tensor)])
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
        i_sens = 5 + N_LEAVES_STAGE0 * len(stage0)         # index of the sensor position in `need`
        # only the gradients somebody asked for are computed: the kernel has a lean instantiation for "facet poses
        # only" (config 5: an alignment fit w.r.t. rotations) that drops the sensor / source / scale / weight adjoints
        g_spos = zeros(3) if need[i_sens] else None
        g_srot = zeros(3) if need[i_sens + 1] else None
        later = [g for k in stages if k != 0 for g in stages[k]]
        n2 = sum(len(g) for g in later)
        need_later = need[i_sens + 2:]
        want_pose = n2 > 0 and any(need_later[N_LEAVES_LATER * i + j] for i in range(len(later)) for j in (0, 1))
        want_surf = n2 > 0 and any(need_later[N_LEAVES_LATER * i + j] for i in range(len(later)) for j in (2, 3, 4))
        g_mpos = zeros(n2, 3) if want_pose else None
        g_mrot = zeros(n2, 3) if want_pose else None
        g_msurf = zeros(n2, 4) if want_surf else None       # d/d(curvature, conic, offset x, offset y) per mirror
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
            # per-sample gradients cost global atomics per ray: only when asked for
            gw = zeros(F, M, 1) if need[li + 3] else None
            gpts = zeros(F, M, 3) if need[li + 4] else None
            gnq = zeros(F, M, 3) if (need[li + 5] or need[li + 6]) else None
            if sc is not None and F * M:
                fa = g._facets_struct(keep)
                sub = N.IactScene.from_buffer_copy(sc)
                # the VJP kernel walks one group's facets: point the scene at that slice of the tables
                sub.n_facets = F
                sub.world = sc.world + off * M * 8 * 4
                sub.bounds = sc.bounds + off * 4 * 4
                if sc.chunk_bounds:
                    sub.chunk_bounds = sc.chunk_bounds + off * ((M + 31) // 32) * N.RUN_BOUND_FLOATS * 4
                gr_struct = N.IactGrads(N.ptr(gp), N.ptr(gr), N.ptr(gs), N.ptr(gw), N.ptr(g_val), N.ptr(g_src),
                                        N.ptr(g_spos), N.ptr(g_srot), N.ptr(g_mpos), N.ptr(g_mrot),
                                        N.ptr(gpts), N.ptr(gnq), N.ptr(g_msurf))
                N.check(N.lib().iact_render_vjp(sub, fa, N.ptr(src), N.ptr(val), src.shape[0],
                                                _stype(ctx.source_type), N.ptr(g_img), gr_struct, N.stream_ptr()),
                        "render_vjp")
            # nw = R (n_l + scale * delta_l): the kernel returns d/d(n_l + scale delta_l)
            g_nrm = gnq if need[li + 5] else None
            g_dlt = gnq * g.perturbation_scale.detach()[:, None, None] if need[li + 6] else None
            grads += [gp, gr, gs, gw, gpts, g_nrm, g_dlt]
            li += N_LEAVES_STAGE0
            off += F
        stage_grads, off2 = [], 0
        for i, g in enumerate(later):
            n = len(g)
            nd = need_later[N_LEAVES_LATER * i:N_LEAVES_LATER * (i + 1)]
            sl = slice(off2, off2 + n)
            stage_grads += [g_mpos[sl] if want_pose else None, g_mrot[sl] if want_pose else None,
                            g_msurf[sl, 2:4] if nd[2] else None,
                            g_msurf[sl, 0].sum() if nd[3] else None,          # curvature and conic are shared by the group
                            g_msurf[sl, 1].sum() if nd[4] else None]
            off2 += n
        return (None, None, None, g_src, g_val, *grads, g_spos, g_srot, *stage_grads)


def render_with_grad(tel, sources, values, source_type, sensor_idx):
    from .render import scene_device
    dev = scene_device(tel)
    src = f32(sources, dev).reshape(-1, 3).contiguous()
    val = f32(values, dev).reshape(-1).contiguous()
    return _Render.apply(tel, source_type, sensor_idx, src, val, *_leaves(tel, sensor_idx))
