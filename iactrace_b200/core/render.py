"""Render drivers (mirror of reference ``iactrace/core/render.py:174-324``).

``render`` / ``render_debug`` / ``render_response_matrix`` pack the Telescope into the flat
``IactScene`` descriptor of the C ABI and launch the trace kernel.  Inputs may be NumPy arrays,
torch tensors (CPU or CUDA) or anything ``torch.as_tensor`` understands; outputs are float32 CUDA
tensors written by the kernel on the current CUDA stream (no hidden synchronisation).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _native as N
from .. import config
from .._util import f32, contig


def _get_stages(mirror_groups):
    """Group mirror groups by optical stage, ascending (``render.py:12-18``)."""
    by = {}
    for g in mirror_groups:
        by.setdefault(g.optical_stage, []).append(g)
    return dict(sorted(by.items()))


def _tensor_sig(ts):
    return tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in ts)


def _world_tables(tel, stage0, later_stages=False):
    """Packed world-frame sample table (F,M,8) + facet bounding spheres, cached on the Telescope."""
    srcs = []
    for g in stage0:
        srcs += [g.positions, g.rotations, g.perturbation_scale, g.points, g.normals, g.perturbation_delta, g.weights]
    M0 = stage0[0].points.shape[1]
    n_obs = sum(len(g) for g in (tel.obstruction_groups or []))
    # with optical stages >= 1 the leg towards the next mirror is culled per 32-ray run from the rays
    # themselves (occluded_leg_culled), which needs spatially compact runs whatever the primitive count
    binned = (bool(config.bin_samples_min) and M0 >= config.bin_samples_min
              and (n_obs >= config.bin_obstructions_min or (later_stages and n_obs >= 1)))
    sig = (_tensor_sig(srcs), binned)
    hit = tel._cache.get("world")
    if hit is not None and hit[0] == sig:
        return hit[1], hit[2], hit[3]
    M = stage0[0].points.shape[1]
    for g in stage0:
        if g.points.shape[1] != M:
            raise ValueError("all stage-0 mirror groups must carry the same number of samples")
    F = sum(len(g) for g in stage0)
    dev = stage0[0].points.device
    world = torch.empty((F, M, 8), dtype=torch.float32, device=dev)
    bounds = torch.empty((F, 4), dtype=torch.float32, device=dev)
    chunks = torch.empty((F, (M + 31) // 32, N.RUN_BOUND_FLOATS), dtype=torch.float32, device=dev) if binned else None
    grid_side = max(1, min(16, int(round(math.sqrt(M / 32.0)))))
    off = 0
    for g in stage0:
        keep = []
        if len(g) * M:
            fa = g._facets_struct(keep)
            if binned:
                N.check(N.lib().iact_transform_to_world_binned(fa, off, grid_side, N.ptr(world), N.ptr(bounds),
                                                               N.ptr(chunks), N.stream_ptr()), "transform_to_world_binned")
            else:
                N.check(N.lib().iact_transform_to_world(fa, off, N.ptr(world), N.ptr(bounds), N.stream_ptr()),
                        "transform_to_world")
        off += len(g)
    tel._cache["world"] = (sig, world, bounds, chunks)
    return world, bounds, chunks


def _scalar(v) -> float:
    """Surface scalars are Python floats in the reference (surfaces.py:11-12); a fit may hold them as 0-dim tensors."""
    return float(v.detach()) if isinstance(v, torch.Tensor) else float(v)


def _stage_sig(groups):
    ts = []
    for g in groups:
        ts += [g.positions, g.rotations, g.offsets, g.radii if g.kind == "disk" else g.vertices]
        ts += [v for v in (g.curvature, g.conic) if isinstance(v, torch.Tensor)]
    return (_tensor_sig(ts), tuple((None if isinstance(g.curvature, torch.Tensor) else g.curvature,
                                    None if isinstance(g.conic, torch.Tensor) else g.conic,
                                    tuple(float(a) for a in g.aspheric)) for g in groups))


def _stage_tables(tel, groups, dev):
    """Flat (n, IACT_MIRROR_REC) record table for one optical stage >= 1."""
    recs, verts = [], []
    for g in groups:
        pos = g.positions.detach().cpu().numpy()
        rot = g.rotations.detach().cpu().numpy()
        off = g.offsets.detach().cpu().numpy()
        asph = np.asarray(g.aspheric, np.float32)
        if len(asph) > N.MAX_ASPH:
            raise ValueError(f"at most {N.MAX_ASPH} aspheric terms are supported")
        for i in range(len(g)):
            r = np.zeros(N.MIRROR_REC, np.float32)
            r[0:3], r[3:6], r[6:8] = pos[i], rot[i], off[i]
            r[8], r[9], r[10] = _scalar(g.curvature), _scalar(g.conic), len(asph)
            r[11:11 + len(asph)] = asph
            if g.kind == "disk":
                r[19], r[20] = 0, float(g.radii[i])
            else:
                v = g.vertices[i].detach().cpu().numpy().reshape(-1, 2)
                r[19], r[21], r[22] = 1, len(v), sum(len(x) for x in verts)
                verts.append(v)
            recs.append(r)
    rec_t = f32(np.stack(recs), dev)
    vert_t = f32(np.concatenate(verts), dev) if verts else None
    return rec_t, vert_t


def _obstruction_tables(tel):
    """Per-type dense obstruction tables.  Several groups of one type are concatenated: the shadow
    mask is a product over groups of min-over-primitives (``render.py:37-41``), i.e. an any-hit."""
    from .obstructions import CylinderGroup, BoxGroup, SphereGroup, OrientedBoxGroup, TriangleGroup
    groups = tel.obstruction_groups or []
    # keyed on (storage, version) of every obstruction tensor: an in-place edit or a swapped group rebuilds the tables
    sig = tuple((type(g).__name__, _tensor_sig([v for v in vars(g).values() if isinstance(v, torch.Tensor)])) for g in groups)
    hit = tel._cache.get("obs")
    if hit is not None and hit[0] == sig:
        return hit[1]

    def cat(cls, names):
        gs = [g for g in groups if isinstance(g, cls)]
        if not gs:
            return None
        return [contig(torch.cat([getattr(g, n).detach().reshape(len(g), -1) for g in gs]).to(torch.float32)) for n in names]

    for g in groups:
        if not isinstance(g, (CylinderGroup, BoxGroup, SphereGroup, OrientedBoxGroup, TriangleGroup)):
            raise TypeError(f"Unknown obstruction group type: {type(g).__name__}")
    tabs = dict(cyl=cat(CylinderGroup, ("p1", "p2", "r")), box=cat(BoxGroup, ("p1", "p2")),
                sph=cat(SphereGroup, ("centers", "radii")),
                obox=cat(OrientedBoxGroup, ("centers", "half_extents", "rotations")),
                tri=cat(TriangleGroup, ("v0", "v1", "v2")))
    tel._cache["obs"] = (sig, tabs)
    return tabs


def build_scene(tel, sensor_idx: int, keep: list, cull: bool | None = None):
    """Pack ``tel`` into an ``IactScene``; returns (scene | None if no primary stage, sensor)."""
    sensor = tel.sensors[sensor_idx]
    stages = _get_stages(tel.mirror_groups)
    if not stages or 0 not in stages:
        return None, sensor
    sc = N.IactScene()
    world, bounds, chunks = _world_tables(tel, stages[0], later_stages=len(stages) > 1)
    keep += [world, bounds, chunks]
    sc.n_facets, sc.n_samples = world.shape[0], world.shape[1]
    sc.world, sc.bounds, sc.chunk_bounds = N.ptr(world), N.ptr(bounds), N.ptr(chunks)
    dev = world.device
    tabs = _obstruction_tables(tel)
    for name, fields in (("cyl", ("cyl_p1", "cyl_p2", "cyl_r")), ("box", ("box_p1", "box_p2")),
                         ("sph", ("sph_c", "sph_r")), ("obox", ("obox_c", "obox_h", "obox_R")),
                         ("tri", ("tri_v0", "tri_v1", "tri_v2"))):
        t = tabs[name]
        if t is None:
            continue
        t = [x.to(dev) for x in t]
        keep += t
        setattr(sc, "n_" + name, t[0].shape[0])
        for f, x in zip(fields, t):
            setattr(sc, f, N.ptr(x))
    later = [k for k in stages if k != 0]
    if len(later) > N.MAX_STAGES:
        raise ValueError(f"at most {N.MAX_STAGES} optical stages beyond the primary are supported")
    sc.n_stages = len(later)
    for i, k in enumerate(later):
        # keyed like the world table: the alignment-fit loop (render -> backward -> optimizer.step()) edits the
        # stage >= 1 poses in place, and the next render must see them
        ck, sig = ("stage", k), _stage_sig(stages[k])
        hit = tel._cache.get(ck)
        if hit is None or hit[0] != sig:
            hit = (sig,) + _stage_tables(tel, stages[k], dev)
            tel._cache[ck] = hit
        rec, verts = hit[1], hit[2]
        keep += [rec, verts]
        sc.stages[i].n_mirrors = rec.shape[0]
        sc.stages[i].records = N.ptr(rec)
        sc.stages[i].verts = N.ptr(verts)
    sc.sensor = sensor._struct(keep)
    sc.cull = int(config.cull_obstructions if cull is None else cull)
    return sc, sensor


def scene_device(tel) -> torch.device:
    """The CUDA device the telescope's tensors live on (the kernels dereference them, so inputs, outputs and the
    launch must be there too); the current device for an empty telescope."""
    N.require_cuda()
    devs = {g.positions.device for g in tel.mirror_groups}
    devs |= {v.device for g in (tel.obstruction_groups or []) for v in vars(g).values() if isinstance(v, torch.Tensor)}
    devs |= {s.position.device for s in tel.sensors if isinstance(getattr(s, "position", None), torch.Tensor)}
    if len(devs) > 1:
        raise ValueError(f"telescope tensors live on several devices: {sorted(map(str, devs))}")
    dev = devs.pop() if devs else torch.device("cuda", torch.cuda.current_device())
    if dev.type != "cuda":
        raise RuntimeError("iactrace_b200: the telescope was built without a CUDA device; the ray-tracing path has no CPU fallback")
    return dev


def _on_scene_device(fn):
    """Run a render entry point with the telescope's device current (kernel launches, allocations and the stream
    all follow the current device)."""
    import functools

    @functools.wraps(fn)
    def wrapped(tel, *a, **k):
        dev = scene_device(tel)
        if dev.index == torch.cuda.current_device():
            return fn(tel, *a, **k)
        with torch.cuda.device(dev):
            return fn(tel, *a, **k)
    return wrapped


def _inputs(tel, sources, values):
    dev = scene_device(tel)
    src = contig(f32(sources, dev).detach().reshape(-1, 3))
    val = contig(f32(values, dev).detach().reshape(-1))
    if src.shape[0] != val.shape[0]:
        raise ValueError(f"sources {tuple(src.shape)} and values {tuple(val.shape)} disagree")
    return src, val, dev


def _out(t):
    if isinstance(t, tuple):
        return tuple(_out(x) for x in t)
    return t.cpu().numpy() if config.return_numpy else t


def _stype(source_type) -> int:
    # any string other than 'point' is treated as parallel (render.py:129-133)
    return N.SOURCE_POINT if source_type == "point" else N.SOURCE_PARALLEL


@_on_scene_device
def render(tel, sources, values, source_type="point", sensor_idx: int = 0) -> torch.Tensor:
    """Render sources through the telescope onto a sensor -> image of the sensor's shape."""
    from .autograd import needs_grad, render_with_grad
    from .streaming import window_plan, iter_windows
    plan = window_plan(tel)
    if plan is not None:                      # large draw: L2-sized sample windows, window images added (streaming.py)
        saved, config.return_numpy = config.return_numpy, False
        try:
            out = None
            for tw in iter_windows(tel, plan):
                img = render(tw, sources, values, source_type, sensor_idx)
                out = img if out is None else out + img
        finally:
            config.return_numpy = saved
        return out if out.requires_grad else _out(out)
    if needs_grad(tel, sources, values, sensor_idx):
        return render_with_grad(tel, sources, values, source_type, sensor_idx)
    src, val, dev = _inputs(tel, sources, values)
    keep = []
    sc, sensor = build_scene(tel, sensor_idx, keep)
    out = torch.empty(sensor.get_accumulator_shape(), dtype=torch.float32, device=dev)
    if sc is None:
        return _out(out.zero_())
    N.check(N.lib().iact_render(sc, N.ptr(src), N.ptr(val), src.shape[0], _stype(source_type), N.ptr(out),
                                N.stream_ptr()), "render")
    return _out(out)


@_on_scene_device
def render_debug(tel, sources, values, source_type="point", sensor_idx: int = 0, return_pixels: bool = False):
    """Raw hits without accumulation -> (points (F*S*M,2), values (F*S*M,)), facet-major then source
    then sample (``render.py:223-268``).  ``return_pixels`` adds the int32 pixel id each ray is
    assigned by the (hard) sensor, -1 = rejected."""
    if any(getattr(g, "sample_stream", None) is not None for g in tel.mirror_groups):
        raise NotImplementedError("render_debug needs materialised samples: use MCIntegrator(n_samples, stream=False) "
                                  "(the per-ray output is n_facets x n_sources x n_samples rows)")
    src, val, dev = _inputs(tel, sources, values)
    keep = []
    sc, _ = build_scene(tel, sensor_idx, keep)
    if sc is None:
        e = (torch.zeros((0, 2), device=dev), torch.zeros((0,), device=dev))
        return e + (torch.zeros((0,), dtype=torch.int32, device=dev),) if return_pixels else e
    n = sc.n_facets * src.shape[0] * sc.n_samples
    xy = torch.empty((n, 2), dtype=torch.float32, device=dev)
    v = torch.empty((n,), dtype=torch.float32, device=dev)
    pix = torch.empty((n,), dtype=torch.int32, device=dev) if return_pixels else None
    N.check(N.lib().iact_render_debug(sc, N.ptr(src), N.ptr(val), src.shape[0], _stype(source_type), N.ptr(xy),
                                      N.ptr(v), N.ptr(pix), N.stream_ptr()), "render_debug")
    return _out((xy, v, pix) if return_pixels else (xy, v))


@_on_scene_device
def render_response_matrix(tel, sources, values, source_type="point", sensor_idx: int = 0) -> torch.Tensor:
    """Source-to-pixel response matrix (S, n_pixels): row i is the flattened image of source i alone."""
    from .streaming import window_plan, iter_windows
    plan = window_plan(tel)
    if plan is not None:
        saved, config.return_numpy = config.return_numpy, False
        try:
            out = None
            for tw in iter_windows(tel, plan):
                m = render_response_matrix(tw, sources, values, source_type, sensor_idx)
                out = m if out is None else out.add_(m)
        finally:
            config.return_numpy = saved
        return _out(out)
    src, val, dev = _inputs(tel, sources, values)
    keep = []
    sc, sensor = build_scene(tel, sensor_idx, keep)
    npix = math.prod(sensor.get_accumulator_shape())
    out = torch.empty((src.shape[0], npix), dtype=torch.float32, device=dev)
    if sc is None:
        return _out(out.zero_())
    N.check(N.lib().iact_response_matrix(sc, N.ptr(src), N.ptr(val), src.shape[0], _stype(source_type), N.ptr(out),
                                         N.stream_ptr()), "render_response_matrix")
    return _out(out)
