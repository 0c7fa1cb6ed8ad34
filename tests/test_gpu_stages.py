"""Stage >= 1 (Cassegrain secondary: conic guess + 10 Newton steps), all five obstruction
primitives, and the soft sensors, against the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import iactrace_b200 as I
from iactrace_b200 import config as Rm
from iactrace_b200.core import (Box, Cylinder, OrientedBox, Sphere, Triangle, group_obstructions, render,
                                render_debug)
from iactrace_b200.io import build_telescope, load_packed_config
from iactrace_b200.sensors import DifferentiableHexagonalSensor, DifferentiableSquareSensor
from oracle import trace as otrace
from _bridge import to_oracle_scene, subset_config, cassegrain_config, parallel_grid


def _star_field(n, seed=42):
    rng = np.random.default_rng(seed)
    fov = np.deg2rad(3.0)
    d = np.stack([rng.uniform(-fov / 2, fov / 2, n), rng.uniform(-fov / 2, fov / 2, n), -np.ones(n)], 1)
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    f = (10 ** (-3 * rng.uniform(size=n))).astype(np.float32)
    return d, f


@pytest.mark.parametrize("with_obs", [False, True])
def test_cassegrain_rays_and_image(with_obs):
    """BASELINE config 3 geometry at oracle-sized sample counts."""
    tel = build_telescope(cassegrain_config(with_obs), I.MCIntegrator(48), I.random.key(0))
    src, val = _star_field(12)
    osc = to_oracle_scene(tel)
    xy, v = render_debug(tel, src, val, "parallel", 0)
    xy, v = xy.cpu().numpy(), v.cpu().numpy()
    oxy, ov = otrace.render_debug(osc, src, val, "parallel", 0, np.float64)
    lit, olit = v != 0, ov != 0
    assert (lit != olit).mean() < 5e-4
    both = lit & olit
    assert both.mean() > 0.3                      # a good share of rays make it through both mirrors
    np.testing.assert_allclose(v[both], ov[both], rtol=2e-5)
    # two reflections and a 6.45 m back focal lever: f32 coordinates within 5e-6 m of the f64 oracle
    assert np.abs(xy[both] - oxy[both]).max() < 5e-6
    # the f32 oracle sits at the same distance from f64 (this is rounding, not a kernel defect)
    oxy32, ov32 = otrace.render_debug(osc, src, val, "parallel", 0, np.float32)
    b32 = both & (ov32 != 0)
    assert np.abs(oxy32[b32] - oxy[b32]).max() < 2e-5
    img = render(tel, src, val, "parallel", 0).cpu().numpy()
    oimg = otrace.render(osc, src, val, "parallel", 0, np.float64)
    assert abs(img.sum() - oimg.sum()) < 2e-3 * oimg.sum()
    if with_obs:
        clear = render(tel.clear_obstructions(), src, val, "parallel", 0).cpu().numpy()
        assert img.sum() < clear.sum()


def test_all_obstruction_primitives():
    """cylinder, box, sphere, oriented box, triangle: shadow decisions equal the oracle's per ray."""
    cfg = subset_config(load_packed_config("CT3"), mirror_step=3)
    cfg = dict(cfg, obstructions=[])
    base = build_telescope(cfg, I.MCIntegrator(64), I.random.key(5))
    th = np.deg2rad(30.0)
    Rz = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]], np.float32)
    prims = {
        "cylinder": [Cylinder([-6, 0.2, 5.0], [6, -0.1, 5.5], 0.25), Cylinder([2, 2, 0.5], [3.5, 1.0, 9], 0.4)],
        "box": [Box([-1.5, -1.0, 7.0], [1.0, 2.0, 7.4]), Box([3.0, 3.0, 2.0], [2.0, 4.5, 6.0])],
        "sphere": [Sphere([0.5, -2.0, 6.0], 1.1), Sphere([-3.0, 3.0, 4.0], 0.6)],
        "oriented_box": [OrientedBox([1.0, 1.0, 6.0], [1.5, 0.4, 0.3], Rz), OrientedBox([-3, -2, 5], [0.5, 0.5, 2], Rz.T)],
        "triangle": [Triangle([-4, -4, 6], [4, -3, 6.5], [0, 3, 5.5]), Triangle([3, 3, 3], [5, 3, 3], [4, 5, 4])],
    }
    src = np.array([[0, 0, 1e10], [2e8, -1e8, 1e10]], np.float32)
    val = np.ones(2, np.float32)
    for name, plist in list(prims.items()) + [("mixed", sum(prims.values(), []))]:
        tel = I.Telescope(base.mirror_groups, group_obstructions(plist), base.sensors)
        xy, v = render_debug(tel, src, val, "point", 1)
        v = v.cpu().numpy()
        _, ov = otrace.render_debug(to_oracle_scene(tel), src, val, "point", 1, np.float64)
        flips = (v != 0) != (ov != 0)
        frac = (ov == 0).mean()
        assert 0.005 < frac < 0.9, (name, frac)
        assert flips.mean() < 3e-5, (name, flips.sum())
        np.testing.assert_allclose(v[~flips], ov[~flips], rtol=1e-5, atol=0)


def test_cylinder_caps_axis_parallel_rays_and_origins_inside():
    """The interval form of the cylinder test (csrc/iact_trace.cuh hit_cylinder) against the oracle's literal
    candidate tests (intersections.py:44-87): cap hits, rays parallel / nearly parallel to the axis (literal
    fallback below 1.8 deg), rays at moderate angles, and ray origins inside a cylinder."""
    cfg = subset_config(load_packed_config("CT3"), mirror_step=3)
    cfg = dict(cfg, obstructions=[])
    base = build_telescope(cfg, I.MCIntegrator(96), I.random.key(11))
    plist = [
        Cylinder([0.5, 0.5, 4.0], [0.5, 0.5, 6.0], 0.8),        # axis along z: on-axis rays run parallel to it, hit its caps
        Cylinder([-2.0, 1.0, 5.0], [-2.0, 1.0, 5.05], 1.2),     # flat disc: almost only cap hits
        Cylinder([3.0, -2.5, 3.0], [2.0, -1.5, 9.0], 0.5),      # inclined strut
        Cylinder([-3.5, -3.0, -1.0], [-3.5, -3.0, 3.0], 1.0),   # swallows some facets: ray origins inside the solid
        Cylinder([1.0, -4.0, 2.0], [1.05, -4.0, 8.0], 0.3),     # 0.5 deg off the z axis
    ]
    tel = I.Telescope(base.mirror_groups, group_obstructions(plist), base.sensors)
    ang = np.deg2rad([0.0, 0.5, 1.5, 2.5, 8.0, 25.0])
    src = np.stack([-np.sin(ang), np.zeros_like(ang), -np.cos(ang)], axis=1).astype(np.float32)
    val = np.ones(len(src), np.float32)
    for cull in (True, False):
        Rm.cull_obstructions = cull
        try:
            xy, v = render_debug(tel, src, val, "parallel", 1)
        finally:
            Rm.cull_obstructions = True
        v = v.cpu().numpy()
        _, ov = otrace.render_debug(to_oracle_scene(tel), src, val, "parallel", 1, np.float64)
        flips = (v != 0) != (ov != 0)
        frac = (ov == 0).mean()
        assert 0.02 < frac < 0.9, frac
        assert flips.mean() < 3e-5, (cull, flips.sum())
        np.testing.assert_allclose(v[~flips], ov[~flips], rtol=1e-5, atol=0)
    # point sources a few metres away: wide range of ray/axis angles within one beam
    psrc = np.array([[0.5, 0.5, 30.0], [6.0, -3.0, 12.0], [-2.0, 1.0, 20.0]], np.float32)
    _, v = render_debug(tel, psrc, np.ones(3, np.float32), "point", 1)
    _, ov = otrace.render_debug(to_oracle_scene(tel), psrc, np.ones(3, np.float32), "point", 1, np.float64)
    v = v.cpu().numpy()
    flips = (v != 0) != (ov != 0)
    assert flips.mean() < 3e-5, flips.sum()
    np.testing.assert_allclose(v[~flips], ov[~flips], rtol=1e-5, atol=0)


def test_soft_sensors_forward():
    cfg = subset_config(load_packed_config("CT5"), mirror_step=9)
    tel = build_telescope(cfg, I.MCIntegrator(32), I.random.key(0))
    hard = tel.sensors[0]
    soft = DifferentiableHexagonalSensor(hard.position, hard.rotation, hard.hex_centers, sigma=0.5, kernel_size=1,
                                         grid=hard.grid_constants())
    lid = tel.sensors[2]
    soft_sq = DifferentiableSquareSensor(lid.position, lid.rotation, 96, 96, (-0.75, 0.75, -0.715, 0.715),
                                         sigma=0.6, kernel_size=2)
    tel2 = tel.replace_sensor(soft, 0).replace_sensor(soft_sq, 2)
    src = np.array([[0, 0, 1e10], [1.5e8, 1e8, 1e10]], np.float32)
    val = np.array([1.0, 2.0], np.float32)
    from _parity import compare_soft_image
    for idx in (0, 2):
        img = render(tel2, src, val, "point", idx).cpu().numpy().astype(np.float64)
        # (1) splat arithmetic at 1e-4 against the float64 splat of the kernel's own hits; (2) end to end at 5e-3,
        # the distance at which the float32 ORACLE sits from the float64 one as well (Gaussian taps amplify the
        # float32 rounding of the hit coordinates; tests/_parity.py::compare_soft_image)
        st = compare_soft_image(tel2, src, val, "point", idx)
        print("soft sensor", idx, st)
        # same regime as the float32 oracle's own distance from float64 (measured on B200: kernel 2e-3, float32 oracle
        # 6e-4 .. 8e-4 on these 24 / 78 lit pixels; per-facet pose rounding does not average out inside a pixel)
        assert st["max_rel_err_vs_f64_oracle"] <= max(10.0 * st["f32_oracle_vs_f64_oracle"], 1e-3)
        # splatting conserves flux that lands well inside the camera
        hard_img = render(tel, src, val, "point", idx).cpu().numpy()
        assert abs(img.sum() - hard_img.sum()) < 0.05 * hard_img.sum()


def test_sensor_accumulate_entry_point():
    cfg = subset_config(load_packed_config("CT3"), n_mirrors=4)
    tel = build_telescope(cfg, I.MCIntegrator(128), I.random.key(0))
    src = np.array([[0, 0, 1e10]], np.float32)
    pts, vals = tel(src, np.ones(1, np.float32), debug=True)
    for idx in (0, 1):
        img = tel.sensors[idx].accumulate(pts[:, 0], pts[:, 1], vals)
        ref = tel(src, np.ones(1, np.float32), sensor_idx=idx)
        # debug hits were intersected with sensor 0's plane; sensor 1 sits 28 mm lower, so only idx 0 is identical
        if idx == 0:
            torch.testing.assert_close(img, ref, rtol=1e-5, atol=1e-8)
        assert img.shape == ref.shape


def test_stage_leg_culling_and_newton_exit_are_exact():
    """The leg towards the secondary is culled per 32-ray run (leg_masks / occluded_leg_culled) on a spatially binned
    sample table, and the Newton scan leaves early once t repeats: both must reproduce the brute-force,
    unbinned run bit for bit, per ray, with obstruction clouds of every type around the light path."""
    from iactrace_b200 import config as Rm
    rng = np.random.default_rng(7)
    base = build_telescope(cassegrain_config(True), I.MCIntegrator(320), I.random.key(3))
    src, val = _star_field(9, seed=1)
    obs = list(cassegrain_config(True)["obstructions"])
    prims = [Cylinder(o["p1"], o["p2"], o["r"]) if o["type"] == "cylinder" else
             Box(o["p1"], o["p2"]) if o["type"] == "box" else Sphere(o["center"], o["r"]) for o in obs]
    for _ in range(30):                                   # clutter between the primary and the secondary
        a = rng.uniform([-3, -3, 0.5], [3, 3, 6.5])
        prims.append(Cylinder(a, a + rng.normal(size=3) * rng.uniform(0.1, 1.5), float(rng.uniform(0.005, 0.05))))
    for _ in range(5):
        prims.append(Sphere(rng.uniform([-3, -3, 0.5], [3, 3, 6.5]), float(rng.uniform(0.02, 0.15))))
        a = rng.uniform([-3, -3, 0.5], [3, 3, 6.5])
        prims.append(Box(a, a + rng.uniform(0.02, 0.3, 3)))
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        prims.append(OrientedBox(rng.uniform([-3, -3, 0.5], [3, 3, 6.5]), rng.uniform(0.02, 0.2, 3), q.astype(np.float32)))
        v0 = rng.uniform([-3, -3, 0.5], [3, 3, 6.5])
        prims.append(Triangle(v0, v0 + rng.normal(size=3) * 0.3, v0 + rng.normal(size=3) * 0.3))
    # <= 32 primitives: the leg is culled per run from the bounding spheres and normal cones (leg_masks), for shared
    # directions and for point sources near enough to have parallax over a run; more: per iteration from the rays
    near = (-300.0 * src).astype(np.float32)
    for plist, srcs, stype in ((prims[:6], src, "parallel"), (prims[:30], src, "parallel"), (prims[:30], near, "point"),
                               (prims, src, "parallel")):
        tel = I.Telescope(base.mirror_groups, group_obstructions(plist), base.sensors)
        out = []
        for cull, bin_min in ((True, 256), (False, 0)):
            Rm.cull_obstructions, old_bin = cull, Rm.bin_samples_min
            Rm.bin_samples_min = bin_min
            try:
                tel._cache.pop("world", None)
                xy, v = render_debug(tel, srcs, val, stype, 0)
                img = render(tel, srcs, val, stype, 0)
                out.append((xy.cpu().numpy(), v.cpu().numpy(), img.cpu().numpy()))
            finally:
                Rm.cull_obstructions, Rm.bin_samples_min = True, old_bin
        assert np.array_equal(out[0][1], out[1][1])
        assert np.array_equal(out[0][0], out[1][0])
        np.testing.assert_allclose(out[0][2], out[1][2], rtol=1e-4, atol=1e-6 * out[1][2].max())
        lit = (out[0][1] != 0).mean()
        assert 0.05 < lit < 0.95, lit
