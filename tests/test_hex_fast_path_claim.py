"""The hex fast path of the trace kernel (csrc/iact_render.cu, trace_ray): a hit whose hex norm about the centre of a
cached cell (q0, r0) is < 0.9999 is assigned to that cell without rounding.  This CPU test checks the claim behind
it with the oracle's float32 restatement of the reference (hexagonal.py:22-47): for such points `_axial_round` of
`_cartesian_to_axial` returns exactly (q0, r0), for the three HESS camera geometries and at the largest table
coordinates, and the norm the edge rejection sees is the one the fast path computed (same centre arithmetic)."""
import numpy as np
import pytest

from oracle import scene as osc, trace as otrace

f32 = np.float32
SQRT3 = 1.7320508075688772


@pytest.mark.parametrize("size,qmax", [(0.0242122, 20), (0.0288668, 27), (0.0242122, 29), (1.0, 200)])
def test_points_inside_a_cell_round_to_it(size, qmax):
    rng = np.random.default_rng(0)
    n = 400_000
    inradius = size * SQRT3 / 2
    q0 = rng.integers(-qmax, qmax + 1, n).astype(f32)
    r0 = rng.integers(-qmax, qmax + 1, n).astype(f32)
    # centre as kernel and reference compute it (hexagonal.py:27-29 with the constants folded in double, then float32)
    cx = f32(size * SQRT3) * (q0 + r0 / f32(2))
    cy = f32(size * 1.5) * r0
    # points all over the cell, concentrated towards its boundary
    ang = rng.uniform(0, 2 * np.pi, n)
    rad = inradius * 1.2 * rng.uniform(0, 1, n) ** 0.3
    xg = (cx + (rad * np.cos(ang)).astype(f32)).astype(f32)
    yg = (cy + (rad * np.sin(ang)).astype(f32)).astype(f32)
    hn = otrace.hex_norm(xg - cx, yg - cy, inradius, f32)
    safe = hn < f32(0.9999)
    assert 0.5 < safe.mean() < 0.99
    q, r = osc.cartesian_to_axial(xg, yg, size, f32)
    qi, ri = otrace.axial_round(q, r)
    assert np.array_equal(qi[safe], q0[safe]) and np.array_equal(ri[safe], r0[safe])
    # the margin is not vacuous: beyond the cell the rounding does move on
    out = hn > f32(1.0001)
    assert out.any() and not np.any((qi[out] == q0[out]) & (ri[out] == r0[out]))


@pytest.mark.parametrize("scene,sensor_idx", [("CT3", 0), ("CT5", 0), ("CT5", 1)])
def test_hits_outside_the_bounding_circle_belong_to_no_pixel(scene, sensor_idx):
    """The second early exit of trace_ray: a hit whose squared distance from the grid offset exceeds
    ``hex_outer_radius``^2 (max pixel-centre distance + 1.001 hex_size + 1e-5, sensors/hexagonal.py) is dropped
    without rounding.  Claim: the reference's lookup (hexagonal.py:155-191) assigns such a hit to no pixel."""
    from iactrace_b200.io import load_packed_config
    from iactrace_b200.sensors import HexagonalSensor
    cfg = load_packed_config(scene)["sensors"][sensor_idx]
    centers = np.stack([np.asarray(cfg["centers_x"], f32), np.asarray(cfg["centers_y"], f32)], axis=1)
    s = HexagonalSensor(cfg["position"], cfg["orientation"], centers, cfg.get("edge_width", 0.0))
    so = osc.make_hex_sensor(np.asarray(cfg["position"], f32), np.asarray(cfg["orientation"], f32), centers,
                             cfg.get("edge_width", 0.0))
    rng = np.random.default_rng(1)
    n = 300_000
    ang = rng.uniform(0, 2 * np.pi, n)
    rad = s.outer_radius * rng.uniform(0.9, 1.3, n)
    x = (f32(s.grid_offset[0]) + (rad * np.cos(ang)).astype(f32)).astype(f32)
    y = (f32(s.grid_offset[1]) + (rad * np.sin(ang)).astype(f32)).astype(f32)
    # the kernel's test, in its arithmetic: grid coordinates (hex_grid_coords), r^2 = fma(xg, xg, yg * yg) > r_out2
    tx, ty = x - f32(s.grid_offset[0]), y - f32(s.grid_offset[1])
    cr, sr = f32(np.cos(f32(-s.grid_rotation))), f32(np.sin(f32(-s.grid_rotation)))
    xg, yg = cr * tx - sr * ty, sr * tx + cr * ty
    out = (xg.astype(np.float64) ** 2 + yg.astype(np.float64) ** 2).astype(f32) > f32(s.outer_radius ** 2)
    assert 0.3 < out.mean() < 0.95
    _, valid, _ = otrace.hex_index(so, x, y, f32)
    assert not valid[out].any()
    assert valid[~out].any()                      # the circle is not so large that it never triggers near the rim
