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
