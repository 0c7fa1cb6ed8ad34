"""The claim behind the cylinder test of the CUDA path (csrc/iact_trace.cuh, cyl_dir / cyl_ray): the discriminant
b^2 - 4ac of the reference (intersections.py:55-57) equals 4 a r^2 - (oc.w)^2 with w = 2 rdp x ax, and evaluated in
float32 the second form decides "the ray line passes within r of the axis" correctly down to micrometres, where the
literal form is noise within millimetres of the silhouette -- and biased towards "hit".  Checked here with NumPy on
telescope-sized geometry (ray origins tens of metres from a thin strut)."""
import numpy as np

f32 = np.float32


def _setup(n, rng, miss_by):
    """n rays whose lines pass the axis of a random thin cylinder at distance r + miss_by (signed, metres)."""
    ax = rng.normal(size=(n, 3)); ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    keep = np.abs(np.sum(ax * u, axis=1)) < 0.95                      # not the near-axial (literal-form) regime
    ax, u = ax[keep], u[keep]
    n = len(ax)
    r = rng.uniform(0.005, 0.03, n)
    nrm = np.cross(u, ax); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)      # common normal of the two lines
    p1 = rng.uniform(-15, 15, (n, 3))
    dist = r + miss_by
    # origin: on the line at signed distance `dist` from the axis, 10-40 m down the ray and anywhere along the axis
    o = p1 + nrm * dist[:, None] + ax * rng.uniform(-3, 3, (n, 1)) - u * rng.uniform(10, 40, (n, 1))
    return ax, u, r, p1, o


def _forms(ax, u, r, p1, o, dt):
    ax, u, r, p1, o = (x.astype(dt) for x in (ax, u, r, p1, o))
    dot = lambda a, b: (a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1] + a[:, 2] * b[:, 2]).astype(dt)
    oc = (o - p1).astype(dt)
    oc_ax, rd_ax = dot(oc, ax), dot(u, ax)
    ocp = (oc - oc_ax[:, None] * ax).astype(dt)
    rdp = (u - rd_ax[:, None] * ax).astype(dt)
    a = dot(rdp, rdp)
    b = (dt(2) * dot(ocp, rdp)).astype(dt)
    c = (dot(ocp, ocp) - r * r).astype(dt)
    literal = (b * b - dt(4) * a * c).astype(dt)
    w = np.cross((dt(2) * rdp).astype(dt), ax).astype(dt)
    g = dot(oc, w)
    triple = (dt(4) * a * (r * r) - g * g).astype(dt)
    return literal, triple


def test_the_two_forms_are_the_same_quantity():
    rng = np.random.default_rng(0)
    geo = _setup(20000, rng, rng.uniform(-0.02, 0.02, 20000)[:0].sum() + 0.004)
    lit, tri = _forms(*geo, np.float64)
    scale = np.abs(lit).max()
    assert np.allclose(lit, tri, rtol=0, atol=1e-9 * max(scale, 1.0))


def test_float32_decisions_near_the_silhouette():
    rng = np.random.default_rng(1)
    n = 200000
    wrong_lit, wrong_tri, biased_hits, total = {}, {}, 0, 0
    for miss in (-1e-3, -1e-4, -2e-5, 2e-5, 1e-4, 1e-3):          # inside (negative) / outside the silhouette by this much
        geo = _setup(n, rng, miss)
        truth = np.full(len(geo[0]), miss < 0)
        lit32, tri32 = _forms(*geo, f32)
        lit64, tri64 = _forms(*geo, np.float64)
        assert np.array_equal(lit64 >= 0, truth) and np.array_equal(tri64 >= 0, truth)      # float64: both exact here
        wrong_lit[miss] = float(np.mean((lit32 >= 0) != truth))
        wrong_tri[miss] = float(np.mean((tri32 >= 0) != truth))
    # the form the kernel uses: no wrong decision 20 micrometres from the silhouette or farther
    assert all(v == 0.0 for v in wrong_tri.values()), wrong_tri
    # the literal form in float32: a coin toss at 20 um and at 0.1 mm, still wrong for a noticeable share at 1 mm
    assert wrong_lit[2e-5] > 0.2 and wrong_lit[-2e-5] > 0.1 and wrong_lit[1e-4] > 0.1, wrong_lit
    assert wrong_lit[1e-3] > 1e-3, wrong_lit
    # ... and biased: more misses called hits than hits called misses (disc >= 0 includes the cancelled zero)
    assert wrong_lit[2e-5] > wrong_lit[-2e-5] and wrong_lit[1e-4] > wrong_lit[-1e-4], wrong_lit
