"""The CUDA cylinder test (csrc/iact_trace.cuh hit_cylinder) decides "some valid candidate hit has t < 1e10" from the
interval in which the ray is inside the solid cylinder instead of validating the reference's four candidates one by
one (intersections.py:44-87).  This CPU test restates both forms in NumPy on adversarial ray/cylinder pairs: identical
decisions in float64, and in float32 only cap-rim grazing rays may differ."""
import numpy as np

from oracle import trace as otrace


def _pairs(n, seed):
    rng = np.random.default_rng(seed)
    p1 = rng.normal(size=(n, 3)) * 3
    ax = rng.normal(size=(n, 3)); ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    h = rng.uniform(0.05, 6, size=(n, 1))
    p2 = p1 + ax * h
    r = rng.uniform(0.01, 0.5, size=n)
    # aim at points on / near the surface, the caps and the rims, from 5 cm to 40 m away, some origins inside
    target = p1 + ax * rng.uniform(-0.3, 1.3, size=(n, 1)) * h + rng.normal(size=(n, 3)) * r[:, None] * 1.2
    o = target + rng.normal(size=(n, 3)) * rng.choice([0.05, 1, 10, 40], size=(n, 1))
    d = target - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    k = n // 10                                              # rays parallel / nearly parallel to the axis
    d[:k] = ax[:k] + rng.normal(size=(k, 3)) * rng.choice([0.0, 1e-5, 1e-3, 3e-2], size=(k, 1))
    d[:k] /= np.linalg.norm(d[:k], axis=1, keepdims=True)
    return o, d, p1, p2, r


def _both_forms(o, d, p1, p2, r, dt):
    o, d, p1, p2, r = (x.astype(dt) for x in (o, d, p1, p2, r))
    eps = dt(1e-8)
    axis = p2 - p1
    hh = np.sqrt((axis * axis).sum(1))
    axn = axis / hh[:, None]
    oc = o - p1
    oc_ax = (oc * axn).sum(1); rd_ax = (d * axn).sum(1)
    ocp = oc - oc_ax[:, None] * axn; rdp = d - rd_ax[:, None] * axn
    a = (rdp * rdp).sum(1); b = dt(2) * (ocp * rdp).sum(1); cc = (ocp * ocp).sum(1) - r * r
    disc = b * b - dt(4) * a * cc
    with np.errstate(all="ignore"):
        sq = np.sqrt(np.maximum(disc, 0)); inv = dt(1) / (dt(2) * a + eps)
        t1 = (-b - sq) * inv; t2 = (-b + sq) * inv
        ia = dt(1) / (rd_ax + eps); tb = -oc_ax * ia; tt = (hh - oc_ax) * ia
        # the reference's candidates, one by one
        y1 = oc_ax + t1 * rd_ax; y2 = oc_ax + t2 * rd_ax
        pb = ocp + tb[:, None] * rdp; pt = ocp + tt[:, None] * rdp
        literal = (((t1 > eps) & (y1 >= 0) & (y1 <= hh) & (disc >= 0) & (t1 < 1e10))
                   | ((t2 > eps) & (y2 >= 0) & (y2 <= hh) & (disc >= 0) & (t2 < 1e10))
                   | ((tb > eps) & ((pb * pb).sum(1) <= r * r) & (tb < 1e10))
                   | ((tt > eps) & ((pt * pt).sum(1) <= r * r) & (tt < 1e10)))
        # the kernel's interval form, with its fallback to the literal test within 1.8 deg of the axis
        lo = np.maximum(t1, np.minimum(tb, tt)); hi = np.minimum(t2, np.maximum(tb, tt))
        tc = np.where(lo > eps, lo, hi)
        interval = (disc >= 0) & (lo <= hi) & (tc > eps) & (tc < 1e10)
    return literal, np.where(a >= dt(1e-3), interval, literal), a


def test_literal_restatement_is_the_oracle():
    o, d, p1, p2, r = _pairs(2000, 0)
    literal, _, _ = _both_forms(o, d, p1, p2, r, np.float64)
    t = np.array([otrace.intersect_cylinder(o[i].astype(np.float64), d[i].astype(np.float64), p1[i:i + 1], p2[i:i + 1],
                                            r[i:i + 1], np.float64)[0] for i in range(len(o))])
    assert np.array_equal(literal, t < 1e10)


def test_interval_form_equals_candidate_tests():
    o, d, p1, p2, r = _pairs(1_000_000, 1)
    lit64, int64, a = _both_forms(o, d, p1, p2, r, np.float64)
    assert 0.2 < lit64.mean() < 0.6 and 0.8 < (a >= 1e-3).mean() < 0.99          # both branches well exercised
    assert np.array_equal(lit64, int64)
    lit32, int32, _ = _both_forms(o, d, p1, p2, r, np.float32)
    assert (lit32 != int32).mean() < 2e-4                                         # cap-rim grazing only (measured 4.5e-5)
