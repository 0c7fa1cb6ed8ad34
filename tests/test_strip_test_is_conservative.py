"""Level-3 culling for far / parallel sources (csrc/iact_cull.cuh strip_masks): a run of table rows with bounding
sphere (c, R) drops cylinder e when |n_e.c - n_e.p1_e| > R + r_e + margins, n_e = unit(u x axis_e).  This CPU test
restates that rule in float32 NumPy and checks on random geometry that a dropped cylinder is never hit -- according
to the oracle's literal float32 cylinder test (intersections.py:44-87) -- by any ray that starts inside the sphere
and runs along the beam axis (plus the residual tilt the rule allows for)."""
import numpy as np

from oracle import trace as otrace

f32 = np.float32


def _strip_record(u, p1, p2, r, c_facet, R_facet, invD):
    """(n, k, rr) as lane e computes them; rr = inf: always kept."""
    ax = p2 - p1
    w = np.cross(u, ax).astype(f32)
    w2, a2 = f32(w @ w), f32(ax @ ax)
    if not (w2 >= f32(2.5e-3) * a2 and a2 > f32(1e-20)):
        return np.zeros(3, f32), f32(0), f32(np.inf)
    n = (w / np.sqrt(w2)).astype(f32)
    k = f32(n @ p1)
    rq = f32(abs(r) * f32(1.0001))                                       # the proxy radius staged in shared memory
    tfar = f32(1.1) * (max(f32((p1 - c_facet) @ u), f32((p2 - c_facet) @ u), f32(0)) + rq + R_facet)
    rr = rq + f32(2e-3) + f32(1e-5) * tfar + R_facet * f32(1.5708) * tfar * invD
    return n, k, f32(rr)


def test_dropped_cylinders_are_never_hit():
    rng = np.random.default_rng(7)
    n_dropped = n_kept = n_hit_kept = 0
    for trial in range(400):
        u = rng.normal(size=3); u /= np.linalg.norm(u); u = u.astype(f32)            # beam axis: towards the source
        c_facet = (rng.normal(size=3) * 5).astype(f32)
        R_facet = f32(rng.uniform(0.3, 0.8))
        D = 10 ** rng.uniform(8, 11)                                                  # far source: R / D < 1e-7
        invD = f32(1.0 / D)
        # a strut somewhere in front of the facet, any orientation, 1-20 cm thick
        mid = c_facet + u * f32(rng.uniform(2, 40)) + (rng.normal(size=3) * rng.choice([0.2, 1.0, 3.0])).astype(f32)
        axd = rng.normal(size=3); axd /= np.linalg.norm(axd)
        half = rng.uniform(0.5, 15)
        p1 = (mid - axd * half).astype(f32); p2 = (mid + axd * half).astype(f32)
        r = f32(rng.uniform(0.01, 0.2))
        n, k, rr = _strip_record(u, p1, p2, r, c_facet, R_facet, invD)
        # runs: small spheres inside the facet's sphere
        for _ in range(6):
            R_run = f32(rng.uniform(0.03, 0.3))
            off = rng.normal(size=3); off *= rng.uniform(0, R_facet - R_run) / np.linalg.norm(off)
            c_run = (c_facet + off).astype(f32)
            keep = not (abs(f32(n @ c_run) - k) > R_run + rr)
            # rays of that run: origins in the sphere, direction along u up to the residual tilt (R / D + rounding)
            m = 400
            o = rng.normal(size=(m, 3)); o *= (R_run * rng.uniform(0, 1, size=(m, 1)) ** (1 / 3)) / np.linalg.norm(o, axis=1, keepdims=True)
            o = (c_run + o).astype(f32)
            d = (u + rng.normal(size=(m, 3)) * 2e-7).astype(f32)
            d /= np.linalg.norm(d, axis=1, keepdims=True)
            t = otrace.intersect_cylinder(o, d.astype(f32), p1[None], p2[None], np.array([r], f32), f32)[:, 0]
            hit = t < 1e10
            if keep:
                n_kept += 1; n_hit_kept += int(hit.any())
            else:
                n_dropped += 1
                assert not hit.any(), (trial, float(np.abs(n @ c_run - k)), float(R_run + rr))
    # the rule must actually cull, and what it keeps must often be needed
    assert n_dropped > 300 and n_kept > 300 and n_hit_kept > 0.3 * n_kept, (n_dropped, n_kept, n_hit_kept)
