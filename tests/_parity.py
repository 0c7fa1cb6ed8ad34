"""Parity machinery shared by the GPU tests, ``__graft_entry__.smoke()`` and ``tools/parity_report.py``.

The bars (BASELINE.json ``north_star``): images within 1e-4 relative per pixel, pixel indices bit-exact for rays
away from pixel edges.  A float32 kernel and a float64 oracle cannot agree on a ray that sits within rounding
noise of a pixel edge or of an obstruction's silhouette ("ambiguous" rays): such a ray lands in the neighbouring
pixel or flips its shadow decision.  The checks therefore work per ray first and per pixel second:

* every non-ambiguous ray: same shadow decision, same pixel, value within 1e-5 relative, hit within ``xy_tol``;
* rays the two sides treat differently (shadow flip, or another pixel) are COUNTED and bounded (``flip_budget``,
  ``edge_budget``), and every ray that changes pixel must lie within twice its own coordinate error of a pixel edge;
* image, pixels into which neither side puts such a ray ("clean"): 1e-4 relative against the float64 binning of the
  oracle's rays -- and the test asserts that lit pixels really were compared (count and flux share);
* image, every pixel, unconditionally: |img - oracle| <= 1e-4 oracle + (sum of |value| of the ambiguous rays that
  either side puts into that pixel);
* image against the float64 binning of the kernel's own per-ray output, every pixel.
"""
from __future__ import annotations

import numpy as np

EDGE_MARGIN = 1e-5   # metres: rays closer than this to a pixel edge may land on either side


def compare_rays(tel, src, val, stype, sensor_idx, xy_tol=2e-5, flip_budget=2e-5, edge_budget=5e-2):
    """Per-ray parity of ``render_debug`` against the float64 oracle.  Returns the arrays the image checks need
    plus a ``stats`` dict with the measured rates."""
    from iactrace_b200.core import render_debug
    from oracle import trace as otrace
    from _bridge import to_oracle_scene
    osc = to_oracle_scene(tel)
    xy, v, pix = render_debug(tel, src, val, stype, sensor_idx, return_pixels=True)
    xy, v, pix = xy.cpu().numpy(), v.cpu().numpy(), pix.cpu().numpy()
    oxy, ov = otrace.render_debug(osc, src, val, stype, sensor_idx, np.float64)
    r = ray_parity(xy, v, pix, oxy, ov, osc["sensors"][sensor_idx], xy_tol=xy_tol, flip_budget=flip_budget,
                   edge_budget=edge_budget)
    r["osc"] = osc
    return r


def subset_rays(r, mask):
    """The rays of ``r`` selected by ``mask`` (e.g. one source's rays for a response-matrix row)."""
    return {k: (a[mask] if isinstance(a, np.ndarray) and a.shape[:1] == mask.shape else a) for k, a in r.items()}


def ray_parity(xy, v, pix, oxy, ov, s, xy_tol=2e-5, flip_budget=2e-5, edge_budget=5e-2, value_rtol=1e-5, index_dt=np.float64):
    """Per-ray comparison of the kernel's (xy, v, pix) with reference rays (oxy, ov) on oracle sensor ``s``; the
    reference pixel index is the oracle's binning of (oxy) in ``index_dt`` arithmetic."""
    from oracle import trace as otrace
    assert xy.shape == oxy.shape and v.shape == ov.shape
    lit, olit = v != 0, ov != 0
    flips = lit != olit
    assert flips.mean() <= flip_budget, f"{flips.sum()} shadow flips of {flips.size} (budget {flip_budget:g})"
    both = lit & olit
    np.testing.assert_allclose(v[both], ov[both], rtol=value_rtol)
    ok = both & (np.abs(oxy[:, 0]) < 1e9)
    xy_err = float(np.abs(xy[ok] - oxy[ok]).max()) if ok.any() else 0.0
    assert xy_err < xy_tol, xy_err
    oidx, ovalid, edge = otrace.pixel_index(s, oxy[:, 0].astype(index_dt), oxy[:, 1].astype(index_dt), index_dt)
    edge = np.asarray(edge, np.float64)
    opix = np.where(ovalid, oidx, -1)
    # distance (metres) of the oracle's hit to the nearest binning decision boundary
    if s["type"] == "hexagonal":
        inr = s["hex_inradius"]
        thr = 1.0 - s["edge_width"] / inr
        d_edge = np.minimum(np.abs(edge - thr), np.abs(edge - 1.0)) * inr
    else:
        d_edge = np.minimum(np.abs(edge - s["edge_width"]), np.abs(edge))
        # the outer border of the pixel grid is an edge too (x0, x0 + W dx, ...)
        xr = (oxy[:, 0] - s["x0"]) / s["dx"]
        yr = (oxy[:, 1] - s["y0"]) / s["dy"]
        for c, n, d in ((xr, s["width"], s["dx"]), (yr, s["height"], s["dy"])):
            d_edge = np.minimum(d_edge, np.minimum(np.abs(c), np.abs(c - n)) * abs(d))
    near = d_edge < EDGE_MARGIN
    chk = ok & ~near
    mism = int((pix[chk] != opix[chk]).sum())
    assert mism == 0, f"{mism} pixel mismatches away from edges"
    n_edge = int((ok & near).sum())
    # every ray the two sides bin differently must be explained by its own coordinate rounding: the oracle's hit lies
    # within (twice) the kernel-vs-oracle coordinate distance of a decision boundary
    moved = ok & (pix != opix)
    err_r = np.abs(xy - oxy).max(axis=1)
    unexplained = moved & ~(d_edge <= 2.0 * err_r + 1e-6)
    assert not unexplained.any(), f"{int(unexplained.sum())} rays change pixel without sitting on an edge"
    assert moved.sum() <= edge_budget * max(int(ok.sum()), 1), f"{int(moved.sum())} rays change pixel of {int(ok.sum())}"
    with np.errstate(all="ignore"):
        rel = np.abs(v[both] - ov[both]) / np.abs(ov[both])
    stats = dict(n_rays=int(v.size), n_lit=int(both.sum()), shadow_flips=int(flips.sum()), flip_rate=float(flips.mean()),
                 rays_within_10um_of_an_edge=n_edge, rays_binned_differently=int(moved.sum()),
                 binned_differently_rate=float(moved.sum() / max(int(ok.sum()), 1)),
                 max_value_rel_err=float(rel.max()) if rel.size else 0.0, max_xy_err_m=xy_err,
                 pixel_mismatch_away_from_edges=mism, shadowed_fraction=float((~olit).mean()))
    return dict(xy=xy, v=v, pix=pix, oxy=oxy, ov=ov, opix=opix, ambiguous=moved | flips, near_edge=ok & near, stats=stats)


def compare_image(img, r, rtol=1e-4, min_lit=1, min_flux_share=0.9, dense_rtol=None):
    """Image checks described in the module docstring; ``r`` comes from ``compare_rays``.  Returns stats.

    ``dense_rtol``: additionally assert |img - oracle| <= dense_rtol * oracle on EVERY pixel that collects at least
    3e4 rays, ambiguous rays included: there one ray weighs a third of the bar, so the handful of rays that flip or
    change pixel cannot hide a real discrepancy (an on-axis spot on a hex camera: ~9e4 rays in the central pixel)."""
    got = np.asarray(img, np.float64).reshape(-1)
    npx = got.size
    amb = r["ambiguous"]
    w_o = np.where(r["opix"] >= 0, np.abs(r["ov"]), 0.0)
    oimg = np.bincount(r["opix"][r["opix"] >= 0], weights=r["ov"][r["opix"] >= 0], minlength=npx)
    # slack: what the ambiguous rays can move, per pixel, on either side
    slack = np.zeros(npx)
    for arr, w in ((r["pix"], np.abs(r["v"]).astype(np.float64)), (r["opix"], w_o)):
        sel = amb & (arr >= 0)
        slack += np.bincount(arr[sel], weights=w[sel], minlength=npx)
    clean = slack == 0
    lit_clean = clean & (oimg > 0)
    n_lit = int((oimg > 0).sum())
    assert int(lit_clean.sum()) >= min(min_lit, n_lit), f"only {int(lit_clean.sum())} lit pixels were compared"
    share = float(oimg[lit_clean].sum() / max(oimg.sum(), 1e-300))
    assert share >= min_flux_share or n_lit == 0, f"clean pixels carry only {share:.3f} of the flux"
    atol = 1e-7 * max(oimg.max(), 1e-30)
    np.testing.assert_allclose(got[clean], oimg[clean], rtol=rtol, atol=atol)
    # unconditional per-pixel bound
    excess = np.abs(got - oimg) - (rtol * np.abs(oimg) + slack * (1 + 1e-5) + atol)
    assert excess.max() <= 0, f"pixel {int(excess.argmax())}: |img - oracle| exceeds the ambiguity bound by {excess.max():.3e}"
    # the image equals the float64 binning of the kernel's own per-ray output everywhere
    own = np.bincount(r["pix"][r["pix"] >= 0], weights=r["v"][r["pix"] >= 0].astype(np.float64), minlength=npx)
    np.testing.assert_allclose(got, own, rtol=5e-5, atol=1e-7 * max(own.max(), 1e-30))
    bright = np.bincount(r["opix"][r["opix"] >= 0], minlength=npx) >= 30000
    with np.errstate(all="ignore"):
        rel = np.abs(got - oimg)[lit_clean] / oimg[lit_clean]
        rel_own = (np.abs(got - own) / own)[own > 0]
        rel_bright = (np.abs(got - oimg) / oimg)[bright & (oimg > 0)]
    if dense_rtol is not None:
        assert rel_bright.size and rel_bright.max() <= dense_rtol, f"dense pixels differ by {rel_bright.max() if rel_bright.size else -1:.3e}"
    return dict(n_pixels=npx, lit_pixels=n_lit, lit_pixels_compared=int(lit_clean.sum()), flux_share_compared=share,
                max_rel_err_clean_pixels=float(rel.max()) if rel.size else 0.0,
                max_rel_err_vs_own_rays=float(rel_own.max()) if rel_own.size else 0.0,
                max_rel_err_dense_pixels_incl_ambiguous_rays=float(rel_bright.max()) if rel_bright.size else 0.0,
                dense_pixels=int(bright.sum()), tainted_pixels=int((~clean).sum()))


def compare_soft_image(tel, src, val, stype, sensor_idx, rtol_own=1e-4, rtol_oracle=5e-3):
    """Soft (Gaussian-splat) sensors.  Two statements:

    1. splat arithmetic: the image equals the float64 splat (oracle ``accumulate``) of the kernel's OWN per-ray hits
       to ``rtol_own`` -- this isolates the sensor code from the optics;
    2. end to end against the float64 oracle to ``rtol_oracle``.  This bound is looser than 1e-4 for a reason that has
       nothing to do with the kernel: the Gaussian taps amplify the float32 rounding of the hit coordinates
       (2e-5 .. 6e-5 m on 15-36 m lever arms; d ln w = hd dhd / sigma^2 reaches 1e-2 for an outer tap of one ray).
       The float32 ORACLE sits at the same distance from the float64 one; the function returns both distances.
    """
    from iactrace_b200.core import render, render_debug
    from oracle import trace as otrace
    from _bridge import to_oracle_scene
    osc = to_oracle_scene(tel)
    s = osc["sensors"][sensor_idx]
    img = render(tel, src, val, stype, sensor_idx).cpu().numpy().astype(np.float64)
    xy, v = render_debug(tel, src, val, stype, sensor_idx)
    xy, v = xy.cpu().numpy(), v.cpu().numpy()
    own = otrace.accumulate(s, xy[:, 0].astype(np.float64), xy[:, 1].astype(np.float64), v.astype(np.float64), np.float64)
    assert img.shape == own.shape
    np.testing.assert_allclose(img, own, rtol=rtol_own, atol=1e-6 * own.max())
    o64 = otrace.render(osc, src, val, stype, sensor_idx, np.float64)
    o32 = otrace.render(osc, src, val, stype, sensor_idx, np.float32).astype(np.float64)
    np.testing.assert_allclose(img, o64, rtol=rtol_oracle, atol=2e-5 * o64.max())

    def dist(a, b):
        m = b > 1e-3 * b.max()
        return float((np.abs(a - b)[m] / b[m]).max())
    return dict(max_rel_err_vs_own_rays=dist(img, own), max_rel_err_vs_f64_oracle=dist(img, o64),
                f32_oracle_vs_f64_oracle=dist(o32, o64), lit_pixels=int((o64 > 1e-3 * o64.max()).sum()))
