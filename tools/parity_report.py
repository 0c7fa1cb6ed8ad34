#!/usr/bin/env python
"""Measured parity figures of the CUDA path against the float64 oracle and the executed-reference fixtures, for
BASELINE configs 1-4 at sizes the oracle finishes in seconds.  Run on a GPU box:

    python tools/parity_report.py > gpurun_out/parity.json

Output (copied to profiles/parity_rNN.json): per case the shadow-flip rate, the rate of rays that land in another
pixel (all of them verified to sit on a pixel edge), the largest per-ray value / coordinate error, and the largest
per-pixel image error on the pixels no such ray touches, with how many lit pixels and how much flux that covers.
"""
from __future__ import annotations

import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]


def main():
    import torch
    import iactrace_b200 as I
    from iactrace_b200.core import render, render_response_matrix
    from iactrace_b200.io import build_telescope, load_packed_config
    from iactrace_b200.workloads import cassegrain_config, point_grid, parallel_grid, star_field
    from _bridge import subset_config
    from _parity import compare_rays, compare_image, compare_soft_image, subset_rays

    out = {"device": torch.cuda.get_device_name(0), "cases": {}}

    def case(name, tel, src, val, stype, si, xy_tol, matrix=False, **kw):
        t0 = time.time()
        r = compare_rays(tel, src, val, stype, si, xy_tol=xy_tol, flip_budget=1e-3)
        if matrix:
            M = render_response_matrix(tel, src, val, stype, si).cpu().numpy()
            n_m = tel.mirror_groups[0].points.shape[1]
            src_of_ray = (np.arange(r["v"].size) // n_m) % len(src)
            rows = [compare_image(M[i], subset_rays(r, src_of_ray == i), min_lit=0, min_flux_share=0.0) for i in range(len(src))]
            st = dict(rows=len(rows), lit_pixels=sum(x["lit_pixels"] for x in rows),
                      lit_pixels_compared=sum(x["lit_pixels_compared"] for x in rows),
                      max_rel_err_clean_pixels=max(x["max_rel_err_clean_pixels"] for x in rows),
                      max_rel_err_vs_own_rays=max(x["max_rel_err_vs_own_rays"] for x in rows))
        else:
            img = render(tel, src, val, stype, si).cpu().numpy()
            st = compare_image(img, r, min_lit=0, min_flux_share=0.0, **kw)
        out["cases"][name] = dict(rays=r["stats"], image=st, seconds=round(time.time() - t0, 1))
        print(name, out["cases"][name], file=sys.stderr, flush=True)

    # config 1: CT3, on-axis point source at 1e10, MCIntegrator(1000), both sensors (full size)
    ct3 = build_telescope(load_packed_config("CT3"), I.MCIntegrator(1000), I.random.key(0))
    on_axis = np.array([[0.0, 0.0, 1e10]], np.float32)
    case("config1_ct3_on_axis_M1000_hex", ct3, on_axis, np.ones(1, np.float32), "point", 0, 2e-5)
    case("config1_ct3_on_axis_M1000_lid", ct3, on_axis, np.ones(1, np.float32), "point", 1, 2e-5)
    # config 2 geometry: full CT5 (876 facets, 271 primitives), M = 115, 6 of the 4096 grid sources
    ct5 = build_telescope(load_packed_config("CT5"), I.MCIntegrator(115), I.random.key(0))
    src = point_grid(64, 1.5)[[0, 777, 2080, 2500, 3333, 4095]]
    val = np.linspace(0.5, 1.5, len(src)).astype(np.float32)
    case("config2_ct5_M115_6_sources_hex", ct5, src, val, "point", 0, 6e-5)
    case("config2_ct5_M115_6_sources_lid", ct5, src, val, "point", 2, 6e-5)
    # config 3 geometry: Cassegrain with obstructions, 24 stars x 6 segments x 512 samples
    cas = build_telescope(cassegrain_config(True), I.MCIntegrator(512), I.random.key(0))
    d, flux = star_field(24, 3.0)
    case("config3_cassegrain_M512_24_stars", cas, d, flux, "parallel", 0, 5e-6)
    # config 4 geometry: CT3 + roughness 24", response matrix over a 6 x 6 grid of the 64 x 64 directions, M = 64
    ct3r = build_telescope(load_packed_config("CT3"), I.MCIntegrator(64), I.random.key(42)).apply_roughness(24)
    dirs = parallel_grid(64, 5.5).reshape(64, 64, 3)[4::11, 4::11].reshape(-1, 3)
    case("config4_ct3_matrix_M64_36_directions", ct3r, dirs, np.ones(len(dirs), np.float32), "parallel", 0, 2e-5, matrix=True)
    # soft sensors (config 5's forward)
    from iactrace_b200.sensors import DifferentiableHexagonalSensor
    hard = ct5.sensors[0]
    t5 = build_telescope(subset_config(load_packed_config("CT5"), mirror_step=9), I.MCIntegrator(32), I.random.key(0))
    t5 = t5.replace_sensor(DifferentiableHexagonalSensor(hard.position, hard.rotation, hard.hex_centers, 0.5, 1,
                                                         grid=hard.grid_constants()), 0)
    out["cases"]["config5_forward_soft_hex"] = compare_soft_image(t5, src[:2], val[:2], "point", 0)

    # executed-reference fixtures (tests/golden): flips per case, to size the budgets in test_gpu_golden.py
    from golden.cases import ALL_CASES, case_values
    from iactrace_b200 import random as R
    from iactrace_b200.core import render_debug
    gold = dict(np.load(ROOT / "tests" / "golden" / "reference_golden.npz"))
    large = ROOT / "tests" / "golden" / "reference_golden_large.npz"
    if large.exists():
        gold.update(np.load(large))
    g = {}
    for name, c in ALL_CASES.items():
        if f"{name}/s{c['sensors'][0]}/debug_vals" not in gold:
            continue
        R.set_rng_mode(c["mode"])
        try:
            tel = build_telescope(c["cfg"](), I.MCIntegrator(c["M"]), R.key(c["seed"]))
        finally:
            R.set_rng_mode(R.PARTITIONABLE)
        if c["rough"]:
            tel = tel.apply_roughness(c["rough"])
        for si in c["sensors"]:
            xy, v = render_debug(tel, c["src"], case_values(name), c["stype"], si)
            xy, v = xy.cpu().numpy(), v.cpu().numpy()
            gp, gv = gold[f"{name}/s{si}/debug_pts"], gold[f"{name}/s{si}/debug_vals"]
            flips = (v != 0) != (gv != 0)
            ok = ~flips & (gv != 0) & (np.abs(gp[:, 0]) < 1e9)
            g[f"{name}/s{si}"] = dict(n_rays=int(v.size), shadow_flips=int(flips.sum()),
                                      max_xy_err_m=float(np.abs(xy[ok] - gp[ok]).max()),
                                      max_value_rel_err=float((np.abs(v[ok] - gv[ok]) / np.abs(gv[ok])).max()))
    out["executed_reference_fixtures"] = g
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
