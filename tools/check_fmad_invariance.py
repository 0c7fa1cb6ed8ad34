#!/usr/bin/env python
"""The stage-0 per-ray chain is written with explicitly rounded intrinsics, so nvcc's mul+add contraction must not
matter: render_debug / render / response matrix from the product build and from a --fmad=false build of the same
sources (variants/libiactrace_b200_nofma.so) have to be BIT-IDENTICAL.  Run on a GPU box:

    python tools/check_fmad_invariance.py            # spawns itself once per library
"""
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]


def produce(path):
    import iactrace_b200 as I
    from iactrace_b200.core import render, render_debug, render_response_matrix
    from iactrace_b200.io import build_telescope, load_packed_config
    from iactrace_b200.workloads import point_grid, parallel_grid
    from _bridge import subset_config
    out = {}
    for name, M, step in (("CT5", 115, 5), ("CT3", 300, 3)):
        tel = build_telescope(subset_config(load_packed_config(name), mirror_step=step), I.MCIntegrator(M), I.random.key(0))
        for stype, src in (("point", point_grid(4, 1.5)), ("parallel", parallel_grid(4, 3.0)),
                           ("point", np.array([[3.0, -2.0, 60.0], [0.0, 0.0, 36.0], [40.0, 5.0, 20.0]], np.float32))):
            val = np.linspace(0.5, 1.5, len(src)).astype(np.float32)
            for si in (0, len(tel.sensors) - 1):
                k = f"{name}/{stype}{len(src)}/s{si}/"
                xy, v, pix = render_debug(tel, src, val, stype, si, return_pixels=True)
                out[k + "xy"], out[k + "v"], out[k + "pix"] = xy.cpu().numpy(), v.cpu().numpy(), pix.cpu().numpy()
                if si == 0:
                    out[k + "matrix"] = render_response_matrix(tel, src, val, stype, si).cpu().numpy()
    np.savez(path, **out)


def main():
    if len(sys.argv) > 1:
        return produce(sys.argv[1])
    tmp = Path(tempfile.mkdtemp())
    libs = {"product": "", "nofma": str(ROOT / "variants" / "libiactrace_b200_nofma.so")}
    for tag, lib in libs.items():
        env = dict(os.environ, IACTRACE_B200_LIB=lib)
        subprocess.run([sys.executable, __file__, str(tmp / f"{tag}.npz")], check=True, env=env)
    a, b = np.load(tmp / "product.npz"), np.load(tmp / "nofma.npz")
    bad = 0
    for k in a.files:
        if k.endswith("/matrix"):
            # the shared-memory histogram is filled by atomics in launch-dependent order: float32 summation noise
            same = np.allclose(a[k], b[k], rtol=2e-6, atol=1e-9)
        else:
            same = np.array_equal(a[k], b[k], equal_nan=True)
        if not same:
            bad += 1
            d = a[k] != b[k]
            print(f"DIFF {k}: {int(d.sum())} of {d.size} elements")
            for i in np.flatnonzero(d.reshape(-1))[:6]:
                print(f"    [{i}] product {a[k].reshape(-1)[i]!r}  nofma {b[k].reshape(-1)[i]!r}")
    n_rays = sum(a[k].size for k in a.files if k.endswith("/v"))
    print(f"fmad invariance: {len(a.files)} arrays, {n_rays} rays, {bad} arrays differ")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
