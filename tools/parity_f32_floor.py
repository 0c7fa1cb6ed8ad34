#!/usr/bin/env python
"""What float32 itself costs against the float64 oracle: the FLOAT32 ORACLE (the reference's arithmetic, op by op, in
NumPy float32) plays the kernel in tests/_parity.py.  CPU only.  Its shadow-flip and moved-pixel rates are the floor
any float32 implementation of the reference sits on; profiles/parity_rNN.json holds the CUDA kernels' rates on the
same scenes (tools/parity_report.py).

    python tools/parity_f32_floor.py > profiles/parity_r02_f32_floor.json
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
from oracle import prng, scene as oscene, trace as otrace  # noqa: E402
from iactrace_b200.io import load_packed_config  # noqa: E402
from iactrace_b200.workloads import point_grid  # noqa: E402
from _parity import ray_parity  # noqa: E402

out = {"note": "float32 NumPy oracle vs float64 NumPy oracle, same sample tables (oracle sampler, seed 0)", "cases": {}}
for name, scene, M, src, sensors in (
        ("config1_ct3_on_axis_M1000", "CT3", 1000, np.array([[0, 0, 1e10]], np.float32), (0, 1)),
        ("config2_ct5_M115_6_sources", "CT5", 115, point_grid(64, 1.5)[[0, 777, 2080, 2500, 3333, 4095]], (0, 2))):
    sc = oscene.build_scene(load_packed_config(scene), M, prng.key(0))
    val = np.linspace(0.5, 1.5, len(src)).astype(np.float32)
    for si in sensors:
        xy, v = otrace.render_debug(sc, src, val, "point", si, np.float32)
        oxy, ov = otrace.render_debug(sc, src, val, "point", si, np.float64)
        s = sc["sensors"][si]
        idx, valid, _ = otrace.pixel_index(s, xy[:, 0], xy[:, 1], np.float32)
        pix = np.where(valid, idx, -1)
        r = ray_parity(xy, v, pix, oxy, ov, s, xy_tol=1e-4, flip_budget=1e-2, edge_budget=0.5)
        out["cases"][f"{name}_{s['type']}"] = r["stats"]
        print(name, si, r["stats"], file=sys.stderr, flush=True)
print(json.dumps(out, indent=1))
