"""Generate tests/golden/reference_golden.npz by EXECUTING THE REFERENCE's own sources.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # the six small cases + unit vectors -> reference_golden.npz
    python tests/golden/make_golden.py --large    # full CT3 / CT5 and the uneven-polygon sampler cases
                                                  #   -> reference_golden_large.npz (about ten minutes)

JAX/Equinox are not installable here, so the reference's unmodified Python runs on
oracle/jaxshim (NumPy stand-in; see its README for what that does and does not pin).
"""
import json
import sys
import tempfile
from pathlib import Path

import numpy as np
import yaml

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT / "oracle" / "jaxshim"), "/root/reference", str(ROOT), str(ROOT / "tests")]

import jax  # noqa: E402  (the shim)
import jax.numpy as jnp  # noqa: E402
from jax import random as jrandom  # noqa: E402
import iactrace  # noqa: E402  (the reference)
from iactrace import MCIntegrator, Telescope  # noqa: E402
from iactrace.core import (render_response_matrix, euler_to_matrix, reflect, intersect_plane, intersect_cylinder,  # noqa: E402
                           intersect_box, intersect_sphere, intersect_oriented_box, intersect_triangle,
                           intersect_conic, AsphericSurface)
from iactrace.sensors.hexagonal import DifferentiableHexagonalSensor, _detect_hex_grid  # noqa: E402
from iactrace.sensors.square import DifferentiableSquareSensor  # noqa: E402
from oracle import prng  # noqa: E402
from golden.cases import CASES, LARGE_CASES, case_values  # noqa: E402

LARGE = "--large" in sys.argv
if LARGE:
    CASES = LARGE_CASES

A = lambda x: np.asarray(x)
out = {}
meta = {"reference_version": iactrace.__version__, "cases": {}}

for name, c in CASES.items():
    jrandom.MODE = prng.PARTITIONABLE if c["mode"] == "partitionable" else prng.LEGACY
    cfg = c["cfg"]()
    with tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False) as f:
        yaml.safe_dump(cfg, f)
    tel = Telescope.from_yaml(f.name, MCIntegrator(c["M"]), key=jax.random.key(c["seed"]))
    if c["rough"]:
        tel = tel.apply_roughness(c["rough"])
    for gi, g in enumerate(tel.mirror_groups):
        for fld in ("points", "normals", "perturbation_delta", "weights", "perturbation_scale", "positions", "rotations"):
            out[f"{name}/group{gi}/{fld}"] = A(getattr(g, fld))
    src, val = jnp.asarray(c["src"]), jnp.asarray(case_values(name))
    for si in c["sensors"]:
        pts, vals = tel(src, val, c["stype"], sensor_idx=si, debug=True)
        out[f"{name}/s{si}/debug_pts"], out[f"{name}/s{si}/debug_vals"] = A(pts), A(vals)
        out[f"{name}/s{si}/image"] = A(tel(src, val, c["stype"], sensor_idx=si))
        if A(out[f"{name}/s{si}/image"]).size < 5000:
            out[f"{name}/s{si}/matrix"] = A(render_response_matrix(tel, src, val, c["stype"], sensor_idx=si))
        s = tel.sensors[si]
        if hasattr(s, "hex_size"):
            out[f"{name}/s{si}/hexgrid"] = np.array([s.hex_size, s.hex_inradius, s.grid_rotation, s.grid_offset[0],
                                                     s.grid_offset[1], s.q_min, s.r_min], np.float64)
            out[f"{name}/s{si}/lookup"] = A(s.lookup_table)
    meta["cases"][name] = dict(M=c["M"], seed=c["seed"], mode=c["mode"], n_groups=len(tel.mirror_groups),
                               info={k: (v if not hasattr(v, "tolist") else A(v).tolist()) for k, v in tel.get_info().items()})
    if name == "ct3_point":
        # parameter-edit operations (operations.py:118-229) and the soft sensors
        t2 = tel.apply_misalignment_to_group(0, 15, 10, jax.random.key(4242)).apply_displacement_to_group(0, 0.02, jax.random.key(4242))
        out["ops/misaligned_rotations"] = A(t2.mirror_groups[0].rotations)
        out["ops/displaced_positions"] = A(t2.mirror_groups[0].positions)
        t3 = tel.resample_mirrors(MCIntegrator(5), jax.random.key(9))
        out["ops/resampled_points"] = A(t3.mirror_groups[0].points)
        hard = tel.sensors[0]
        soft = DifferentiableHexagonalSensor(hard.position, hard.rotation, hard.hex_centers, sigma=0.5, kernel_size=1)
        lid = tel.sensors[1]
        soft_sq = DifferentiableSquareSensor(lid.position, lid.rotation, 48, 32, (-0.768, 0.768, -0.512, 0.512), sigma=0.7, kernel_size=2)
        t4 = tel.replace_sensor(soft, 0).replace_sensor(soft_sq, 1)
        out["soft/hex_image"] = A(t4(src, val, "point", sensor_idx=0))
        out["soft/square_image"] = A(t4(src, val, "point", sensor_idx=1))

if LARGE:
    np.savez_compressed(ROOT / "tests" / "golden" / "reference_golden_large.npz", **out)
    (ROOT / "tests" / "golden" / "reference_golden_large.json").write_text(json.dumps(meta, indent=1, default=str))
    print(f"wrote {len(out)} arrays,", (ROOT / "tests" / "golden" / "reference_golden_large.npz").stat().st_size, "bytes")
    sys.exit(0)

# unit-level vectors for every primitive (intersections.py, reflection.py, transforms.py)
jrandom.MODE = prng.PARTITIONABLE
rng = np.random.default_rng(123)
n = 64
o = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
d = rng.normal(size=(n, 3)).astype(np.float32)
d /= np.linalg.norm(d, axis=1, keepdims=True)
prim = dict(cyl=(np.float32([-1, 0.5, 2]), np.float32([2, -0.5, 4]), 0.8), box=(np.float32([-1, -2, 1]), np.float32([1.5, 0.5, 3])),
            sph=(np.float32([0.5, 0.5, 3]), 1.7), tri=(np.float32([-4, -4, 3]), np.float32([4, -3, 3.5]), np.float32([0, 4, 2.5])))
th = np.deg2rad(30.0)
Rz = np.float32([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
out["unit/o"], out["unit/d"] = o, d
J = jnp.asarray
out["unit/cylinder"] = A([intersect_cylinder(J(o[i]), J(d[i]), J(prim["cyl"][0]), J(prim["cyl"][1]), prim["cyl"][2]) for i in range(n)])
out["unit/box"] = A([intersect_box(J(o[i]), J(d[i]), J(prim["box"][0]), J(prim["box"][1])) for i in range(n)])
out["unit/sphere"] = A([intersect_sphere(J(o[i]), J(d[i]), J(prim["sph"][0]), prim["sph"][1]) for i in range(n)])
out["unit/obox"] = A([intersect_oriented_box(J(o[i]), J(d[i]), J(np.float32([0.5, 0, 3])), J(np.float32([1.5, 0.6, 1.0])), J(Rz)) for i in range(n)])
out["unit/triangle"] = A([intersect_triangle(J(o[i]), J(d[i]), *(J(v) for v in prim["tri"])) for i in range(n)])
out["unit/plane"] = A([intersect_plane(J(o[i]), J(d[i]), J(np.float32([0.1, -0.2, 5])), euler_to_matrix(J(np.float32([3, -2, 20])))) for i in range(n)])
out["unit/euler"] = A([euler_to_matrix(J(e)) for e in np.float32([[0, 0, 0], [90, 0, 0], [10, -20, 30], [180, 0, 0], [-7.5, 12.25, 359]])])
rr, cc = reflect(J(d), J(np.roll(d, 1, axis=0)))
out["unit/reflect"], out["unit/reflect_cos"] = A(rr), A(cc)
surf = AsphericSurface(curvature=-0.05, conic=-1.0, aspheric=jnp.array([]))
oo = np.float32(o * [0.3, 0.3, 0] + [0, 0, 4])
dd = np.float32(d * [0.2, 0.2, 0] + [0, 0, -1]); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
res = [surf.intersect(J(oo[i]), J(dd[i]), J(np.float32([0.2, -0.1]))) for i in range(n)]
out["unit/surf_o"], out["unit/surf_d"] = oo, dd
out["unit/surf_t"] = A([r[0] for r in res]); out["unit/surf_pt"] = A([r[1] for r in res]); out["unit/surf_n"] = A([r[2] for r in res])
out["unit/conic_t"] = A([intersect_conic(J(oo[i]), J(dd[i]), 0.05, -1.0) for i in range(n)])
out["unit/random_normal_key4242_n8"] = A(jrandom.normal(jrandom.key(4242), (8,)))

np.savez_compressed(ROOT / "tests" / "golden" / "reference_golden.npz", **out)
(ROOT / "tests" / "golden" / "reference_golden.json").write_text(json.dumps(meta, indent=1, default=str))
print(f"wrote {len(out)} arrays,", (ROOT / "tests" / "golden" / "reference_golden.npz").stat().st_size, "bytes")
