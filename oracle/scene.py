"""Oracle scene model: YAML -> plain dict of NumPy arrays.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Follows:
* ``iactrace/io/yaml_loader.py:29-172``  load_telescope / build_telescope / _parse_*
* ``iactrace/telescope/mirrors.py:232-336`` group_mirrors / _group_by_surface_params
* ``iactrace/core/obstructions.py:258-278`` group_obstructions (order cyl, box, sphere, obox, tri)
* ``iactrace/sensors/square.py:43-61``   SquareSensor.__init__
* ``iactrace/sensors/hexagonal.py:61-144`` _detect_hex_grid / _build_lookup_table / HexagonalSensor.__init__
* ``iactrace/telescope/operations.py:118-229`` apply_roughness / misalignment / displacement

A scene is ``{"name", "groups": [...], "obstructions": [...], "sensors": [...]}``;
every float array is float32 exactly as the reference would hold it.
"""
from __future__ import annotations

import copy
from collections import defaultdict

import numpy as np

from . import prng, sample

SQRT3 = 1.7320508075688772
SQRT3_2 = 0.8660254037844386
SQRT3_3 = 0.5773502691896257
f32 = np.float32


# --------------------------------------------------------------------------- sensors
def make_square_sensor(position, rotation, width, height, bounds, edge_width=0.0):
    """square.py:43-61."""
    xmin, xmax, ymin, ymax = bounds
    return dict(type="square", position=np.asarray(position, f32), rotation=np.asarray(rotation, f32),
                width=int(width), height=int(height), x0=float(xmin), y0=float(ymin),
                dx=float((xmax - xmin) / width), dy=float((ymax - ymin) / height),
                edge_width=float(edge_width))


def rotate2d(x, y, angle, dt=f32):
    """hexagonal.py:16-19 (``angle`` is a Python float folded to ``dt``)."""
    c, s = dt(np.cos(dt(angle))), dt(np.sin(dt(angle)))
    return c * x - s * y, s * x + c * y


def cartesian_to_axial(x, y, size, dt=f32):
    """hexagonal.py:22-24."""
    return (dt(SQRT3_3) * x - y / dt(3)) / dt(size), (dt(2) * y / dt(3)) / dt(size)


def detect_hex_grid(centers):
    """hexagonal.py:61-80, float32 like the reference (argmin = first minimum)."""
    centers = np.asarray(centers, f32)
    n = len(centers)
    diff = centers[:, None] - centers[None, :]
    dist_sq = np.sum(diff ** 2, axis=2, dtype=f32)
    dist_sq[np.eye(n, dtype=bool)] = np.inf
    min_dist = np.sqrt(np.min(dist_sq))
    idx = int(np.argmin(dist_sq))
    vec = diff[idx // n, idx % n]
    angle = np.mod(np.arctan2(vec[1], vec[0]), f32(np.pi / 3))
    offset = centers[int(np.argmin(np.sum(centers ** 2, axis=1, dtype=f32)))]
    return f32(min_dist / f32(SQRT3)), f32(angle), offset


def build_lookup_table(centers, hex_size, rotation, offset):
    """hexagonal.py:83-101."""
    centers = np.asarray(centers, f32)
    x = centers[:, 0] - offset[0]
    y = centers[:, 1] - offset[1]
    xr, yr = rotate2d(x, y, -rotation)
    q, r = cartesian_to_axial(xr, yr, hex_size)
    qi = np.round(q).astype(np.int32)
    ri = np.round(r).astype(np.int32)
    q_min, q_max = int(qi.min()), int(qi.max())
    r_min, r_max = int(ri.min()), int(ri.max())
    table = np.full((q_max - q_min + 1, r_max - r_min + 1), -1, dtype=np.int32)
    table[qi - q_min, ri - r_min] = np.arange(len(centers), dtype=np.int32)
    return table, q_min, r_min


def make_hex_sensor(position, rotation, hex_centers, edge_width=0.0, grid=None):
    """hexagonal.py:121-144.  ``grid`` overrides the (tie-sensitive) detected grid
    constants so both sides of a parity run can be fed the same numbers."""
    centers = np.asarray(hex_centers, f32)
    if grid is None:
        size, rot, offset = detect_hex_grid(centers)
        hex_size = float(size)
        hex_inradius = float(f32(size * f32(SQRT3_2)))
        grid_rotation = float(rot)
        grid_offset = (float(offset[0]), float(offset[1]))
        table, q_min, r_min = build_lookup_table(centers, hex_size, grid_rotation, offset)
    else:
        hex_size, hex_inradius = grid["hex_size"], grid["hex_inradius"]
        grid_rotation, grid_offset = grid["grid_rotation"], tuple(grid["grid_offset"])
        table, q_min, r_min = np.asarray(grid["lookup_table"], np.int32), grid["q_min"], grid["r_min"]
    return dict(type="hexagonal", position=np.asarray(position, f32), rotation=np.asarray(rotation, f32),
                centers=centers, n_pixels=len(centers), edge_width=float(edge_width),
                hex_size=hex_size, hex_inradius=hex_inradius, grid_rotation=grid_rotation,
                grid_offset=grid_offset, lookup_table=table, q_min=q_min, r_min=r_min)


def make_soft_hex_sensor(hard, sigma=0.5, kernel_size=1):
    """hexagonal.py:197-258 DifferentiableHexagonalSensor built on a hard sensor's grid."""
    s = dict(hard)
    offs = [(q, r) for q in range(-kernel_size, kernel_size + 1)
            for r in range(-kernel_size, kernel_size + 1)
            if max(abs(q), abs(r), abs(-q - r)) <= kernel_size]
    s.update(type="soft_hexagonal", sigma=float(sigma),
             nb_q=np.array([o[0] for o in offs], np.int32), nb_r=np.array([o[1] for o in offs], np.int32))
    return s


def make_soft_square_sensor(position, rotation, width, height, bounds=(-1, 1, -1, 1), sigma=0.1, kernel_size=2):
    """square.py:94-141 DifferentiableSquareSensor."""
    s = make_square_sensor(position, rotation, width, height, bounds)
    rng = np.arange(-kernel_size, kernel_size + 1)
    ox, oy = np.meshgrid(rng, rng, indexing="xy")
    s.update(type="soft_square", sigma=float(sigma), kernel_size=int(kernel_size),
             offset_x=ox.ravel().astype(np.int32), offset_y=oy.ravel().astype(np.int32))
    return s


# --------------------------------------------------------------------------- mirrors
def make_mirror(position, rotation, curvature, conic, aspheric, aperture, stage=0, offset=(0.0, 0.0)):
    """mirrors.py:10-35 Mirror value type (aperture = ('disk', r) | ('polygon', verts))."""
    return dict(position=np.asarray(position, f32), rotation=np.asarray(rotation, f32),
                curvature=float(curvature), conic=float(conic),
                aspheric=np.asarray(aspheric, f32), aperture=aperture, stage=int(stage),
                offset=np.asarray(offset, f32))


def _empty_group(kind, stage, ms):
    n = len(ms)
    g = dict(kind=kind, stage=stage,
             positions=np.stack([m["position"] for m in ms]).astype(f32),
             rotations=np.stack([m["rotation"] for m in ms]).astype(f32),
             offsets=np.stack([m["offset"] for m in ms]).astype(f32),
             curvature=ms[0]["curvature"], conic=ms[0]["conic"], aspheric=ms[0]["aspheric"],
             points=np.zeros((n, 0, 3), f32), normals=np.zeros((n, 0, 3), f32),
             weights=np.zeros((n, 0, 1), f32), delta=np.zeros((n, 0, 3), f32),
             scale=np.zeros(n, f32))
    if kind == "disk":
        g["radii"] = np.array([m["aperture"][1] for m in ms], f32)
    else:
        g["vertices"] = np.stack([np.asarray(m["aperture"][1], f32) for m in ms])
    return g


def _by_surface(ms):
    """mirrors.py:316-336 (dict insertion order)."""
    d = defaultdict(list)
    for m in ms:
        d[(m["curvature"], m["conic"], tuple(np.asarray(m["aspheric"]).tolist()))].append(m)
    return d


def group_mirrors(mirrors):
    """mirrors.py:232-313: stage asc -> disk groups by surface -> polygon groups by n_verts, surface."""
    groups = []
    by_stage = defaultdict(list)
    for m in mirrors:
        by_stage[m["stage"]].append(m)
    for stage, sm in sorted(by_stage.items()):
        for _, ml in _by_surface([m for m in sm if m["aperture"][0] == "disk"]).items():
            groups.append(_empty_group("disk", stage, ml))
        by_nv = defaultdict(list)
        for m in sm:
            if m["aperture"][0] == "polygon":
                by_nv[len(m["aperture"][1])].append(m)
        for _, ml in by_nv.items():
            for _, mll in _by_surface(ml).items():
                groups.append(_empty_group("polygon", stage, mll))
    return groups


def group_obstructions(obs):
    """obstructions.py:258-278.  ``obs`` = list of (type, *arrays); returns grouped list in
    the fixed type order, each group a dict of stacked float32 arrays."""
    out = []
    cyl = [o for o in obs if o[0] == "cylinder"]
    box = [o for o in obs if o[0] == "box"]
    sph = [o for o in obs if o[0] == "sphere"]
    obx = [o for o in obs if o[0] == "oriented_box"]
    tri = [o for o in obs if o[0] == "triangle"]
    if cyl:
        out.append(dict(type="cylinder", p1=np.array([o[1] for o in cyl], f32),
                        p2=np.array([o[2] for o in cyl], f32), r=np.array([o[3] for o in cyl], f32)))
    if box:
        out.append(dict(type="box", p1=np.array([o[1] for o in box], f32), p2=np.array([o[2] for o in box], f32)))
    if sph:
        out.append(dict(type="sphere", centers=np.array([o[1] for o in sph], f32),
                        radii=np.array([o[2] for o in sph], f32)))
    if obx:
        out.append(dict(type="oriented_box", centers=np.array([o[1] for o in obx], f32),
                        half_extents=np.array([o[2] for o in obx], f32),
                        rotations=np.array([o[3] for o in obx], f32)))
    if tri:
        out.append(dict(type="triangle", v0=np.array([o[1] for o in tri], f32),
                        v1=np.array([o[2] for o in tri], f32), v2=np.array([o[3] for o in tri], f32)))
    return out


def parse_config(config):
    """yaml_loader.py:66-70,82-85,95-172 -> (name, mirrors, obstruction list, sensors)."""
    name = config.get("telescope", {}).get("name", "telescope")
    templates = config.get("mirror_templates", {})
    mirrors = []
    for m in config.get("mirrors", []):
        ap = m["aperture"]
        if ap["type"] == "circular":
            aperture = ("disk", float(ap["radius"]))
        elif ap["type"] == "polygon":
            aperture = ("polygon", np.asarray(ap["vertices"], f32))
        else:
            raise ValueError(f"Unknown aperture type: {ap['type']}")
        s = templates[m["template"]]["surface"]
        mirrors.append(make_mirror(m["position"], m["orientation"], s["curvature"], s["conic"],
                                   s.get("aspheric", []), aperture, m.get("stage", 0),
                                   m.get("offset", [0.0, 0.0])))
    obs = []
    for o in config.get("obstructions", []):
        t = o["type"]
        if t == "cylinder":
            obs.append((t, o["p1"], o["p2"], float(o["r"])))
        elif t == "box":
            obs.append((t, o["p1"], o["p2"]))
        elif t == "sphere":
            obs.append((t, o["center"], float(o["r"])))
        elif t == "oriented_box":
            obs.append((t, o["center"], o["half_extents"], o["rotation"]))
        elif t == "triangle":
            obs.append((t, o["v0"], o["v1"], o["v2"]))
        else:
            raise ValueError(f"Unknown obstruction type: {t}")
    sensors = []
    for s in config.get("sensors", []):
        ew = s.get("edge_width", 0.0)
        if s["type"] == "square":
            sensors.append(make_square_sensor(s["position"], s["orientation"], s["width"], s["height"],
                                              tuple(s["bounds"]), ew))
        elif s["type"] == "hexagonal":
            centers = np.array([s["centers_x"], s["centers_y"]], f32).T
            sensors.append(make_hex_sensor(s["position"], s["orientation"], centers, ew))
        else:
            raise ValueError(f"Unknown sensor type: {s['type']}")
    return name, mirrors, obs, sensors


def build_scene(config, n_samples, key=None, mode=prng.PARTITIONABLE, dt=f32):
    """yaml_loader.py:52-92 incl. the key chain ``key, subkey = split(key)`` per stage-0 group."""
    if key is None:
        key = prng.key(0)
    name, mirrors, obs, sensors = parse_config(config)
    groups = group_mirrors(mirrors)
    sampled = []
    for g in groups:
        if g["stage"] == 0:
            key, sub = prng.split(key, 2, mode)
            sampled.append(sample.sample_group(g, sub, n_samples, mode, dt))
        else:
            sampled.append(g)
    return dict(name=name, groups=sampled, obstructions=group_obstructions(obs), sensors=sensors)


def load_yaml(path, n_samples, key=None, mode=prng.PARTITIONABLE, dt=f32):
    """yaml_loader.py:29-49."""
    import yaml
    with open(path) as f:
        config = yaml.safe_load(f)
    return build_scene(config, n_samples, key, mode, dt)


# --------------------------------------------------------------------------- operations
def apply_roughness(scene, arcsec):
    """operations.py:118-135."""
    out = copy.copy(scene)
    sigma = f32(arcsec * np.pi / (180.0 * 3600.0))   # Python-double arithmetic, folded to f32 by jnp.full
    out["groups"] = [dict(g, scale=np.full(len(g["positions"]), sigma, f32)) for g in scene["groups"]]
    return out


def apply_misalignment_to_group(scene, gi, sigma_h, sigma_v, key, mode=prng.PARTITIONABLE):
    """operations.py:161-198."""
    out = copy.copy(scene)
    g = dict(scene["groups"][gi])
    n = len(g["positions"])
    k1, k2 = prng.split(key, 2, mode)
    dh = prng.normal(k1, n, mode) * f32(sigma_h / 3600.0)
    dv = prng.normal(k2, n, mode) * f32(sigma_v / 3600.0)
    rot = g["rotations"].copy()
    rot[:, 0] += dh
    rot[:, 1] += dv
    g["rotations"] = rot
    out["groups"] = list(scene["groups"])
    out["groups"][gi] = g
    return out


def apply_displacement_to_group(scene, gi, sigma_z, key, mode=prng.PARTITIONABLE):
    """operations.py:201-229."""
    out = copy.copy(scene)
    g = dict(scene["groups"][gi])
    n = len(g["positions"])
    dz = prng.normal(key, n, mode) * f32(sigma_z)
    pos = g["positions"].copy()
    pos[:, 2] += dz
    g["positions"] = pos
    out["groups"] = list(scene["groups"])
    out["groups"][gi] = g
    return out
