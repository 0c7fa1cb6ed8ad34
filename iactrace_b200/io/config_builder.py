"""Programmatic authoring of telescope configs (the role of the reference's ``configs/builder.py``).

A ``TelescopeConfigBuilder`` collects plain dict entries in the YAML schema read by
``io.yaml_loader`` and can write them to a file, hand them to ``build_telescope`` directly, or both.
Beyond the reference's writer it covers the whole schema: optical stages and parent-surface offsets,
all five obstruction types, and the soft (differentiable) sensors.
"""
from __future__ import annotations

from pathlib import Path

import yaml


def _floats(seq):
    return [float(v) for v in seq]


class TelescopeConfigBuilder:
    def __init__(self, name: str, units: str = "m"):
        self.config = {"telescope": {"name": name, "units": units}, "mirror_templates": {}, "mirrors": [],
                       "obstructions": [], "sensors": []}

    # ---- mirrors
    def add_mirror_template(self, name, curvature, conic, aspheric_coeffs=()):
        self.config["mirror_templates"][name] = {
            "surface": {"curvature": float(curvature), "conic": float(conic), "aspheric": _floats(aspheric_coeffs)}}
        return self

    def add_mirror(self, mirror_id, template, position, orientation, aperture: dict, stage: int = 0, offset=None):
        if template not in self.config["mirror_templates"]:
            raise KeyError(f"unknown mirror template {template!r}")
        entry = {"id": mirror_id, "template": template, "position": _floats(position),
                 "orientation": _floats(orientation), "aperture": aperture}
        if stage:
            entry["stage"] = int(stage)
        if offset is not None:
            entry["offset"] = _floats(offset)
        self.config["mirrors"].append(entry)
        return self

    def add_mirror_circular(self, mirror_id, template, position, orientation, radius, **kw):
        return self.add_mirror(mirror_id, template, position, orientation, {"type": "circular", "radius": float(radius)}, **kw)

    def add_mirror_polygon(self, mirror_id, template, position, orientation, vertices, **kw):
        return self.add_mirror(mirror_id, template, position, orientation,
                               {"type": "polygon", "vertices": [_floats(v) for v in vertices]}, **kw)

    # ---- obstructions
    _OBSTRUCTION_FIELDS = {"cylinder": ("p1", "p2", "r"), "box": ("p1", "p2"), "sphere": ("center", "r"),
                           "oriented_box": ("center", "half_extents", "rotation"), "triangle": ("v0", "v1", "v2")}

    def add_obstruction(self, obs_id, kind: str, **fields):
        if kind not in self._OBSTRUCTION_FIELDS:
            raise ValueError(f"Unknown obstruction type: {kind}")
        want = self._OBSTRUCTION_FIELDS[kind]
        if set(fields) != set(want):
            raise ValueError(f"{kind} needs exactly the fields {want}")
        entry = {"id": obs_id, "type": kind}
        for k in want:
            v = fields[k]
            entry[k] = float(v) if k == "r" else ([_floats(row) for row in v] if k == "rotation" else _floats(v))
        self.config["obstructions"].append(entry)
        return self

    def add_obstruction_box(self, obs_id, p1, p2):
        return self.add_obstruction(obs_id, "box", p1=p1, p2=p2)

    def add_obstruction_cylinder(self, obs_id, p1, p2, radius):
        return self.add_obstruction(obs_id, "cylinder", p1=p1, p2=p2, r=radius)

    # ---- sensors
    def add_square_sensor_array(self, sensor_id, position, orientation, width, height, bounds, edge_width=None,
                                soft: dict | None = None):
        entry = {"id": sensor_id, "type": "differentiable_square" if soft else "square", "position": _floats(position),
                 "orientation": _floats(orientation), "width": int(width), "height": int(height), "bounds": _floats(bounds)}
        if edge_width is not None:
            entry["edge_width"] = float(edge_width)
        entry.update(soft or {})
        self.config["sensors"].append(entry)
        return self

    def add_hexagon_sensor_array(self, sensor_id, position, orientation, pixel_x, pixel_y, edge_width=None,
                                 soft: dict | None = None):
        entry = {"id": sensor_id, "type": "differentiable_hexagonal" if soft else "hexagonal",
                 "position": _floats(position), "orientation": _floats(orientation),
                 "centers_x": _floats(pixel_x), "centers_y": _floats(pixel_y)}
        if edge_width is not None:
            entry["edge_width"] = float(edge_width)
        entry.update(soft or {})
        self.config["sensors"].append(entry)
        return self

    # ---- output
    def to_dict(self) -> dict:
        return self.config

    def save(self, filename, precision: int | None = None):
        """Write the config as YAML; ``precision`` rounds floats to that many decimals first."""
        def rnd(node):
            if isinstance(node, float) and precision is not None:
                return round(node, precision)
            if isinstance(node, dict):
                return {k: rnd(v) for k, v in node.items()}
            if isinstance(node, list):
                return [rnd(v) for v in node]
            return node
        Path(filename).write_text(yaml.safe_dump(rnd(self.config), default_flow_style=False, sort_keys=False))
        return Path(filename)

    def build(self, integrator, key=None):
        from .yaml_loader import build_telescope
        return build_telescope(self.config, integrator, key)
