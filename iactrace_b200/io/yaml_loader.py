"""YAML -> Telescope (mirror of reference ``iactrace/io/yaml_loader.py``; same schema, same
exceptions, same PRNG key chain: ``key, subkey = split(key)`` per stage-0 group)."""
from __future__ import annotations

from pathlib import Path
from typing import Any

import numpy as np
import yaml

from .. import random as R
from ..core import (AsphericSurface, Box, Cylinder, DiskAperture, OrientedBox, PolygonAperture, Sphere, Triangle,
                    group_obstructions)
from ..sensors import (DifferentiableHexagonalSensor, DifferentiableSquareSensor, HexagonalSensor, SquareSensor)
from ..telescope import Mirror, Telescope, group_mirrors

try:  # libyaml makes the 32 k-line CT5 file load in well under a second
    _Loader = yaml.CSafeLoader
except AttributeError:  # pragma: no cover
    _Loader = yaml.SafeLoader


def load_telescope(filename: str | Path, integrator, key=None) -> Telescope:
    """Load a telescope from a YAML configuration file; ``key`` defaults to ``key(0)``."""
    if key is None:
        key = R.key(0)
    with open(filename, "r") as f:
        config = yaml.load(f, Loader=_Loader)
    return build_telescope(config, integrator, key)


def build_telescope(config: dict[str, Any], integrator, key) -> Telescope:
    """Build a telescope from a parsed config dict.  ``integrator=None`` skips sampling (extension:
    host-only inspection of a scene on a machine without a GPU)."""
    name = config.get("telescope", {}).get("name", "telescope")
    templates = config.get("mirror_templates", {})
    mirrors = [_parse_mirror(m, templates) for m in config.get("mirrors", [])]
    mirror_groups = group_mirrors(mirrors)
    key = R.as_key(key)
    sampled = []
    for group in mirror_groups:
        if group.optical_stage == 0 and integrator is not None:
            key, subkey = R.split(key)
            sampled.append(integrator.sample_group(group, subkey))
        else:
            sampled.append(group)   # stage >= 1 mirrors are intersected, not sampled
    obstructions = [_parse_obstruction(o) for o in config.get("obstructions", [])]
    sensors = [_parse_sensor(s) for s in config.get("sensors", [])]
    return Telescope(mirror_groups=sampled, obstruction_groups=group_obstructions(obstructions), sensors=sensors,
                     name=name)


def _parse_mirror(m, templates) -> Mirror:
    aperture = _parse_aperture(m["aperture"])
    surface = AsphericSurface.from_template(templates[m["template"]])
    return Mirror(position=m["position"], rotation=m["orientation"], surface=surface, aperture=aperture,
                  optical_stage=m.get("stage", 0), offset=m.get("offset", [0.0, 0.0]))


def _parse_aperture(config):
    atype = config["type"]
    if atype == "circular":
        return DiskAperture(config["radius"])
    if atype == "polygon":
        return PolygonAperture(config["vertices"])
    raise ValueError(f"Unknown aperture type: {atype}")


def _parse_obstruction(config):
    otype = config["type"]
    if otype == "cylinder":
        return Cylinder(config["p1"], config["p2"], config["r"])
    if otype == "box":
        return Box(config["p1"], config["p2"])
    if otype == "sphere":
        return Sphere(config["center"], config["r"])
    if otype == "oriented_box":
        return OrientedBox(config["center"], config["half_extents"], np.array(config["rotation"]))
    if otype == "triangle":
        return Triangle(config["v0"], config["v1"], config["v2"])
    raise ValueError(f"Unknown obstruction type: {otype}")


def _parse_sensor(config):
    stype = config["type"]
    edge_width = config.get("edge_width", 0.0)
    if stype == "square":
        return SquareSensor(position=config["position"], rotation=config["orientation"], width=config["width"],
                            height=config["height"], bounds=tuple(config["bounds"]), edge_width=edge_width)
    if stype == "hexagonal":
        centers = np.array([config["centers_x"], config["centers_y"]], dtype=np.float32).T
        return HexagonalSensor(position=config["position"], rotation=config["orientation"], hex_centers=centers,
                               edge_width=edge_width)
    # extension over the reference schema: the soft (Gaussian-splat) sensors are YAML-constructible
    if stype == "differentiable_square":
        return DifferentiableSquareSensor(position=config["position"], rotation=config["orientation"],
                                          width=config["width"], height=config["height"],
                                          bounds=tuple(config.get("bounds", (-1, 1, -1, 1))),
                                          sigma=config.get("sigma", 0.1), kernel_size=config.get("kernel_size", 2))
    if stype == "differentiable_hexagonal":
        centers = np.array([config["centers_x"], config["centers_y"]], dtype=np.float32).T
        return DifferentiableHexagonalSensor(position=config["position"], rotation=config["orientation"],
                                             hex_centers=centers, sigma=config.get("sigma", 0.5),
                                             kernel_size=config.get("kernel_size", 1))
    raise ValueError(f"Unknown sensor type: {stype}")
