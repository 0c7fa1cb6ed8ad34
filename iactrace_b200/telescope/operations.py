"""Functional "return a new Telescope" edits (mirror of reference ``iactrace/telescope/operations.py``).

Every function leaves its input untouched and returns a shallow copy with the edited field;
derived device tables are rebuilt lazily on the next render.
"""
from __future__ import annotations

import math
from typing import Any

import torch

from .. import random as R
from .._util import f32, replace


def _with_group(telescope, group_idx, new_group):
    groups = list(telescope.mirror_groups)
    groups[group_idx] = new_group
    return replace(telescope, mirror_groups=groups)


# ---- mirror operations
def resample_mirrors(telescope, integrator, key):
    """Resample all mirror groups: ``keys = split(key, n_groups)`` (``operations.py:25-40``)."""
    keys = R.split(key, len(telescope.mirror_groups))
    return replace(telescope, mirror_groups=[integrator.sample_group(g, k) for g, k in zip(telescope.mirror_groups, keys)])


def set_mirror_positions(telescope, group_idx: int, positions):
    return _with_group(telescope, group_idx, replace(telescope.mirror_groups[group_idx], positions=f32(positions)))


def set_mirror_rotations(telescope, group_idx: int, rotations):
    return _with_group(telescope, group_idx, replace(telescope.mirror_groups[group_idx], rotations=f32(rotations)))


def scale_mirror_weights(telescope, group_idx: int, scale_factors):
    """Multiply the per-sample weights of each mirror by a factor (``operations.py:87-115``)."""
    g = telescope.mirror_groups[group_idx]
    s = f32(scale_factors)
    if s.ndim == 0:
        s = s.expand(len(g))
    return _with_group(telescope, group_idx, replace(g, weights=g.weights * s[:, None, None]))


def _sigma_rad(roughness):
    return roughness * math.pi / (180.0 * 3600.0)


def apply_roughness(telescope, roughness: float):
    """Roughness in arcsec for every mirror of every group (``operations.py:118-135``)."""
    groups = [replace(g, perturbation_scale=torch.full((len(g),), _sigma_rad(roughness), dtype=torch.float32,
                                                       device=g.positions.device)) for g in telescope.mirror_groups]
    return replace(telescope, mirror_groups=groups)


def apply_roughness_to_group(telescope, group_idx: int, roughness: float):
    g = telescope.mirror_groups[group_idx]
    scale = torch.full((len(g),), _sigma_rad(roughness), dtype=torch.float32, device=g.positions.device)
    return _with_group(telescope, group_idx, replace(g, perturbation_scale=scale))


def apply_misalignment_to_group(telescope, group_idx: int, sigma_h: float, sigma_v: float, key):
    """Gaussian tip/tilt misalignment in arcsec (``operations.py:161-198``): ``k1,k2 = split(key)``,
    N(0, sigma/3600 deg) added to rotations[:,0] and [:,1]."""
    g = telescope.mirror_groups[group_idx]
    n = len(g)
    k1, k2 = R.split(key)
    dh = R.normal(k1, n) * (sigma_h / 3600.0)
    dv = R.normal(k2, n) * (sigma_v / 3600.0)
    # out of place, like the reference's `.at[:, 0].add(...)`: the edit stays differentiable w.r.t. the original rotations
    dev = g.rotations.device
    rot = g.rotations + torch.stack([dh.to(dev), dv.to(dev), torch.zeros(n, dtype=torch.float32, device=dev)], dim=1)
    return _with_group(telescope, group_idx, replace(g, rotations=rot))


def apply_displacement_to_group(telescope, group_idx: int, sigma_z: float, key):
    """Gaussian z displacement of mirror positions (``operations.py:201-229``)."""
    g = telescope.mirror_groups[group_idx]
    dz = R.normal(key, len(g)) * sigma_z
    dev = g.positions.device
    zero = torch.zeros(len(g), dtype=torch.float32, device=dev)
    pos = g.positions + torch.stack([zero, zero, dz.to(dev)], dim=1)       # out of place: differentiable (see above)
    return _with_group(telescope, group_idx, replace(g, positions=pos))


def get_mirrors_by_stage(telescope, stage: int) -> list[int]:
    return [i for i, g in enumerate(telescope.mirror_groups) if g.optical_stage == stage]


def get_mirror_count(telescope) -> int:
    return sum(len(g) for g in telescope.mirror_groups)


# ---- sensor operations
def add_sensor(telescope, sensor):
    return replace(telescope, sensors=list(telescope.sensors) + [sensor])


def _check_sensor_idx(telescope, idx):
    if idx < 0 or idx >= len(telescope.sensors):
        raise IndexError(f"Sensor index {idx} out of range (0-{len(telescope.sensors) - 1})")


def replace_sensor(telescope, sensor, idx: int = 0):
    _check_sensor_idx(telescope, idx)
    sensors = list(telescope.sensors)
    sensors[idx] = sensor
    return replace(telescope, sensors=sensors)


def remove_sensor(telescope, idx: int = 0):
    _check_sensor_idx(telescope, idx)
    return replace(telescope, sensors=[s for i, s in enumerate(telescope.sensors) if i != idx])


def set_sensor_position(telescope, idx: int, position):
    sensors = list(telescope.sensors)
    sensors[idx] = replace(sensors[idx], position=f32(position))
    return replace(telescope, sensors=sensors)


def set_sensor_rotation(telescope, idx: int, rotation):
    sensors = list(telescope.sensors)
    sensors[idx] = replace(sensors[idx], rotation=f32(rotation))
    return replace(telescope, sensors=sensors)


def focus(telescope, delta_z: float, sensor_idx: int = 0):
    """Move a sensor along z (``operations.py:355-371``)."""
    old = telescope.sensors[sensor_idx].position
    dz = delta_z if isinstance(delta_z, torch.Tensor) else torch.tensor(float(delta_z), dtype=torch.float32)
    e_z = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float32, device=old.device)
    return set_sensor_position(telescope, sensor_idx, old + dz.to(old.device) * e_z)   # differentiable in delta_z and position


def get_sensor_count(telescope) -> int:
    return len(telescope.sensors)


# ---- obstruction operations
def add_obstruction(telescope, obstruction):
    return replace(telescope, obstruction_groups=list(telescope.obstruction_groups or []) + [obstruction])


def remove_obstruction(telescope, group_idx: int):
    if not telescope.obstruction_groups:
        raise IndexError("No obstruction groups to remove")
    if group_idx < 0 or group_idx >= len(telescope.obstruction_groups):
        raise IndexError(f"Obstruction group index {group_idx} out of range "
                         f"(0-{len(telescope.obstruction_groups) - 1})")
    return replace(telescope, obstruction_groups=[g for i, g in enumerate(telescope.obstruction_groups) if i != group_idx])


def clear_obstructions(telescope):
    return replace(telescope, obstruction_groups=[])


def get_obstruction_count(telescope) -> int:
    if not telescope.obstruction_groups:
        return 0
    return sum(len(g) for g in telescope.obstruction_groups)


# ---- convenience
def clone(telescope):
    """Independent copy: every tensor is cloned (``operations.py:463-472``)."""
    def cp(o):
        new = replace(o)
        for k, v in vars(new).items():
            if isinstance(v, torch.Tensor):
                object.__setattr__(new, k, v.detach().clone().requires_grad_(v.requires_grad))
        return new
    return replace(telescope, mirror_groups=[cp(g) for g in telescope.mirror_groups],
                   obstruction_groups=[cp(g) for g in (telescope.obstruction_groups or [])],
                   sensors=[cp(s) for s in telescope.sensors])


def get_info(telescope) -> dict[str, Any]:
    """Summary dict (``operations.py:475-542``)."""
    from ..sensors import HexagonalSensor, SquareSensor
    from .mirrors import AsphericDiskMirrorGroup, AsphericPolygonMirrorGroup
    stages, mirror_types = set(), []
    for g in telescope.mirror_groups:
        stages.add(g.optical_stage)
        mirror_types.append("disk" if isinstance(g, AsphericDiskMirrorGroup)
                            else "polygon" if isinstance(g, AsphericPolygonMirrorGroup) else "unknown")
    sensor_types = ["hexagonal" if isinstance(s, HexagonalSensor) else "square" if isinstance(s, SquareSensor)
                    else type(s).__name__ for s in telescope.sensors]
    if telescope.mirror_groups:
        allp = torch.cat([g.positions.detach() for g in telescope.mirror_groups], dim=0)
        bmin, bmax = allp.min(dim=0).values, allp.max(dim=0).values
    else:
        bmin = bmax = torch.zeros(3)
    return {"name": telescope.name, "n_mirror_groups": len(telescope.mirror_groups),
            "n_mirrors": get_mirror_count(telescope), "optical_stages": sorted(stages),
            "mirror_types": mirror_types, "n_sensors": len(telescope.sensors), "sensor_types": sensor_types,
            "n_obstruction_groups": len(telescope.obstruction_groups) if telescope.obstruction_groups else 0,
            "n_obstructions": get_obstruction_count(telescope), "bbox_min": bmin, "bbox_max": bmax}
