"""The boundary object (mirror of reference ``iactrace/telescope/telescope.py``)."""
from __future__ import annotations

from typing import Any

from ..core.render import render, render_debug
from . import operations as ops


class Telescope:
    """IACT telescope configuration: mirror groups, obstruction groups, sensors.

    Immutable by convention: every edit method returns a new Telescope.  Calling the object
    renders sources onto a sensor with the CUDA trace kernel.
    """

    def __init__(self, mirror_groups, obstruction_groups=None, sensors=None, name: str = "telescope") -> None:
        self.mirror_groups = mirror_groups
        self.obstruction_groups = obstruction_groups
        self.sensors = list(sensors) if sensors else []
        self.name = name
        self._cache: dict = {}

    def __call__(self, sources, values, source_type="point", sensor_idx: int = 0, debug: bool = False):
        """Render sources (N,3) with fluxes (N,) -> image, or raw (pts, values) hits if ``debug``.

        ``sources`` are positions for ``'point'`` and propagation directions for ``'parallel'``.
        Shadowing applies whenever ``obstruction_groups`` is non-empty.
        """
        if debug:
            return render_debug(self, sources, values, source_type, sensor_idx)
        return render(self, sources, values, source_type, sensor_idx)

    @classmethod
    def from_yaml(cls, filename, integrator, key=None) -> "Telescope":
        from ..io.yaml_loader import load_telescope
        return load_telescope(filename, integrator, key)

    # convenience wrappers over telescope.operations
    def resample_mirrors(self, integrator, key): return ops.resample_mirrors(self, integrator, key)
    def set_mirror_positions(self, group_idx, positions): return ops.set_mirror_positions(self, group_idx, positions)
    def set_mirror_rotations(self, group_idx, rotations): return ops.set_mirror_rotations(self, group_idx, rotations)
    def scale_mirror_weights(self, group_idx, scale_factors): return ops.scale_mirror_weights(self, group_idx, scale_factors)
    def apply_roughness(self, roughness_arcsec): return ops.apply_roughness(self, roughness_arcsec)
    def apply_roughness_to_group(self, group_idx, roughness): return ops.apply_roughness_to_group(self, group_idx, roughness)
    def apply_misalignment_to_group(self, group_idx, sigma_h, sigma_v, key): return ops.apply_misalignment_to_group(self, group_idx, sigma_h, sigma_v, key)
    def apply_displacement_to_group(self, group_idx, sigma_z, key): return ops.apply_displacement_to_group(self, group_idx, sigma_z, key)
    def get_mirrors_by_stage(self, stage): return ops.get_mirrors_by_stage(self, stage)
    def get_mirror_count(self): return ops.get_mirror_count(self)
    def add_sensor(self, sensor): return ops.add_sensor(self, sensor)
    def replace_sensor(self, sensor, idx=0): return ops.replace_sensor(self, sensor, idx)
    def remove_sensor(self, idx=0): return ops.remove_sensor(self, idx)
    def set_sensor_position(self, idx, position): return ops.set_sensor_position(self, idx, position)
    def set_sensor_rotation(self, idx, rotation): return ops.set_sensor_rotation(self, idx, rotation)
    def focus(self, delta_z, sensor_idx=0): return ops.focus(self, delta_z, sensor_idx)
    def get_sensor_count(self): return ops.get_sensor_count(self)
    def add_obstruction(self, obstruction): return ops.add_obstruction(self, obstruction)
    def remove_obstruction(self, group_idx): return ops.remove_obstruction(self, group_idx)
    def clear_obstructions(self): return ops.clear_obstructions(self)
    def get_obstruction_count(self): return ops.get_obstruction_count(self)
    def clone(self): return ops.clone(self)
    def get_info(self) -> dict[str, Any]: return ops.get_info(self)
