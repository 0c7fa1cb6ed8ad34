"""``squareshow`` / ``hexshow`` (reference ``iactrace/viz/plotting.py:6-82``), import-guarded."""
from __future__ import annotations

import numpy as np


def _plt():
    try:
        import matplotlib.pyplot as plt
        return plt
    except ImportError as e:  # pragma: no cover
        raise ImportError("matplotlib is required for iactrace_b200.viz") from e


def _np(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


def squareshow(image, sensor, ax=None, **kwargs):
    """Show a square-sensor image with physical extents."""
    plt = _plt()
    ax = ax or plt.gca()
    ext = [sensor.x0, sensor.x0 + sensor.dx * sensor.width, sensor.y0, sensor.y0 + sensor.dy * sensor.height]
    ax.imshow(_np(image), origin="lower", extent=ext, **kwargs)
    return ax


def hexshow(image, sensor, ax=None, **kwargs):
    """Show a hexagonal-sensor image as a collection of hexagons."""
    plt = _plt()
    from matplotlib.collections import RegularPolyCollection
    ax = ax or plt.gca()
    c = _np(sensor.hex_centers)
    coll = RegularPolyCollection(6, rotation=sensor.grid_rotation, sizes=(1,), offsets=c,
                                 transOffset=ax.transData, **kwargs)
    coll.set_array(_np(image))
    ax.add_collection(coll)
    ax.set_xlim(c[:, 0].min() - sensor.hex_size, c[:, 0].max() + sensor.hex_size)
    ax.set_ylim(c[:, 1].min() - sensor.hex_size, c[:, 1].max() + sensor.hex_size)
    ax.set_aspect("equal")
    return ax
