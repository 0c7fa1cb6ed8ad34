"""Inspection helpers (the role of reference ``iactrace/utils/filtering.py:120-129`` show_structure)."""
from __future__ import annotations

import torch


def iter_leaves(telescope):
    """Yield ``(path, tensor)`` for every tensor leaf, with the path strings the reference's glob
    patterns use (e.g. ``mirror_groups.0.rotations``)."""
    for attr in ("mirror_groups", "obstruction_groups", "sensors"):
        for i, obj in enumerate(getattr(telescope, attr) or []):
            for k, v in vars(obj).items():
                if isinstance(v, torch.Tensor) and not k.startswith("_"):
                    yield f"{attr}.{i}.{k}", v


def show_structure(telescope) -> None:
    """Print the telescope's tensor leaves with shapes, dtypes and whether they carry gradients."""
    print("Model structure:")
    for path, t in iter_leaves(telescope):
        print(f"  {path}: {tuple(t.shape)} {str(t.dtype).replace('torch.', '')}{' (requires_grad)' if t.requires_grad else ''}")


def trainable(telescope, pattern: str):
    """Mark the leaves whose path matches the glob ``pattern`` (``*`` = one path element) as trainable
    and return them: the analogue of choosing leaves with ``eqx.partition`` in the reference."""
    import fnmatch
    out = []
    for path, t in iter_leaves(telescope):
        if fnmatch.fnmatchcase(path, pattern) and t.is_floating_point():
            t.requires_grad_(True)
            out.append((path, t))
    return out
