import numpy as _np

from . import numpy as _jnp


def tree_stack(items):
    """Stack a list of identical pytrees (tuples / None / arrays / scalars) leaf-wise."""
    if not items:
        return None
    first = items[0]
    if first is None:
        return None
    if isinstance(first, tuple):
        return tuple(tree_stack([it[i] for it in items]) for i in range(len(first)))
    if isinstance(first, list):
        return [tree_stack([it[i] for it in items]) for i in range(len(first))]
    return _jnp._wrap(_np.stack([_np.asarray(it) for it in items]))
