import copy as _copy

import numpy as _np


class GetAttrKey:
    def __init__(self, name):
        self.name = name


class SequenceKey:
    def __init__(self, idx):
        self.idx = idx


class DictKey:
    def __init__(self, key):
        self.key = key


def tree_map(f, tree, *rest):
    from equinox import Module
    if isinstance(tree, Module):
        new = _copy.copy(tree)
        for k, v in vars(tree).items():
            object.__setattr__(new, k, tree_map(f, v))
        return new
    if isinstance(tree, list):
        return [tree_map(f, v) for v in tree]
    if isinstance(tree, tuple):
        return tuple(tree_map(f, v) for v in tree)
    if isinstance(tree, dict):
        return {k: tree_map(f, v) for k, v in tree.items()}
    if tree is None:
        return None
    if isinstance(tree, (_np.ndarray, _np.generic, float, int)):
        return f(tree)
    return tree


def tree_map_with_path(f, tree, *rest):  # pragma: no cover - filtering utilities are out of scope
    raise NotImplementedError
