"""Minimal stand-in for the slice of Equinox the reference uses (see ../README.md)."""
import copy as _copy
import dataclasses as _dc


def field(*, static=False, default=_dc.MISSING, **_):
    return _dc.field(default=default, metadata={"static": static}) if default is not _dc.MISSING else _dc.field(metadata={"static": static})


class Module:
    """Plain class; subclasses that write no __init__ get a dataclass-style one from their annotations."""

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        if "__init__" not in cls.__dict__ and getattr(cls, "__annotations__", None) and not getattr(cls, "__abstractmethods__", None):
            has_custom = any("__init__" in b.__dict__ for b in cls.__mro__[1:] if b not in (Module, object))
            if not has_custom:
                _dc.dataclass(cls, eq=False, repr=False)


def is_array(x):
    import numpy as np
    return isinstance(x, np.ndarray)


def tree_at(where, pytree, replace):
    targets = where(pytree)
    multiple = isinstance(targets, tuple)
    tl = list(targets) if multiple else [targets]
    rl = list(replace) if multiple else [replace]
    idmap = {id(t): r for t, r in zip(tl, rl)}
    found = set()

    def rebuild(node):
        if id(node) in idmap:
            found.add(id(node))
            return idmap[id(node)]
        if isinstance(node, Module):
            new = _copy.copy(node)
            for k, v in vars(node).items():
                object.__setattr__(new, k, rebuild(v))
            return new
        if isinstance(node, list):
            return [rebuild(v) for v in node]
        if isinstance(node, tuple):
            return tuple(rebuild(v) for v in node)
        return node

    out = rebuild(pytree)
    assert len(found) == len(idmap), "tree_at: a selected node was not found in the tree"
    return out


def partition(*a, **k):  # pragma: no cover - filtering utilities are out of scope
    raise NotImplementedError


combine = filter_jit = filter_grad = partition
