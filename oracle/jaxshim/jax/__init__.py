"""Minimal eager stand-in for the slice of JAX the reference uses (see ../README.md)."""
import numpy as _np

from . import numpy, random, lax, ops, tree_util  # noqa: F401
from .numpy import Array  # noqa: F401
from ._tree import tree_stack as _tree_stack

__version__ = "0.0-shim"


def jit(fun=None, static_argnames=None, static_argnums=None, **_):
    if fun is None:
        return lambda f: f
    return fun


class Dual:
    """Forward-mode dual number over float32 scalars (enough for jax.grad of scalar functions)."""
    __array_ufunc__ = None
    __array_priority__ = 1000

    def __init__(self, v, d):
        self.v, self.d = _np.float32(v), _np.float32(d)

    @staticmethod
    def _c(o):
        return o if isinstance(o, Dual) else Dual(o, 0.0)

    def __add__(self, o): o = Dual._c(o); return Dual(self.v + o.v, self.d + o.d)
    __radd__ = __add__
    def __sub__(self, o): o = Dual._c(o); return Dual(self.v - o.v, self.d - o.d)
    def __rsub__(self, o): o = Dual._c(o); return Dual(o.v - self.v, o.d - self.d)
    def __mul__(self, o): o = Dual._c(o); return Dual(self.v * o.v, self.d * o.v + self.v * o.d)
    __rmul__ = __mul__

    def __truediv__(self, o):
        o = Dual._c(o)
        with _np.errstate(all="ignore"):
            return Dual(self.v / o.v, self.d / o.v - self.v * o.d / (o.v * o.v))

    def __rtruediv__(self, o): return Dual._c(o) / self
    def __neg__(self): return Dual(-self.v, -self.d)

    def __pow__(self, p):
        with _np.errstate(all="ignore"):
            return Dual(self.v ** p, _np.float32(p) * self.v ** (p - 1) * self.d)

    def sqrt(self):
        with _np.errstate(all="ignore"):
            s = _np.sqrt(self.v)
            return Dual(s, self.d * (_np.float32(0.5) / s))

    def abs(self): return Dual(_np.abs(self.v), _np.sign(self.v) * self.d)


def grad(f):
    def df(x, *args):
        out = f(Dual(x, 1.0), *args)
        return out.d if isinstance(out, Dual) else _np.float32(0.0)
    return df


def vmap(f, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _np.asarray(a).shape[ax]
                break
        outs = []
        for i in range(n):
            call = [a if ax is None else numpy._wrap(_np.take(_np.asarray(a), i, axis=ax)) for a, ax in zip(args, axes)]
            outs.append(f(*call))
        if n == 0:
            raise ValueError("vmap over an empty axis is not supported by the shim")
        return _tree_stack(outs)
    return mapped
