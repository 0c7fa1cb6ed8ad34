"""jax.numpy stand-in: NumPy with JAX's defaults (float32 / int32, no 64-bit types, `.at[]`)."""
import numpy as _np

pi = _np.pi
inf = _np.inf
float32 = _np.float32
int32 = _np.int32
bool_ = _np.bool_
ndarray = _np.ndarray


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, v):
        out = _np.array(self.arr, copy=True)
        out[self.idx] = v
        return _wrap(out)

    def add(self, v):
        out = _np.array(self.arr, copy=True)
        _np.add.at(out, self.idx, v)
        return _wrap(out)


class Array(_np.ndarray):
    """ndarray with the functional-update `.at[idx].set/add` of jax.Array."""

    @property
    def at(self):
        return _At(self)

    def block_until_ready(self):
        return self


def _narrow(x):
    if isinstance(x, _np.ndarray):
        if x.dtype == _np.float64:
            x = x.astype(_np.float32)
        elif x.dtype == _np.int64:
            x = x.astype(_np.int32)
        return x.view(Array)
    if isinstance(x, _np.float64):
        return _np.float32(x)
    if isinstance(x, _np.int64):
        return _np.int32(x)
    return x


def _wrap(x):
    if isinstance(x, tuple):
        return tuple(_wrap(v) for v in x)
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return _narrow(x)


def _is_dual(x):
    return type(x).__name__ == "Dual"


def _lift(fn):
    def f(*a, **k):
        with _np.errstate(all="ignore"):
            return _wrap(fn(*a, **k))
    f.__name__ = getattr(fn, "__name__", "f")
    return f


def array(x, dtype=None):
    return _wrap(_np.array(x, dtype=dtype))


def asarray(x, dtype=None):
    return _wrap(_np.asarray(x, dtype=dtype))


def sqrt(x):
    if _is_dual(x):
        return x.sqrt()
    with _np.errstate(all="ignore"):
        return _wrap(_np.sqrt(x))


def abs(x):  # noqa: A001
    if _is_dual(x):
        return x.abs()
    return _wrap(_np.abs(x))


def where(c, a, b):
    return _wrap(_np.where(c, a, b))


for _name in ("zeros ones full stack concatenate sum minimum maximum min max dot cross einsum arange "
              "broadcast_to radians cos sin arctan2 arcsin mod eye round floor clip exp argmin meshgrid "
              "linspace roll cumsum sign ones_like zeros_like column_stack rad2deg deg2rad searchsorted "
              "isfinite all any tan log log1p expand_dims squeeze transpose reshape prod mean").split():
    globals()[_name] = _lift(getattr(_np, _name))


class _Linalg:
    norm = staticmethod(_lift(_np.linalg.norm))


linalg = _Linalg()
