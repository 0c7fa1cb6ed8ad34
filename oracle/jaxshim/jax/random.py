"""jax.random stand-in on the oracle's threefry restatement (oracle/prng.py)."""
import numpy as _np

from oracle import prng as _prng
from . import numpy as _jnp

MODE = _prng.PARTITIONABLE     # jax_threefry_partitionable (JAX >= 0.5 default)


def key(seed):
    return _jnp._wrap(_prng.key(seed))


PRNGKey = key


def split(k, num=2):
    return _jnp._wrap(_prng.split(_np.asarray(k, _np.uint32), num, MODE))


def _n(shape):
    shape = tuple(shape) if not isinstance(shape, int) else (shape,)
    return int(_np.prod(shape)) if shape else 1, shape


def uniform(k, shape=(), dtype=None, minval=0.0, maxval=1.0):
    n, shape = _n(shape)
    return _jnp._wrap(_prng.uniform(_np.asarray(k, _np.uint32), n, minval, maxval, MODE).reshape(shape))


def normal(k, shape=(), dtype=None):
    n, shape = _n(shape)
    return _jnp._wrap(_prng.normal(_np.asarray(k, _np.uint32), n, MODE).reshape(shape))


def choice(k, a, shape=(), replace=True, p=None):
    n, shape = _n(shape)
    assert replace and p is not None and isinstance(a, int)
    return _jnp._wrap(_prng.choice_p(_np.asarray(k, _np.uint32), _np.asarray(p), n, MODE).reshape(shape))
