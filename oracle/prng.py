"""Oracle restatement of the JAX PRNG surface the reference relies on.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

The reference draws every random number through ``jax.random`` -- a
third-party dependency that is NOT vendored in ``/root/reference`` and is not
version-pinned (``pyproject.toml:44`` ``jax>=0.4.20``).  Reference call sites:
``io/yaml_loader.py:44,76``; ``core/integrators.py:53,108,113,153,158``;
``utils/sampling.py:19-22,53-58``; ``core/reflection.py:35-37``;
``telescope/operations.py:36,186-188,220``.

What is restated here is the published algorithm of ``jax/_src/prng.py`` and
``jax/_src/random.py``:

* threefry2x32 (Salmon et al., SC'11; 20 rounds, rotations
  13,15,26,6 / 17,29,16,24, parity constant 0x1BD11BDA),
* ``key(seed)``, ``split``, ``bits``, ``uniform``, ``normal`` (via XLA's f32
  ``erf_inv`` -- Giles' polynomial) and ``choice(..., p=...)``,
* in both key-derivation modes: ``partitionable`` (default for JAX >= 0.5.0)
  and ``legacy`` (``jax_threefry_partitionable=False``, default before).

Pinned by the known-answer values in ``tests/test_oracle_prng.py`` (SURVEY.md
App. B).  ``choice`` is restated from memory of the upstream source and is not
covered by a public known-answer value.
"""
from __future__ import annotations

import numpy as np

PARTITIONABLE = "partitionable"
LEGACY = "legacy"

_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(k0, k1, c0, c1):
    """threefry2x32 block function on uint32 arrays (broadcasting)."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, dtype=_U32)
        k1 = np.asarray(k1, dtype=_U32)
        x0 = np.asarray(c0, dtype=_U32).copy()
        x1 = np.asarray(c1, dtype=_U32).copy()
        ks = (k0, k1, k0 ^ k1 ^ _U32(0x1BD11BDA))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + _U32(i + 1)
    return x0, x1


def key(seed: int) -> np.ndarray:
    """``jax.random.key(seed)`` / ``PRNGKey(seed)`` -> raw key words (hi, lo)."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=_U32)


def _legacy_bits_1d(k, n):
    """``threefry_2x32(key, iota(n))`` of legacy JAX: counts cut in halves."""
    cnt = np.arange(n, dtype=_U32)
    odd = n % 2
    if odd:
        cnt = np.concatenate([cnt, np.zeros(1, _U32)])
    h = cnt.size // 2
    y0, y1 = threefry2x32(k[0], k[1], cnt[:h], cnt[h:])
    out = np.concatenate([y0, y1])
    return out[:-1] if odd else out


def split(k, num: int = 2, mode: str = PARTITIONABLE) -> np.ndarray:
    """``jax.random.split(key, num)`` -> (num, 2) uint32."""
    k = np.asarray(k, dtype=_U32)
    if mode == PARTITIONABLE:
        i = np.arange(num, dtype=_U32)
        b0, b1 = threefry2x32(k[0], k[1], np.zeros(num, _U32), i)
        return np.stack([b0, b1], axis=-1)
    if mode == LEGACY:
        return _legacy_bits_1d(k, 2 * num).reshape(num, 2)
    raise ValueError(mode)


def bits(k, n: int, mode: str = PARTITIONABLE) -> np.ndarray:
    """32-bit ``random_bits(key, shape=(n,))``."""
    k = np.asarray(k, dtype=_U32)
    if mode == PARTITIONABLE:
        i = np.arange(n, dtype=_U32)
        b0, b1 = threefry2x32(k[0], k[1], np.zeros(n, _U32), i)
        return b0 ^ b1
    if mode == LEGACY:
        return _legacy_bits_1d(k, n)
    raise ValueError(mode)


def uniform(k, n: int, minval=0.0, maxval=1.0, mode: str = PARTITIONABLE) -> np.ndarray:
    """``jax.random.uniform(key, (n,), float32, minval, maxval)``."""
    b = bits(k, n, mode)
    fb = (b >> _U32(9)) | _U32(0x3F800000)
    f = fb.view(np.float32) - np.float32(1.0)
    lo = np.float32(minval)
    hi = np.float32(maxval)
    return np.maximum(lo, f * (hi - lo) + lo).astype(np.float32)


_ERFINV_LT5 = (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06,
               0.00021858087, -0.00125372503, -0.00417768164, 0.246640727, 1.50140941)
_ERFINV_GE5 = (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844,
               0.00573950773, -0.0076224613, 0.00943887047, 1.00167406, 2.83297682)


def erf_inv_f32(x: np.ndarray) -> np.ndarray:
    """XLA's float32 ``erf_inv`` (Giles' single-precision polynomial)."""
    x = np.asarray(x, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = -np.log1p(-(x * x)).astype(np.float32)
        lt = w < np.float32(5.0)
        w2 = np.where(lt, w - np.float32(2.5), np.sqrt(w) - np.float32(3.0)).astype(np.float32)
        p = np.where(lt, np.float32(_ERFINV_LT5[0]), np.float32(_ERFINV_GE5[0])).astype(np.float32)
        for a, b in zip(_ERFINV_LT5[1:], _ERFINV_GE5[1:]):
            p = (np.where(lt, np.float32(a), np.float32(b)) + p * w2).astype(np.float32)
        r = (p * x).astype(np.float32)
        r = np.where(np.abs(x) == np.float32(1.0), x * np.float32(np.inf), r)
    return r.astype(np.float32)


def normal(k, n: int, mode: str = PARTITIONABLE) -> np.ndarray:
    """``jax.random.normal(key, (n,), float32)``."""
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0), dtype=np.float32)
    u = uniform(k, n, lo, 1.0, mode)
    return (np.float32(np.sqrt(2)) * erf_inv_f32(u)).astype(np.float32)


def choice_p(k, p: np.ndarray, n: int, mode: str = PARTITIONABLE) -> np.ndarray:
    """``jax.random.choice(key, len(p), (n,), p=p)`` (replace=True).

    Upstream: ``p_cuml = cumsum(p); r = p_cuml[-1] * (1 - uniform(key, shape));
    ind = searchsorted(p_cuml, r)`` (side='left').
    """
    p = np.asarray(p, dtype=np.float32)
    cum = np.cumsum(p, dtype=np.float32)
    r = cum[-1] * (np.float32(1.0) - uniform(k, n, mode=mode))
    return np.searchsorted(cum, r, side="left").astype(np.int32)
