"""Host-side PRNG key bookkeeping compatible with ``jax.random`` keys.

Only key derivation lives on the host (a handful of threefry2x32 blocks per
``load_telescope``); every random *number* the path consumes is generated on
the GPU from these keys (``csrc/iact_sample.cu``).  Two key-derivation modes
exist because the reference does not pin its JAX version:
``partitionable`` (JAX >= 0.5.0 default, ours too) and ``legacy``.
"""
from __future__ import annotations

import numpy as np

PARTITIONABLE = "partitionable"
LEGACY = "legacy"

_mode = PARTITIONABLE


def set_rng_mode(mode: str) -> None:
    """Choose the JAX threefry key-derivation mode for everything that follows."""
    global _mode
    if mode not in (PARTITIONABLE, LEGACY):
        raise ValueError(f"unknown rng mode {mode!r}")
    _mode = mode


def get_rng_mode() -> str:
    return _mode


def mode_code(mode: str | None = None) -> int:
    return 0 if (mode or _mode) == PARTITIONABLE else 1


def _threefry(k0, k1, c0, c1):
    u = np.uint32
    rot = ((13, 15, 26, 6), (17, 29, 16, 24))
    with np.errstate(over="ignore"):
        x0 = np.asarray(c0, u) + u(k0)
        x1 = np.asarray(c1, u) + u(k1)
        ks = (u(k0), u(k1), u(k0) ^ u(k1) ^ u(0x1BD11BDA))
        for i in range(5):
            for r in rot[i % 2]:
                x0 = x0 + x1
                x1 = (x1 << u(r)) | (x1 >> u(32 - r))
                x1 = x1 ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + u(i + 1)
    return x0, x1


def key(seed: int) -> np.ndarray:
    """``jax.random.key(seed)`` -> raw threefry key words ``uint32[2]``."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


PRNGKey = key


def as_key(k) -> np.ndarray:
    """Accept None (-> key(0)), an int seed, two uint32 words, or a JAX key object."""
    if k is None:
        return key(0)
    if isinstance(k, (int, np.integer)):
        return key(int(k))
    try:  # typed JAX key
        import jax  # type: ignore
        if hasattr(k, "dtype") and jax.dtypes.issubdtype(k.dtype, jax.dtypes.prng_key):
            k = jax.random.key_data(k)
    except Exception:
        pass
    a = np.asarray(k)
    if a.shape != (2,):
        raise ValueError(f"PRNG key must be a seed or two uint32 words, got shape {a.shape}")
    return a.astype(np.uint32)


def split(k, num: int = 2, mode: str | None = None) -> np.ndarray:
    """``jax.random.split(key, num)`` -> ``uint32[num, 2]``."""
    k = as_key(k)
    mode = mode or _mode
    if mode == PARTITIONABLE:
        b0, b1 = _threefry(k[0], k[1], np.zeros(num, np.uint32), np.arange(num, dtype=np.uint32))
        return np.stack([b0, b1], axis=-1)
    cnt = np.arange(2 * num, dtype=np.uint32)
    y0, y1 = _threefry(k[0], k[1], cnt[:num], cnt[num:])
    return np.concatenate([y0, y1]).reshape(num, 2)


def normal(k, n: int, mode: str | None = None):
    """``jax.random.normal(key, (n,))`` evaluated on the GPU -> float32 CUDA tensor."""
    from . import _native as N
    torch = N.require_cuda()
    out = torch.empty(int(n), dtype=torch.float32, device="cuda")
    N.check(N.lib().iact_random_normal(N.key_arg(as_key(k)), mode_code(mode), int(n), N.ptr(out), N.stream_ptr()),
            "random.normal")
    return out


def uniform(k, n: int, minval: float = 0.0, maxval: float = 1.0, mode: str | None = None):
    """``jax.random.uniform(key, (n,), minval=, maxval=)`` evaluated on the GPU."""
    from . import _native as N
    torch = N.require_cuda()
    out = torch.empty(int(n), dtype=torch.float32, device="cuda")
    N.check(N.lib().iact_random_uniform(N.key_arg(as_key(k)), mode_code(mode), int(n), float(minval), float(maxval),
                                        N.ptr(out), N.stream_ptr()), "random.uniform")
    return out
