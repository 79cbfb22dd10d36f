"""Counter-based PRNG keys for parameter init, GateError jitter and shot sampling.

The reference uses ``jax.random`` (threefry2x32) for ``random.key(seed)``
(model.py:199), ``split`` (model.py:665,1655,1674,1678; script.py:481;
utils.py:9-13), ``uniform`` (model.py:688-713), ``normal`` (unitary.py:238-245)
and ``choice`` (simulation.py:352).  This module restates the published
threefry2x32 construction (Salmon et al., SC'11; 20 rounds) and the
``jax.random`` derivations on top of it (partitionable layout, the default of
the pinned jax 0.9.0.1) in NumPy, vectorised over arrays of keys.

PARITY UNPINNED: jax is not installable in this image, so bit-equality with
``jax.random`` streams cannot be checked here.  What the
backend guarantees instead: the uniform stream is an explicit input of the shot
sampler, and given identical probabilities and uniforms the sampled indices and
int32 counts equal the oracle's bit for bit.

While a circuit is being recorded symbolically, a batched key argument is a
:class:`SymKey`; ``split`` extends its derivation path and ``normal`` allocates
one *noise leaf* per drawn value.  The actual numbers are generated per batch
element at run time by replaying the path on the concrete keys.
"""

from __future__ import annotations

import threading
from contextlib import contextmanager
from typing import List, Optional, Tuple

import numpy as np

from .symbolic import Sym, SymArray

_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, d):
    return (x << _U32(d)) | (x >> _U32(32 - d))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds.  All arguments uint32 arrays (broadcastable)."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, _U32)
        k1 = np.asarray(k1, _U32)
        ks = (k0, k1, k0 ^ k1 ^ _U32(0x1BD11BDA))
        x0 = np.asarray(x0, _U32) + ks[0]
        x1 = np.asarray(x1, _U32) + ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r) ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + _U32(i + 1)
    return x0, x1


def _threefry_scalar(k0: int, k1: int, x0: int, x1: int):
    """The same block function on Python ints - one scalar key is split on every
    ``Model.__call__`` and NumPy's per-call overhead on 2-element arrays dominated it."""
    M = 0xFFFFFFFF
    ks = (k0, k1, k0 ^ k1 ^ 0x1BD11BDA)
    x0 = (x0 + ks[0]) & M
    x1 = (x1 + ks[1]) & M
    for i in range(5):
        for r in _ROT[i % 2]:
            x0 = (x0 + x1) & M
            x1 = (((x1 << r) | (x1 >> (32 - r))) & M) ^ x0
        x0 = (x0 + ks[(i + 1) % 3]) & M
        x1 = (x1 + ks[(i + 2) % 3] + i + 1) & M
    return x0, x1


class PRNGKey:
    """An array of threefry keys: ``data`` is uint32 with shape ``(*batch, 2)``."""

    __slots__ = ("data",)

    def __init__(self, data):
        self.data = np.asarray(data, dtype=_U32)
        assert self.data.shape[-1] == 2

    @property
    def shape(self):
        return self.data.shape[:-1]

    @property
    def ndim(self):
        return self.data.ndim - 1

    @property
    def dtype(self):
        return "key<threefry2x32>"

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, idx):
        return PRNGKey(self.data[idx])

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __repr__(self):
        return f"PRNGKey(shape={self.shape})"


def key(seed: int) -> PRNGKey:
    """``jax.random.key(seed)``: (seed >> 32, seed & 0xffffffff)."""
    seed = int(seed)
    return PRNGKey(np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=_U32))


PRNGKeyFn = key


def _split_data(data: np.ndarray, num: int) -> np.ndarray:
    """(*batch, 2) -> (*batch, num, 2): key_i = threefry(key, (0, i))."""
    k0 = data[..., 0][..., None]
    k1 = data[..., 1][..., None]
    lo = np.arange(num, dtype=_U32)
    hi = np.zeros(num, dtype=_U32)
    a, b = threefry2x32(k0, k1, hi, lo)
    return np.stack([a, b], axis=-1)


def split(k, num: int = 2):
    """``jax.random.split``.  For a concrete key returns a PRNGKey of shape
    ``(num, *k.shape)``; for a :class:`SymKey` a tuple of derived SymKeys."""
    if isinstance(k, SymKey):
        return tuple(SymKey(k.arg, k.path + ((num, i),)) for i in range(num))
    if k.data.ndim == 1 and num <= 8:  # scalar key, few children: pure-int fast path
        k0, k1 = int(k.data[0]), int(k.data[1])
        return PRNGKey(np.array([_threefry_scalar(k0, k1, 0, i) for i in range(num)],
                                dtype=_U32))
    d = _split_data(k.data, num)  # (*batch, num, 2)
    return PRNGKey(np.moveaxis(d, -2, 0))


def safe_random_split(random_key, *args, **kwargs):
    """None-tolerant split (utils.py:9-13)."""
    if random_key is None:
        return None, None
    return split(random_key, *args, **kwargs)


def _bits64(data: np.ndarray, n: int) -> np.ndarray:
    """n uint64 words per key: (*batch, n)."""
    k0 = data[..., 0][..., None]
    k1 = data[..., 1][..., None]
    lo = np.arange(n, dtype=_U32)
    hi = np.zeros(n, dtype=_U32)
    a, b = threefry2x32(k0, k1, hi, lo)
    return (a.astype(np.uint64) << np.uint64(32)) | b.astype(np.uint64)


def _uniform01(data: np.ndarray, n: int) -> np.ndarray:
    """float64 uniforms in [0, 1): mantissa fill of 1.0, minus 1."""
    bits = _bits64(data, n)
    f = ((bits >> np.uint64(12)) | np.uint64(0x3FF0000000000000)).view(np.float64)
    return f - 1.0


def uniform(k: PRNGKey, shape=(), minval=0.0, maxval=1.0) -> np.ndarray:
    """``jax.random.uniform`` (float64) for a single key."""
    n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    u = _uniform01(k.data, n) * (maxval - minval) + minval
    u = np.maximum(minval, u)
    return u.reshape(shape) if len(shape) else u.reshape(())


def _normal_from_data(data: np.ndarray, n: int) -> np.ndarray:
    from scipy.special import erfinv

    lo = np.nextafter(-1.0, 0.0)
    u = _uniform01(data, n) * (1.0 - lo) + lo
    u = np.maximum(lo, u)
    return np.sqrt(2.0) * erfinv(u)


def normal(k, shape=()):
    """``jax.random.normal`` (float64).  Symbolic keys allocate noise leaves."""
    n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    if isinstance(k, SymKey):
        rec = _active_noise()
        if rec is None:
            raise RuntimeError("symbolic key used outside of a Script recording")
        cols = rec.allocate(k, n)
        if not len(shape):
            return Sym.leaf(rec.noise_arg, cols[0])
        arr = np.empty(n, dtype=object)
        for i, c in enumerate(cols):
            arr[i] = Sym.leaf(rec.noise_arg, c)
        return SymArray(arr.reshape(shape))
    z = _normal_from_data(k.data, n)
    return z.reshape(shape) if len(shape) else z.reshape(())


def choice_uniforms(k: PRNGKey, shots: int) -> np.ndarray:
    """Uniform stream a shot sampler consumes: one float64 in [0,1) per shot and key.
    Returns ``(*k.shape, shots)``."""
    return _uniform01(k.data, shots)


# ---------------------------------------------------------------------------
# symbolic keys + noise recipes
# ---------------------------------------------------------------------------
class SymKey:
    """A batched key argument during symbolic recording: argument slot + the
    chain of ``split`` selections ``((num, index), ...)`` applied so far."""

    __slots__ = ("arg", "path")

    def __init__(self, arg: int, path: Tuple[Tuple[int, int], ...] = ()):
        self.arg = arg
        self.path = path

    shape = ()

    def __repr__(self):
        return f"SymKey(arg={self.arg}, path={self.path})"


class NoiseRecorder:
    """Collects the normal draws a symbolic recording asked for."""

    def __init__(self, noise_arg: int):
        self.noise_arg = noise_arg
        self.recipes: List[Tuple[int, tuple, int]] = []  # (key arg, path, count)
        self.n_cols = 0

    def allocate(self, k: SymKey, n: int) -> List[int]:
        cols = list(range(self.n_cols, self.n_cols + n))
        self.recipes.append((k.arg, k.path, n))
        self.n_cols += n
        return cols

    def realise(self, key_args: dict, batch: int) -> np.ndarray:
        """(batch, n_cols) float64 matrix of the recorded draws, from the concrete
        per-element keys (``key_args[arg]`` is a PRNGKey of shape (batch,))."""
        out = np.empty((batch, self.n_cols), dtype=np.float64)
        col = 0
        for arg, path, n in self.recipes:
            data = key_args[arg].data  # (batch, 2)
            for num, idx in path:
                data = _split_data(data, num)[..., idx, :]
            out[:, col : col + n] = _normal_from_data(data, n)
            col += n
        return out


_tls = threading.local()


def _active_noise() -> Optional[NoiseRecorder]:
    st = getattr(_tls, "stack", None)
    return st[-1] if st else None


@contextmanager
def noise_recording(noise_arg: int):
    st = getattr(_tls, "stack", None)
    if st is None:
        st = _tls.stack = []
    rec = NoiseRecorder(noise_arg)
    st.append(rec)
    try:
        yield rec
    finally:
        st.pop()
