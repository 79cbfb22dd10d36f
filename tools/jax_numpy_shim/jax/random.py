"""Deterministic stand-ins so that constructors which draw initial values run.  The draws are
NOT jax.random's threefry streams: the fixture generator always passes explicit parameters
and inputs, and records no output that depends on a random draw of the reference."""
import numpy as _np


def PRNGKey(seed):
    return _np.array([0, int(seed) & 0xFFFFFFFF], dtype=_np.uint32)


key = PRNGKey


def _gen(k):
    return _np.random.default_rng([int(x) for x in _np.asarray(k).ravel()])


def split(k, num=2):
    words = _gen(k).integers(0, 2**32, size=(num, 2), dtype=_np.uint64).astype(_np.uint32)
    return words


def uniform(k, shape=(), dtype=None, minval=0.0, maxval=1.0):
    return _gen(k).uniform(minval, maxval, size=shape)


def normal(k=None, shape=(), dtype=None, key=None):
    return _gen(k if k is not None else key).normal(size=shape)


def choice(*a, **k):
    raise NotImplementedError("jax.random.choice: shots are outside the recorded cases")
