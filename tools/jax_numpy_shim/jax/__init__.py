"""NumPy stand-in for the slice of the JAX API that the reference's ``operations.py`` /
``simulation.py`` / ``tape.py`` touch.

FIXTURE TOOLING ONLY.  JAX is not installed in the build container, so the reference cannot
be imported as is; with this package in front of ``sys.path`` its UNMODIFIED source files
run on NumPy (float64 / complex128, eager, no tracing; ``vmap`` is a Python loop, ``jit`` the
identity with the ``lower().compile()`` surface) and ``tools/gen_golden_reference*.py`` record
their outputs as ``tests/golden/reference_*.npz``.  Nothing in the product, the
tests or the bench imports this directory; it is not a JAX re-implementation (no tracing, no
autodiff, XLA's summation order is not reproduced - irrelevant at the 1e-10 tolerance).
"""
import numpy as _np

from . import numpy  # noqa: F401  (jax.numpy)
from . import random  # noqa: F401


class _Config:
    x64_enabled = True
    jax_enable_x64 = True

    def update(self, name, value):
        if name == "jax_enable_x64":
            self.x64_enabled = self.jax_enable_x64 = bool(value)
            numpy._x64[0] = bool(value)


config = _Config()
Array = _np.ndarray


class _TreeUtil:
    @staticmethod
    def tree_leaves(tree):
        out = []

        def walk(t):
            if isinstance(t, (list, tuple)):
                for x in t:
                    walk(x)
            elif isinstance(t, dict):
                for x in t.values():
                    walk(x)
            elif t is not None:
                out.append(t)

        walk(tree)
        return out


class _Core:
    class Tracer:  # nothing is ever traced here
        pass


class _Lax:
    @staticmethod
    def index_in_dim(a, index, axis=0, keepdims=True):
        out = _np.take(_np.asarray(a), index, axis=axis)
        return _np.expand_dims(out, axis) if keepdims else out

    @staticmethod
    def dynamic_slice_in_dim(a, start, size, axis=0):
        idx = [slice(None)] * _np.ndim(a)
        idx[axis] = slice(int(start), int(start) + int(size))
        return _np.asarray(a)[tuple(idx)]


tree_util = _TreeUtil()
core = _Core()
lax = _Lax()


def clear_caches():
    pass


class _Jitted:
    """Eager function with the ahead-of-time surface the reference calls
    (`jit(f).lower(*args).compile()` -> callable)."""

    def __init__(self, fn):
        self._fn = fn

    def __call__(self, *a, **k):
        return self._fn(*a, **k)

    def lower(self, *a, **k):
        return self

    def compile(self):
        return self._fn


def jit(fn=None, **_kw):
    if fn is None:
        return lambda f: _Jitted(f)
    return _Jitted(fn)


def vmap(fn, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(_np.shape(a)[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = [fn(*[a if ax is None else _np.take(a, i, axis=ax) for a, ax in zip(args, axes)])
                for i in range(n)]
        return numpy.asarray(_np.stack(outs, axis=out_axes))

    return mapped
