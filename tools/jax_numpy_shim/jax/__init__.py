"""NumPy stand-in for the slice of the JAX API that the reference's ``operations.py`` /
``simulation.py`` / ``tape.py`` touch.

FIXTURE TOOLING ONLY.  JAX is not installed in the build container, so the reference cannot
be imported as is; with this package in front of ``sys.path`` its UNMODIFIED source files
run on NumPy (float64 / complex128, eager, no tracing) and ``tools/gen_golden_reference.py``
records their outputs as ``tests/golden/reference_sim.npz``.  Nothing in the product, the
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


def jit(fn=None, **_kw):
    if fn is None:
        return lambda f: f
    return fn


def vmap(fn, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(_np.shape(a)[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = [fn(*[a if ax is None else _np.take(a, i, axis=ax) for a, ax in zip(args, axes)])
                for i in range(n)]
        return numpy.asarray(_np.stack(outs, axis=out_axes))

    return mapped
