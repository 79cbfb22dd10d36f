"""``jax.numpy`` on NumPy: same names, arrays carry the functional ``.at[idx].set / .add``."""
import numpy as _np
from numpy import *  # noqa: F401,F403
from numpy import linalg  # noqa: F401

_x64 = [True]
ndarray = _np.ndarray
complex64, complex128, float32, float64, int32, int64 = (
    _np.complex64, _np.complex128, _np.float32, _np.float64, _np.int32, _np.int64)
pi = _np.pi


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Ref:
            def set(self, value):
                out = _np.array(arr, copy=True)
                out[idx] = value
                return out.view(ShimArray)

            def add(self, value):
                out = _np.array(arr, copy=True)
                _np.add.at(out, idx, value)
                return out.view(ShimArray)

        return _Ref()


class ShimArray(_np.ndarray):
    @property
    def at(self):
        return _At(self)


def _wrap(fn):
    def inner(*a, **k):
        return _np.asarray(fn(*a, **k)).view(ShimArray)

    inner.__name__ = fn.__name__
    return inner


for _name in ("array", "asarray", "zeros", "ones", "eye", "zeros_like", "ones_like", "diag",
              "kron", "stack", "einsum", "outer", "transpose", "conj", "real", "sqrt", "exp",
              "cos", "sin", "abs", "dot", "sum", "trace", "round", "log", "arange", "reshape",
              "concatenate", "tensordot", "matmul", "where", "cumsum", "searchsorted"):
    globals()[_name] = _wrap(getattr(_np, _name))


def __getattr__(name):  # anything else: NumPy's
    return getattr(_np, name)
