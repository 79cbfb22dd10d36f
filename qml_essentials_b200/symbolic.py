"""Affine symbolic scalars used to record a batched circuit ONCE.

The reference traces the circuit function under ``jax.vmap`` with tracer
arguments (script.py:302-315).  Without a tracing compiler we record the tape
once with *affine proxies*: every batched argument is replaced by an array of
``Sym`` leaves, and every gate angle the circuit computes from them stays an
affine form ``c0 + sum_i c_i * leaf_i`` (true for every angle the reference
builds: model.py:744,812; ansaetze.py:934,959; unitary.py:238).  The compiler
turns those forms into the device program's angle table, so the kernels read
``params`` / ``inputs`` once from HBM and index the (inputs x params x pulse)
batch factors in place instead of materialising the repeated batch
(model.py:1452-1481).

Anything non-affine raises :class:`SymbolicError`; ``Script`` then falls back
to concrete per-element recording (still executed on the GPU).
"""

from __future__ import annotations

import numbers
from typing import Dict, Tuple

import numpy as np


class SymbolicError(TypeError):
    """Raised when a recorded value is used in a way an affine proxy cannot follow."""


class NonAffineProduct(SymbolicError):
    """Two proxies were multiplied.  ``args`` lists the argument slots involved so
    the recorder can retry with a broadcast argument baked in as a constant."""

    def __init__(self, args):
        super().__init__("product of two traced values is not affine")
        self.args_involved = tuple(sorted(set(args)))


Leaf = Tuple[int, int]  # (argument slot, flat offset inside one batch element)


def _is_number(x) -> bool:
    return isinstance(x, (numbers.Real, np.floating, np.integer)) and not isinstance(
        x, bool
    )


class Sym:
    """Affine form over batch-argument leaves."""

    __slots__ = ("terms", "const")
    __array_ufunc__ = None  # make numpy scalars/arrays defer to our reflected ops
    __array_priority__ = 1000
    shape = ()
    ndim = 0
    size = 1

    def __init__(self, terms: Dict[Leaf, float] = None, const: float = 0.0):
        self.terms = terms or {}
        self.const = float(const)

    @staticmethod
    def leaf(arg: int, offset: int) -> "Sym":
        return Sym({(arg, offset): 1.0}, 0.0)

    # -- arithmetic ------------------------------------------------------
    def _lift(self, other):
        if isinstance(other, Sym):
            return other
        if _is_number(other):
            return Sym({}, float(other))
        if isinstance(other, np.ndarray) and other.ndim == 0:
            return Sym({}, float(other))
        return None

    def __add__(self, other):
        if isinstance(other, (np.ndarray, SymArray)) and getattr(other, "ndim", 0) > 0:
            return SymArray._wrap(other).__radd__(self)
        o = self._lift(other)
        if o is None:
            return NotImplemented
        terms = dict(self.terms)
        for k, v in o.terms.items():
            nv = terms.get(k, 0.0) + v
            if nv == 0.0:
                terms.pop(k, None)
            else:
                terms[k] = nv
        return Sym(terms, self.const + o.const)

    __radd__ = __add__

    def __neg__(self):
        return Sym({k: -v for k, v in self.terms.items()}, -self.const)

    def __pos__(self):
        return self

    def __sub__(self, other):
        o = self._lift(other)
        if o is None:
            if isinstance(other, (np.ndarray, SymArray)):
                return (-SymArray._wrap(other)).__radd__(self)
            return NotImplemented
        return self + (-o)

    def __rsub__(self, other):
        return (-self) + other

    def __mul__(self, other):
        if isinstance(other, Sym):
            if not other.terms:
                other = other.const
            elif not self.terms:
                return other * self.const
            else:
                raise NonAffineProduct(
                    [a for a, _ in self.terms] + [a for a, _ in other.terms]
                )
        if isinstance(other, (np.ndarray, SymArray)) and getattr(other, "ndim", 0) > 0:
            return SymArray._wrap(other).__rmul__(self)
        if isinstance(other, np.ndarray):
            other = float(other)
        if not _is_number(other):
            return NotImplemented
        c = float(other)
        if c == 0.0:
            return Sym({}, 0.0)
        return Sym({k: v * c for k, v in self.terms.items()}, self.const * c)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Sym) and not other.terms:
            other = other.const
        if isinstance(other, np.ndarray) and other.ndim == 0:
            other = float(other)
        if not _is_number(other):
            raise SymbolicError("division by a batched value is not affine")
        return self * (1.0 / float(other))

    def __rtruediv__(self, other):
        raise SymbolicError("division by a batched value is not affine")

    def __pow__(self, other):
        raise SymbolicError("power of a batched value is not affine")

    # -- things an affine proxy cannot answer ----------------------------
    def __float__(self):
        if not self.terms:
            return self.const
        raise SymbolicError("batched value has no concrete float")

    def __bool__(self):
        raise SymbolicError("truth value of a batched value is undefined while recording")

    def _cmp(self, other):
        raise SymbolicError("comparison of a batched value is undefined while recording")

    __lt__ = __le__ = __gt__ = __ge__ = _cmp

    def __array__(self, *a, **k):
        raise SymbolicError("batched value cannot be converted to a numpy array")

    def reshape(self, *shape):
        return SymArray(np.array([self], dtype=object)).reshape(*shape)

    @property
    def is_constant(self) -> bool:
        return not self.terms

    def __repr__(self):
        body = " + ".join(f"{v:g}*a{a}[{o}]" for (a, o), v in sorted(self.terms.items()))
        return f"Sym({self.const:g}{' + ' + body if body else ''})"


class SymArray:
    """N-d array of :class:`Sym` with the small numpy surface circuits use."""

    __array_ufunc__ = None
    __array_priority__ = 1000

    def __init__(self, data: np.ndarray):
        self._d = data

    @staticmethod
    def leaves(arg: int, shape) -> "SymArray":
        n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
        flat = np.empty(n, dtype=object)
        for i in range(n):
            flat[i] = Sym.leaf(arg, i)
        return SymArray(flat.reshape(shape))

    @staticmethod
    def _wrap(x) -> "SymArray":
        if isinstance(x, SymArray):
            return x
        a = np.asarray(x)
        out = np.empty(a.shape, dtype=object)
        for idx in np.ndindex(a.shape):
            v = a[idx]
            out[idx] = v if isinstance(v, Sym) else Sym({}, float(v))
        return SymArray(out)

    # -- structure -------------------------------------------------------
    @property
    def shape(self):
        return self._d.shape

    @property
    def ndim(self):
        return self._d.ndim

    @property
    def size(self):
        return self._d.size

    @property
    def T(self):
        return SymArray(self._d.T)

    def __len__(self):
        return len(self._d)

    def __iter__(self):
        for i in range(len(self._d)):
            yield self[i]

    def __getitem__(self, idx):
        r = self._d[idx]
        if isinstance(r, np.ndarray):
            if r.ndim == 0:
                return r.item()
            return SymArray(r)
        return r

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return SymArray(self._d.reshape(shape))

    def flatten(self):
        return SymArray(self._d.flatten())

    ravel = flatten

    def squeeze(self, axis=None):
        r = self._d.squeeze(axis)
        return r.item() if r.ndim == 0 else SymArray(r)

    def take(self, indices, axis=None):
        return SymArray(np.take(self._d, np.asarray(indices), axis=axis))

    def repeat(self, n, axis=None):
        return SymArray(np.repeat(self._d, n, axis=axis))

    def sum(self, axis=None):
        r = np.sum(self._d, axis=axis)
        return SymArray(r) if isinstance(r, np.ndarray) and r.ndim else (
            r.item() if isinstance(r, np.ndarray) else r)

    def mean(self, axis=None):
        n = self._d.size if axis is None else self._d.shape[axis]
        return self.sum(axis=axis) / float(n)

    def any(self):
        raise SymbolicError("truth value of a batched array is undefined while recording")

    all = any

    # -- elementwise arithmetic -------------------------------------------
    def _binary(self, other, fn):
        if isinstance(other, SymArray):
            o = other._d
        elif isinstance(other, np.ndarray):
            o = other.astype(object) if other.dtype != object else other
        else:
            o = other
        a, b = np.broadcast_arrays(self._d, np.asarray(o, dtype=object))
        out = np.empty(a.shape, dtype=object)
        for idx in np.ndindex(a.shape):
            x, y = a[idx], b[idx]
            if not isinstance(y, Sym) and not _is_number(y):
                y = float(y)
            out[idx] = fn(x, y)
        return SymArray(out)

    def __add__(self, o):
        return self._binary(o, lambda x, y: x + y)

    __radd__ = __add__

    def __sub__(self, o):
        return self._binary(o, lambda x, y: x - y)

    def __rsub__(self, o):
        return self._binary(o, lambda x, y: y - x)

    def __mul__(self, o):
        return self._binary(o, lambda x, y: x * y)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._binary(o, lambda x, y: x / y)

    def __neg__(self):
        return SymArray(np.vectorize(lambda x: -x, otypes=[object])(self._d))

    def __array__(self, *a, **k):
        raise SymbolicError("batched array cannot be converted to a numpy array")

    def __repr__(self):
        return f"SymArray(shape={self.shape})"


def is_symbolic(x) -> bool:
    """True if ``x`` (scalar or array) still depends on a batch leaf."""
    if isinstance(x, Sym):
        return bool(x.terms)
    if isinstance(x, SymArray):
        return any(isinstance(v, Sym) and v.terms for v in x._d.flat)
    return False


def concrete(x) -> float:
    """Float value of a number or a constant ``Sym``."""
    if isinstance(x, Sym):
        return float(x)
    return float(x)
