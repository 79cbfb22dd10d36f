"""Process-wide defaults of the B200 backend.

The reference switches complex64/complex128 through the global
``jax_enable_x64`` flag (operations.py:12-16, memory.py:26-33).  Here the
precision is an explicit setting: ``set_precision("complex128")`` (default, the
mode every reference test runs in) or ``"complex64"``; a ``Script`` can override
it per instance.
"""

from __future__ import annotations

_PRECISIONS = ("complex64", "complex128")
_state = {"precision": "complex128"}


def set_precision(precision: str) -> None:
    if precision not in _PRECISIONS:
        raise ValueError(f"precision must be one of {_PRECISIONS}, got {precision!r}")
    _state["precision"] = precision


def get_precision() -> str:
    return _state["precision"]


def complex_itemsize(precision: str = None) -> int:
    return 16 if (precision or _state["precision"]) == "complex128" else 8
