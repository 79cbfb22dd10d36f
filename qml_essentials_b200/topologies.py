"""[control, target] wire-pair generators for two-qubit blocks.

Behavioural mirror of the reference's ``qml_essentials/topologies.py``
(topologies.py:22-121); pinned against the reference's own outputs by
``tests/golden/topologies.json``.
"""

from __future__ import annotations

import logging
from typing import Callable, List, Tuple, Union

log = logging.getLogger(__name__)

IntOrFn = Union[int, Callable[[int], int]]


def _resolve(v: IntOrFn, n_qubits: int) -> int:
    return v(n_qubits) if callable(v) else v


class Topology:
    """Static generators of ``(control, target)`` pairs."""

    @classmethod
    def stairs(
        cls,
        n_qubits: int,
        offset: IntOrFn = 0,
        wrap: bool = False,
        reverse: bool = True,
        mirror: bool = True,
        span: IntOrFn = 1,
        stride: int = 1,
        modulo: bool = True,
    ) -> List[Tuple[int, int]]:
        """Nearest-neighbour / spanned ladder of pairs (topologies.py:22-100).

        For every start ``q`` in ``range(0, n_qubits if wrap else n_qubits - 1,
        stride)`` the pair is ``(q + offset, q + offset + span)``; out-of-range
        ends either wrap around (``modulo``) or drop the pair.  ``reverse``
        flips the emission order, ``mirror`` swaps control and target.
        """
        shift = _resolve(offset, n_qubits)
        reach = _resolve(span, n_qubits)
        count = n_qubits if wrap else n_qubits - 1

        pairs: List[Tuple[int, int]] = []
        for start in range(0, count, stride):
            ctrl = start + shift
            tgt = ctrl + reach
            out_of_range = tgt >= n_qubits or ctrl < 0
            if out_of_range and not modulo:
                continue
            ctrl %= n_qubits
            tgt %= n_qubits
            if ctrl == tgt:
                log.warning("Skipping gate where control == target")
                continue
            pairs.append((ctrl, tgt))

        if reverse:
            pairs.reverse()
        if mirror:
            pairs = [(t, c) for (c, t) in pairs]
        return pairs

    @classmethod
    def bricks(cls, n_qubits: int, **kwargs) -> List[Tuple[int, int]]:
        """Every second rung of the ladder, no wrap-around (topologies.py:102-106)."""
        opts = {"stride": 2, "modulo": False}
        opts.update(kwargs)
        return cls.stairs(n_qubits=n_qubits, **opts)

    @classmethod
    def all_to_all(cls, n_qubits: int) -> List[List[int]]:
        """Every ordered pair, highest wires first (topologies.py:108-121)."""
        top = n_qubits - 1
        return [
            [top - a, (top - b) % n_qubits]
            for a in range(n_qubits)
            for b in range(n_qubits)
            if a != b
        ]
