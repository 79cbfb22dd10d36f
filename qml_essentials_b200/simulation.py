"""Tape-level helpers of the simulator seam (host side).

The reference's ``qml_essentials/simulation.py`` holds the einsum kernels; here
only the routing predicates live on the host (simulation.py:25-57).  The
evolution and measurement themselves (simulation.py:65-377) are the CUDA
library's job - see ``csrc/`` and ``include/qmlb200.h``.
"""

from __future__ import annotations

from typing import List, Optional

from .operations import KrausChannel, Operation


def infer_n_qubits(ops: List[Operation], obs: List[Operation]) -> int:
    """``max(wire) + 1`` over operations and observables, at least 1
    (simulation.py:25-39)."""
    wires = {w for o in list(ops) + list(obs) for w in o.wires}
    return max(wires) + 1 if wires else 1


def has_noise(tape: List[Operation]) -> bool:
    return any(isinstance(o, KrausChannel) for o in tape)


def uses_density(tape: List[Operation], type: str) -> bool:
    """Density-matrix simulation iff requested or a channel is on the tape
    (simulation.py:42-57)."""
    return type == "density" or has_noise(tape)


def simulate_and_measure(
    tape: List[Operation],
    n_qubits: int,
    type: str,
    obs: List[Operation],
    use_density: bool,
    shots: Optional[int] = None,
    key=None,
):
    """Run one concrete tape on the GPU and measure it (simulation.py:131-201)."""
    from .script import Script

    def replay():
        from .tape import active_tape

        active_tape().extend(tape)

    return Script(replay, n_qubits=n_qubits).execute(
        type=type, obs=obs, shots=shots, key=key
    )
