"""Entry point for hand-built circuits plus general quantum-info helpers.

Mirror of the reference's ``qml_essentials/jaqsi.py``: re-exports ``Script`` and
offers ``partial_trace`` (jaqsi.py:79-117), ``marginalize_probs`` (jaqsi.py:120-146)
and ``build_parity_observable`` (jaqsi.py:149-167).  The helpers are host-side
NumPy post-processing of results that already left the GPU (the fused on-device
variants live in ``backend``: ``purities``, ``overlap_fidelities``).
"""

from __future__ import annotations

from functools import reduce
from typing import List, Sequence, Tuple, Union

import numpy as np

from .operations import Hermitian, PauliZ  # noqa: F401
from .script import BatchAxis, Script  # noqa: F401


def Hamiltonian(matrix, wires: Union[int, List[int]] = 0, record: bool = False) -> Hermitian:
    """Static Hamiltonian as a :class:`Hermitian` (jaqsi.py:35-57)."""
    return Hermitian(matrix, wires=wires, record=record)


def partial_trace(rho, n_qubits: int, keep: Sequence[int]):
    """Trace out every qubit not in ``keep``; accepts ``(2^n, 2^n)`` or
    ``(B, 2^n, 2^n)``."""
    rho = np.asarray(rho)
    dim = 2**n_qubits
    single = rho.shape == (dim, dim)
    r = rho.reshape((-1,) + (2,) * (2 * n_qubits))
    keep = list(keep)
    gone = [q for q in range(n_qubits) if q not in keep]
    # pair ket axis q with bra axis q for every traced qubit in one einsum
    letters = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXY"
    ket = [letters[q] for q in range(n_qubits)]
    bra = [letters[n_qubits + q] for q in range(n_qubits)]
    for q in gone:
        bra[q] = ket[q]
    out = [ket[q] for q in sorted(keep)] + [bra[q] for q in sorted(keep)]
    res = np.einsum("Z" + "".join(ket) + "".join(bra) + "->Z" + "".join(out), r)
    d = 2 ** len(keep)
    res = res.reshape(-1, d, d)
    return res[0] if single else res


def marginalize_probs(probs, n_qubits: int, keep: Tuple[int]):
    """Marginal over the qubits in ``keep``; always returns ``(B, 2^k)`` like the
    reference (jaqsi.py:120-146)."""
    dim = 2**n_qubits
    p = np.asarray(probs).reshape((-1,) + (2,) * n_qubits)
    drop = tuple(1 + q for q in range(n_qubits) if q not in keep)
    return p.sum(axis=drop).reshape(p.shape[0], -1) if drop else p.reshape(-1, dim)


def build_parity_observable(qubit_group: List[int]) -> Hermitian:
    """Z (x) Z (x) ... on ``qubit_group`` tagged with its Pauli label."""
    mat = reduce(np.kron, [PauliZ._matrix] * len(qubit_group))
    obs = Hermitian(matrix=mat, wires=qubit_group, record=False)
    obs._pauli_label = "Z" * len(qubit_group)
    return obs
