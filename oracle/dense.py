"""Second, independent oracle: dense 2^n x 2^n operators by bit arithmetic.

No einsum and no tensor reshapes - every gate is lifted to the full Hilbert
space entry by entry, then states/densities are evolved with plain matrix
products.  Used only to cross-check ``oracle.sim`` for n <= ~7 (test-only).
Conventions as in SURVEY.md section 8: wire 0 is the most significant bit of the
basis index; a k-qubit matrix is indexed with ``wires[0]`` most significant.
"""

import numpy as np

from . import gates as G

C = np.complex128


def lift(U, wires, n_qubits):
    """Full-space matrix of the k-qubit matrix ``U`` acting on ``wires``."""
    k = len(wires)
    dim = 2**n_qubits
    U = np.asarray(U, dtype=C)
    full = np.zeros((dim, dim), dtype=C)
    shifts = [n_qubits - 1 - w for w in wires]
    mask = 0
    for s in shifts:
        mask |= 1 << s
    for col in range(dim):
        rest = col & ~mask
        sub_in = 0
        for j, s in enumerate(shifts):
            sub_in |= ((col >> s) & 1) << (k - 1 - j)
        for sub_out in range(2**k):
            amp = U[sub_out, sub_in]
            if amp == 0:
                continue
            row = rest
            for j, s in enumerate(shifts):
                row |= ((sub_out >> (k - 1 - j)) & 1) << s
            full[row, col] += amp
    return full


def run(tape, n_qubits, density=None):
    """Return the final statevector, or density matrix if channels are present
    (or ``density=True``)."""
    has_noise = any(G.is_channel(e[0]) for e in tape)
    if density is None:
        density = has_noise
    dim = 2**n_qubits
    psi = np.zeros(dim, dtype=C)
    psi[0] = 1
    rho = np.outer(psi, psi.conj()) if density else None
    for e in tape:
        name, wires = e[0], list(e[1])
        params = list(e[2]) if len(e) > 2 else []
        extra = e[3] if len(e) > 3 else None
        if name == "Barrier":
            continue
        if G.is_channel(name):
            Ks = [lift(K, wires, n_qubits) for K in G.kraus_matrices(name, params, extra)]
            rho = sum(K @ rho @ K.conj().T for K in Ks)
        else:
            U = lift(G.unitary_matrix(name, wires, params, extra), wires, n_qubits)
            if density:
                rho = U @ rho @ U.conj().T
            else:
                psi = U @ psi
    return rho if density else psi


def expval(state_or_rho, ob, n_qubits):
    name, wires = ob[0], list(ob[1])
    params = list(ob[2]) if len(ob) > 2 else []
    extra = ob[3] if len(ob) > 3 else None
    O = lift(G.unitary_matrix(name, wires, params, extra), wires, n_qubits)
    x = np.asarray(state_or_rho)
    if x.ndim == 1:
        return float(np.real(x.conj() @ O @ x))
    return float(np.real(np.trace(O @ x)))
