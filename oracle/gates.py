"""Gate matrices and Kraus sets, restated from the reference (oracle; test-only).

A tape entry is a tuple ``(name, wires, params, extra)``:

* ``name``   reference class name (``"RX"``, ``"CX"``, ``"DepolarizingChannel"`` ...)
* ``wires``  list of ints, ``wires[0]`` most significant (operations.py:38-50,439)
* ``params`` list of floats in the order of the reference's ``_param_names``
* ``extra``  matrix (``"QubitUnitary"``/``"Hermitian"``), list of Kraus matrices
             (``"QubitChannel"``), Pauli word (``"PauliRot"``,
             ``"ControlledPauliRot"`` -> ``(word, n_controls)``) or diagonal
             (``"DiagonalQubitUnitary"``)

All constants are exact doubles (SURVEY.md section 8(c) hazard (i): the
reference's ``H`` constant may carry float32 rounding when x64 is enabled after
import; the oracle deliberately does not reproduce that).
"""

from functools import reduce
from itertools import product

import numpy as np

C = np.complex128

I2 = np.eye(2, dtype=C)  # operations.py:727
X = np.array([[0, 1], [1, 0]], dtype=C)  # operations.py:749
Y = np.array([[0, -1j], [1j, 0]], dtype=C)  # operations.py:765
Z = np.array([[1, 0], [0, -1]], dtype=C)  # operations.py:781
HAD = np.array([[1, 1], [1, -1]], dtype=C) / np.sqrt(2.0)  # operations.py:797
S = np.array([[1, 0], [0, 1j]], dtype=C)  # operations.py:817
SWAP = np.array(
    [[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=C
)  # operations.py:833-835
P0 = np.array([[1, 0], [0, 0]], dtype=C)  # operations.py:1049
P1 = np.array([[0, 0], [0, 1]], dtype=C)  # operations.py:1050
PAULI = {"I": I2, "X": X, "Y": Y, "Z": Z}  # operations.py:994-999

KRAUS_NAMES = {
    "BitFlip",
    "PhaseFlip",
    "DepolarizingChannel",
    "AmplitudeDamping",
    "PhaseDamping",
    "ThermalRelaxationError",
    "QubitChannel",
}


def rot_pauli(theta, P):
    """R_P(theta) = cos(theta/2) I - i sin(theta/2) P (operations.py:1029-1031)."""
    dim = P.shape[0]
    return np.cos(theta / 2) * np.eye(dim, dtype=C) - 1j * np.sin(theta / 2) * P


def pauli_word(word):
    """Kronecker product of single-qubit Paulis (operations.py:1289-1290)."""
    return reduce(np.kron, [PAULI[c] for c in word])


def controlled(target):
    """|0><0| (x) I + |1><1| (x) target (operations.py:1074)."""
    return np.kron(P0, I2) + np.kron(P1, target)


def _perm3(swap_rows):
    m = np.eye(8, dtype=C)
    a, b = swap_rows
    m[[a, b]] = m[[b, a]]
    return m


CCX = _perm3((6, 7))  # operations.py:1112-1124
CSWAP = _perm3((5, 6))  # operations.py:1146-1158


def controlled_pauli_rot(theta, word, n_controls):
    """Identity with R_word(theta) in the last block (operations.py:1397-1411)."""
    R = rot_pauli(theta, pauli_word(word))
    d_t = R.shape[0]
    d_c = 2**n_controls
    mat = np.eye(d_c * d_t, dtype=C)
    start = (d_c - 1) * d_t
    mat[start:, start:] = R
    return mat


def unitary_matrix(name, wires, params, extra=None):
    """Matrix of a non-channel tape entry, indexed (out..., in...) with wires[0] MSB."""
    k = len(wires)
    if name in ("Id", "I"):
        return np.eye(2**k, dtype=C)  # operations.py:727,739-743
    if name == "PauliX":
        return X
    if name == "PauliY":
        return Y
    if name == "PauliZ":
        return Z
    if name == "H":
        return HAD
    if name == "S":
        return S
    if name == "SWAP":
        return SWAP
    if name == "RX":
        return rot_pauli(params[0], X)  # operations.py:1043
    if name == "RY":
        return rot_pauli(params[0], Y)  # operations.py:1044
    if name == "RZ":
        return rot_pauli(params[0], Z)  # operations.py:1045
    if name == "CX":
        return controlled(X)  # operations.py:1098
    if name == "CY":
        return controlled(Y)  # operations.py:1099
    if name == "CZ":
        return controlled(Z)  # operations.py:1100
    if name == "CCX":
        return CCX
    if name == "CSWAP":
        return CSWAP
    if name == "ControlledPhaseShift":
        phi = params[0]  # operations.py:1199-1200
        return np.kron(P0, I2) + np.kron(
            P1, np.array([[1, 0], [0, np.exp(1j * phi)]], dtype=C)
        )
    if name == "Rot":
        phi, theta, omega = params  # operations.py:1235-1242
        return rot_pauli(omega, Z) @ rot_pauli(theta, Y) @ rot_pauli(phi, Z)
    if name == "PauliRot":
        return rot_pauli(params[0], pauli_word(extra))  # operations.py:1289-1295
    if name in ("RXX", "RYY", "RZZ", "RZX"):
        return rot_pauli(params[0], pauli_word(name[1:]))  # operations.py:1348-1351
    if name == "ControlledPauliRot":
        word, n_controls = extra
        return controlled_pauli_rot(params[0], word, n_controls)
    if name in ("CRX", "CRY", "CRZ"):
        return controlled_pauli_rot(params[0], name[2], 1)  # operations.py:1485-1487
    if name in ("DiagonalQubitUnitary", "DiagU"):
        return np.diag(np.asarray(extra, dtype=C))  # operations.py:917
    if name in ("QubitUnitary", "Hermitian", "Operation"):
        return np.asarray(extra, dtype=C)
    raise ValueError(f"oracle: unknown gate {name!r}")


def thermal_relaxation_kraus(pe, t1, t2, tg):
    """operations.py:1854-1895 (both regimes)."""
    eT1 = np.exp(-tg / t1)
    p_reset = 1.0 - eT1
    eT2 = np.exp(-tg / t2)
    if t2 <= t1:
        pz = (1.0 - p_reset) * (1.0 - eT2 / eT1) / 2.0
        pr0 = (1.0 - pe) * p_reset
        pr1 = pe * p_reset
        pid = 1.0 - pz - pr0 - pr1
        return [
            np.sqrt(pid) * I2,
            np.sqrt(pz) * Z,
            np.sqrt(pr0) * np.array([[1, 0], [0, 0]], dtype=C),
            np.sqrt(pr0) * np.array([[0, 1], [0, 0]], dtype=C),
            np.sqrt(pr1) * np.array([[0, 0], [1, 0]], dtype=C),
            np.sqrt(pr1) * np.array([[0, 0], [0, 1]], dtype=C),
        ]
    choi = np.array(
        [
            [1 - pe * p_reset, 0, 0, eT2],
            [0, pe * p_reset, 0, 0],
            [0, 0, (1 - pe) * p_reset, 0],
            [eT2, 0, 0, 1 - (1 - pe) * p_reset],
        ],
        dtype=C,
    )
    lam, vec = np.linalg.eigh(choi)
    return [
        (np.sqrt(np.abs(lam[i])) * vec[:, i].reshape(2, 2, order="F")).astype(C)
        for i in range(4)
    ]


def n_qubit_depolarizing_kraus(p, n):
    """unitary.py:114-146: sqrt(1-p(4^n-1)/4^n) I, then sqrt(p/4^n) P for P != I."""
    paulis = [I2, X, Y, Z]
    dim = 2**n
    ops = [np.sqrt(1 - p * (4**n - 1) / (4**n)) * np.eye(dim, dtype=C)]
    for i, idx in enumerate(product(range(4), repeat=n)):
        if i == 0:
            continue
        ops.append(np.sqrt(p / (4**n)) * reduce(np.kron, [paulis[j] for j in idx]))
    return ops


def kraus_matrices(name, params, extra=None):
    """Kraus operators of a channel tape entry (operations.py:1581-1929)."""
    if name == "BitFlip":
        p = params[0]  # operations.py:1614-1617
        return [np.sqrt(1 - p) * I2, np.sqrt(p) * X]
    if name == "PhaseFlip":
        p = params[0]  # operations.py:1653-1656
        return [np.sqrt(1 - p) * I2, np.sqrt(p) * Z]
    if name == "DepolarizingChannel":
        p = params[0]  # operations.py:1693-1698
        return [
            np.sqrt(1 - p) * I2,
            np.sqrt(p / 3) * X,
            np.sqrt(p / 3) * Y,
            np.sqrt(p / 3) * Z,
        ]
    if name == "AmplitudeDamping":
        g = params[0]  # operations.py:1736-1739
        return [
            np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=C),
            np.array([[0, np.sqrt(g)], [0, 0]], dtype=C),
        ]
    if name == "PhaseDamping":
        g = params[0]  # operations.py:1776-1779
        return [
            np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=C),
            np.array([[0, 0], [0, np.sqrt(g)]], dtype=C),
        ]
    if name == "ThermalRelaxationError":
        return thermal_relaxation_kraus(*params)
    if name == "QubitChannel":
        return [np.asarray(K, dtype=C) for K in extra]
    raise ValueError(f"oracle: unknown channel {name!r}")


def is_channel(name):
    return name in KRAUS_NAMES
