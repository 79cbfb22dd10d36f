"""Unitary gate wrappers with noise insertion (host side).

Mirror of the reference's ``qml_essentials/unitary.py``: every wrapper applies
the optional coherent ``GateError`` jitter to its angle(s) (unitary.py:200-246),
records the gate, then records the incoherent channels on every wire it touched
(unitary.py:150-197).  Angles may be affine proxies of batched arguments.
"""

from __future__ import annotations

import itertools
import logging
from functools import reduce
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

from . import operations as op
from . import rng
from .rng import safe_random_split
from .symbolic import Sym, SymArray

log = logging.getLogger(__name__)

_GOLOMB_RULER_CACHE: Dict[int, Tuple[int, ...]] = {}


def _greedy_golomb(d: int) -> Tuple[int, ...]:
    """Greedy Golomb ruler of order ``d`` (unitary.py:18-50): each new mark is the
    smallest integer whose differences to all earlier marks are unused."""
    if d <= 0:
        return ()
    marks: List[int] = [0]
    used = set()
    nxt = 1
    while len(marks) < d:
        cand = {nxt - m for m in marks}
        if used.isdisjoint(cand):
            marks.append(nxt)
            used |= cand
        nxt += 1
    return tuple(marks)


def golomb_ruler(d: int) -> Tuple[int, ...]:
    """Cached Golomb ruler of order ``d`` (unitary.py:53-84)."""
    if d <= 0:
        raise ValueError(f"Golomb ruler order must be positive, got {d}")
    if d not in _GOLOMB_RULER_CACHE:
        _GOLOMB_RULER_CACHE[d] = _greedy_golomb(d)
    return _GOLOMB_RULER_CACHE[d]


def _n_qubit_depolarizing_kraus(p: float, n: int) -> List[np.ndarray]:
    """unitary.py:114-146."""
    if not (0.0 <= p <= 1.0):
        raise ValueError(f"Probability p must be between 0 and 1, got {p}")
    if n < 2:
        raise ValueError(f"Number of qubits must be >= 2, got {n}")
    paulis = op._PAULI_MATS
    dim, n_words = 2**n, 4**n
    out = [np.sqrt(1 - p * (n_words - 1) / n_words) * np.eye(dim)]
    words = itertools.product(range(4), repeat=n)
    next(words)  # identity word handled above
    scale = np.sqrt(p / n_words)
    for idx in words:
        out.append(scale * reduce(np.kron, [paulis[i] for i in idx]))
    return out


class UnitaryGates:
    """Collection of unitary gates with optional noise simulation."""

    batch_gate_error = True

    @staticmethod
    def NQubitDepolarizingChannel(p: float, wires: List[int]) -> op.QubitChannel:
        return op.QubitChannel(_n_qubit_depolarizing_kraus(p, len(wires)), wires=wires)

    @staticmethod
    def Noise(wires, noise_params: Optional[Dict[str, float]] = None) -> None:
        """Channels after a gate, per wire then per pair (unitary.py:175-197)."""
        if noise_params is None:
            return
        if isinstance(wires, (int, np.integer)):
            wires = [wires]
        per_wire = (
            ("BitFlip", op.BitFlip),
            ("PhaseFlip", op.PhaseFlip),
            ("Depolarizing", op.DepolarizingChannel),
        )
        for w in wires:
            for key, channel in per_wire:
                p = noise_params.get(key, 0.0)
                if p > 0:
                    channel(p, wires=w)
        if len(wires) > 1:
            p = noise_params.get("MultiQubitDepolarizing", 0.0)
            if p > 0:
                UnitaryGates.NQubitDepolarizingChannel(p, list(wires))

    @staticmethod
    def GateError(w, noise_params=None, random_key=None):
        """Gaussian angle jitter (unitary.py:226-246).  With ``batch_gate_error``
        every batch element draws its own value from its own key."""
        if noise_params is not None and noise_params.get("GateError", None) is not None:
            assert random_key is not None, (
                "A random_key must be provided when using GateError"
            )
            if UnitaryGates.batch_gate_error:
                random_key, sub_key = safe_random_split(random_key)
                shape = w.shape if isinstance(w, (np.ndarray, SymArray)) else ()
            else:
                sub_key = rng.key(0)  # same draw for every element
                shape = ()
            sigma = noise_params["GateError"]
            if sigma != 0.0:
                w = w + sigma * rng.normal(sub_key, shape)
        return w, random_key

    # -- gates ----------------------------------------------------------------
    @staticmethod
    def Rot(phi, theta, omega, wires, noise_params=None, random_key=None) -> None:
        if noise_params is not None and "GateError" in noise_params:
            phi, random_key = UnitaryGates.GateError(phi, noise_params, random_key)
            theta, random_key = UnitaryGates.GateError(theta, noise_params, random_key)
            omega, random_key = UnitaryGates.GateError(omega, noise_params, random_key)
        op.Rot(phi, theta, omega, wires=wires)
        UnitaryGates.Noise(wires, noise_params)

    @staticmethod
    def PauliRot(theta, pauli, wires, noise_params=None, random_key=None) -> None:
        if noise_params is not None and "GateError" in noise_params:
            theta, random_key = UnitaryGates.GateError(theta, noise_params, random_key)
        op.PauliRot(theta, pauli, wires=wires)
        UnitaryGates.Noise(wires, noise_params)

    @staticmethod
    def GolombEncoding(w, wires, noise_params=None, random_key=None) -> None:
        """S(x) = exp(-i diag(golomb marks) x) on all given wires (unitary.py:661-701)."""
        wl = list(wires) if isinstance(wires, (list, tuple, range)) else [wires]
        marks = np.array(golomb_ruler(2 ** len(wl)), dtype=float)
        w, random_key = UnitaryGates.GateError(w, noise_params, random_key)
        op.DiagonalQubitUnitary.from_phase(marks, w, wl)
        UnitaryGates.Noise(wl, noise_params)


def _angle_gate(op_name: str, doc_ref: str):
    cls = getattr(op, op_name)

    def gate(w, wires, noise_params=None, random_key=None) -> None:
        w, random_key = UnitaryGates.GateError(w, noise_params, random_key)
        cls(w, wires=wires)
        UnitaryGates.Noise(wires, noise_params)

    gate.__doc__ = f"{op_name} with optional GateError and channel noise ({doc_ref})."
    return staticmethod(gate)


def _fixed_gate(op_name: str, doc_ref: str):
    cls = getattr(op, op_name)

    def gate(wires, noise_params=None, random_key=None) -> None:
        cls(wires=wires)
        UnitaryGates.Noise(wires, noise_params)

    gate.__doc__ = f"{op_name} with optional channel noise ({doc_ref})."
    return staticmethod(gate)


for _n, _ref in (("RX", "unitary.py:313-333"), ("RY", "unitary.py:336-356"),
                 ("RZ", "unitary.py:359-379"), ("CRX", "unitary.py:382-402"),
                 ("CRY", "unitary.py:405-425"), ("CRZ", "unitary.py:428-448"),
                 ("RXX", "unitary.py:451-473"), ("RYY", "unitary.py:476-498"),
                 ("RZZ", "unitary.py:501-523"), ("RZX", "unitary.py:526-549")):
    setattr(UnitaryGates, _n, _angle_gate(_n, _ref))
setattr(UnitaryGates, "CPhase", _angle_gate("ControlledPhaseShift", "unitary.py:552-575"))
for _n, _ref in (("CX", "unitary.py:578-596"), ("CY", "unitary.py:599-617"),
                 ("CZ", "unitary.py:620-638"), ("H", "unitary.py:641-659")):
    setattr(UnitaryGates, _n, _fixed_gate(_n, _ref))
# single-qubit Paulis as state-preparation gates (convenience, same noise rule)
for _n in ("PauliX", "PauliY", "PauliZ", "S"):
    setattr(UnitaryGates, _n, _fixed_gate(_n, "operations.py:746-827"))
