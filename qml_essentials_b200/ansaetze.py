"""Ansaetze, blocks and encodings (host side; defines the tape the compiler eats).

API mirror of the reference's ``qml_essentials/ansaetze.py``: ``Circuit`` ABC
(ansaetze.py:13-155), ``DeclarativeCircuit``/``Block`` (ansaetze.py:158-371), the
23 named ansaetze under ``Ansaetze`` (ansaetze.py:374-756) and ``Encoding``
(ansaetze.py:759-1000).  Layer contents come from one table (``_LAYOUTS``) rather
than one class body per circuit.
"""

from __future__ import annotations

import logging
import warnings
from abc import ABC, abstractmethod
from typing import Any, Callable, List, Optional, Tuple, Union

import numpy as np

from .gates import Gates, PulseInformation
from .topologies import Topology

log = logging.getLogger(__name__)


class Circuit(ABC):
    """Abstract ansatz: subclasses implement ``n_params_per_layer``,
    ``get_control_indices`` and ``build``."""

    def __init__(self) -> None:
        pass

    @abstractmethod
    def n_params_per_layer(self, n_qubits: int) -> int:
        raise NotImplementedError("n_params_per_layer method is not implemented")

    def n_pulse_params_per_layer(self, n_qubits: int) -> int:
        raise NotImplementedError("n_pulse_params_per_layer method is not implemented")

    @abstractmethod
    def get_control_indices(self, n_qubits: int) -> Optional[List[int]]:
        raise NotImplementedError("get_control_indices method is not implemented")

    def get_control_angles(self, w, n_qubits: int):
        """Parameters of the controlled rotations of one layer (ansaetze.py:75-94)."""
        indices = self.get_control_indices(n_qubits)
        if indices is None:
            return np.array([])
        if len(indices) == 3 and None in indices:
            return w[indices[0] : indices[1] : indices[2]]
        return w.take(np.array(indices))

    def _build(self, w, n_qubits: int, **kwargs: Any) -> Any:
        if kwargs.get("gate_mode", "unitary") == "pulse":
            raise NotImplementedError("pulse-level gates are outside the B200 backend scope")
        return self.build(w, n_qubits, **kwargs)

    @abstractmethod
    def build(self, w, n_qubits: int, **kwargs: Any) -> Any:
        raise NotImplementedError("build method is not implemented")

    def __call__(self, *args: Any, **kwds: Any) -> Any:
        self._build(*args, **kwds)


class Block:
    """One gate type applied across all qubits or across a pair topology."""

    def __init__(self, gate, topology: Any = None, **kwargs):
        self.gate = getattr(Gates, gate) if isinstance(gate, str) else gate
        if self.is_entangling:
            assert topology is not None, "Topology must be specified for entangling gates"
        self.topology = topology
        self.kwargs = kwargs

    def __repr__(self):
        if self.topology is None:
            return f"{self.__class__.__name__}({self.gate.__name__})"
        return f"{self.__class__.__name__}({self.topology.__name__}[{self.gate.__name__}])"

    @property
    def is_entangling(self):
        return Gates.is_entangling(self.gate)

    @property
    def is_rotational(self):
        return Gates.is_rotational(self.gate)

    @property
    def is_controlled_rotation(self):
        return self.is_entangling and self.is_rotational

    def enough_qubits(self, n_qubits) -> bool:
        if not self.is_entangling:
            return n_qubits >= 1
        span = self.kwargs.get("span", 1)
        span = span(n_qubits) if callable(span) else span
        return n_qubits >= 2 and n_qubits > span

    def _pairs(self, n_qubits):
        return self.topology(n_qubits=n_qubits, **self.kwargs)

    def _warn_skip(self, n_qubits):
        warnings.warn(
            f"Skipping {self.topology.__name__} with n_qubits={n_qubits} "
            "as there are not enough qubits for this topology."
        )

    def n_params(self, n_qubits: int) -> int:
        assert n_qubits > 0, "Number of qubits must be positive"
        if not self.is_rotational:
            return 0
        if self.is_entangling:
            if not self.enough_qubits(n_qubits):
                self._warn_skip(n_qubits)
                return 0
            return len(self._pairs(n_qubits))
        return 3 * n_qubits if self.gate.__name__ == "Rot" else n_qubits

    def n_pulse_params(self, n_qubits: int) -> int:
        assert n_qubits > 0, "Number of qubits must be positive"
        per_gate = PulseInformation.num_params(self.gate)
        if self.is_entangling:
            if not self.enough_qubits(n_qubits):
                self._warn_skip(n_qubits)
                return 0
            return per_gate * len(self._pairs(n_qubits))
        return per_gate * n_qubits

    def apply(self, n_qubits: int, w=None, w_idx: int = None, **kwargs) -> int:
        """Emit the block's gates; returns the advanced weight index
        (ansaetze.py:323-371)."""
        assert n_qubits > 0, "Number of qubits must be positive"
        if self.is_entangling:
            if not self.enough_qubits(n_qubits):
                self._warn_skip(n_qubits)
                return w_idx
            targets = self._pairs(n_qubits)
        else:
            targets = range(n_qubits)
        triple = self.gate.__name__ == "Rot"
        for wires in targets:
            if not self.is_rotational:
                self.gate(wires=wires, **kwargs)
                continue
            assert w is not None, "w must be provided for rotational gates"
            assert w_idx is not None, "w_idx must be provided for rotational gates"
            if triple:
                self.gate(w[w_idx], w[w_idx + 1], w[w_idx + 2], wires=wires, **kwargs)
                w_idx += 3
            else:
                self.gate(w[w_idx], wires=wires, **kwargs)
                w_idx += 1
        return w_idx


class DeclarativeCircuit(Circuit):
    """Circuit defined by a tuple of :class:`Block` (``structure()``)."""

    @classmethod
    def structure(cls) -> Tuple[Any, ...]:
        raise NotImplementedError

    @classmethod
    def n_params_per_layer(cls, n_qubits: int) -> int:
        return sum(b.n_params(n_qubits) for b in cls.structure())

    @classmethod
    def n_pulse_params_per_layer(cls, n_qubits: int) -> int:
        return sum(b.n_pulse_params(n_qubits) for b in cls.structure())

    @classmethod
    def get_control_indices(cls, n_qubits: int) -> Optional[List]:
        """Slice ``[start, None, None]`` when the controlled-rotation parameters
        form the tail of the layer, else explicit indices (ansaetze.py:181-213)."""
        counts = [(b.is_controlled_rotation, b.n_params(n_qubits)) for b in cls.structure()]
        total = sum(n for _, n in counts)
        picked, pos = [], 0
        for is_ctrl, n in counts:
            if is_ctrl:
                picked.extend(range(pos, pos + n))
            pos += n
        if not picked:
            return None
        if picked == list(range(total - len(picked), total)):
            return [-len(picked), None, None]
        return picked

    @classmethod
    def build(cls, w, n_qubits: int, **kwargs: Any) -> None:
        idx = 0
        for block in cls.structure():
            idx = block.apply(n_qubits, w, idx, **kwargs)
            Gates.Barrier(wires=list(range(n_qubits)), **kwargs)


_S, _B, _A = Topology.stairs, Topology.bricks, Topology.all_to_all
_LADDER_UP = dict(wrap=True, reverse=True, mirror=False)
_SKIP3 = dict(reverse=False, mirror=False, offset=lambda n: n - 1, span=3, wrap=True)

# name -> tuple of (gate, topology, kwargs); the contents of ansaetze.py:410-756
_LAYOUTS = {
    "No_Ansatz": (),
    "Circuit_1": (("RX",), ("RZ",)),
    "Circuit_2": (("RX",), ("RZ",), ("CX", _S, {})),
    "Circuit_3": (("RX",), ("RZ",), ("CRZ", _S, {})),
    "Circuit_4": (("RX",), ("RZ",), ("CRX", _S, {})),
    "Circuit_5": (("RX",), ("RZ",), ("CRZ", _A, {}), ("RX",), ("RZ",)),
    "Circuit_6": (("RX",), ("RZ",), ("CRX", _A, {}), ("RX",), ("RZ",)),
    "Circuit_7": (("RX",), ("RZ",), ("CRZ", _B, {}), ("RX",), ("RZ",),
                  ("CRZ", _B, dict(offset=1))),
    "Circuit_8": (("RX",), ("RZ",), ("CRX", _B, {}), ("RX",), ("RZ",),
                  ("CRX", _B, dict(offset=1))),
    "Circuit_9": (("H",), ("CZ", _S, {}), ("RX",)),
    "Circuit_10": (("RY",), ("CZ", _S, dict(offset=-1, wrap=True)), ("RY",)),
    "Circuit_13": (("RY",), ("CRZ", _S, _LADDER_UP), ("RY",), ("CRZ", _S, _SKIP3)),
    "Circuit_14": (("RY",), ("CRX", _S, _LADDER_UP), ("RY",), ("CRX", _S, _SKIP3)),
    "Circuit_15": (("RY",), ("CX", _S, _LADDER_UP), ("RY",), ("CX", _S, _SKIP3)),
    "Circuit_16": (("RX",), ("RZ",), ("CRZ", _B, {}), ("CRZ", _B, dict(offset=1))),
    "Circuit_17": (("RX",), ("RZ",), ("CRX", _B, {}), ("CRX", _B, dict(offset=1))),
    "Circuit_18": (("RX",), ("RZ",), ("CRZ", _S, dict(wrap=True, mirror=False))),
    "Circuit_19": (("RX",), ("RZ",), ("CRX", _S, dict(wrap=True, mirror=False))),
    "Circuit_20": (("RY",), ("CX", _S, _LADDER_UP), ("RY",),
                   ("CX", _S, dict(reverse=False, offset=lambda n: n - 2, span=1,
                                   wrap=True))),
    "No_Entangling": (("Rot",),),
    "Hardware_Efficient": (("RY",), ("RZ",), ("RY",), ("CX", _B, dict(mirror=False)),
                           ("CX", _B, dict(offset=-1, modulo=True, wrap=True,
                                           mirror=False))),
    "Strongly_Entangling": (("Rot",),
                            ("CX", _S, dict(wrap=True, reverse=False, mirror=False)),
                            ("Rot",),
                            ("CX", _S, dict(reverse=False, span=lambda n: n // 2,
                                            wrap=True, mirror=False))),
}


def _declare(name: str, layout) -> type:
    def structure(cls):
        return tuple(
            Block(gate=item[0], topology=item[1] if len(item) > 1 else None,
                  **(item[2] if len(item) > 2 else {}))
            for item in layout
        )

    return type(name, (DeclarativeCircuit,), {"structure": classmethod(structure)})


class Ansaetze:
    """Namespace of the built-in ansaetze (ansaetze.py:374-756)."""

    @staticmethod
    def get_available(parameterized_only: bool = False):
        names = [n for n in _LAYOUTS if n != "No_Ansatz"]
        # the reference lists the parameterised circuits in this order
        order = [f"Circuit_{i}" for i in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 13, 14, 15, 16,
                                          17, 18, 19, 20)]
        order += ["No_Entangling", "Strongly_Entangling", "Hardware_Efficient"]
        assert set(order) == set(names)
        out = [getattr(Ansaetze, n) for n in order]
        if not parameterized_only:
            out += [Ansaetze.No_Ansatz, Ansaetze.GHZ]
        return out


for _name, _layout in _LAYOUTS.items():
    setattr(Ansaetze, _name, _declare(_name, _layout))


class _GHZ(DeclarativeCircuit):
    """H on wire 0 then a CX chain; no parameters (ansaetze.py:415-433)."""

    @classmethod
    def structure(cls):
        return (Block(gate=Gates.H), Block(gate=Gates.CX, topology=_S, reverse=True))

    @classmethod
    def build(cls, w, n_qubits: int, **kwargs):
        Gates.H(wires=0, **kwargs)
        for q in range(n_qubits - 1):
            Gates.CX(wires=[q, q + 1], **kwargs)

    @classmethod
    def n_pulse_params_per_layer(cls, n_qubits: int) -> int:
        return PulseInformation.num_params("H") + (n_qubits - 1) * PulseInformation.num_params(
            Gates.CX
        )


_GHZ.__name__ = _GHZ.__qualname__ = "GHZ"
Ansaetze.GHZ = _GHZ


class Encoding:
    """Input-encoding strategy: hamming | binary | ternary | golomb
    (ansaetze.py:759-1000, after doi:10.22331/q-2023-12-20-1210)."""

    _STRATEGIES = ("hamming", "binary", "ternary", "golomb")

    def __init__(self, strategy: str, gates: Union[str, Callable, List[Union[str, Callable]]]):
        if strategy not in self._STRATEGIES:
            raise ValueError(
                f"Encoding strategy {strategy} not implemented. "
                "Available options: ['hamming', 'binary', 'ternary', 'golomb']"
            )
        self._strategy = strategy
        wrap = getattr(self, strategy)
        if strategy == "golomb":
            self._gates = []
            self.callable = [wrap(None)]
        else:
            try:
                self._gates = Gates.parse_gates(gates, Gates)
            except ValueError as e:
                raise ValueError(f"Error parsing encodings: {e}")
            self.callable = [wrap(g) for g in self._gates]

    def __len__(self):
        return len(self.callable)

    def __getitem__(self, idx):
        return self.callable[idx]

    @property
    def is_golomb(self) -> bool:
        return self._strategy == "golomb"

    def _golomb_max(self) -> int:
        from .unitary import golomb_ruler

        n_qubits = getattr(self, "_n_qubits", None)
        if n_qubits is None:
            raise ValueError("Golomb encoding requires n_qubits to be set")
        return max(golomb_ruler(2**n_qubits))

    def get_n_freqs(self, omegas) -> int:
        """Number of frequencies (both signs + zero) after ``omegas`` encodings."""
        if self._strategy == "hamming":
            return int(2 * omegas + 1)
        if self._strategy == "binary":
            return int(2 ** (omegas + 1) - 1)
        if self._strategy == "ternary":
            return int(3**omegas)
        return int(2 * omegas * self._golomb_max() + 1)

    def get_spectrum(self, omegas):
        """Integer frequency support after ``omegas`` encodings."""
        if self._strategy == "hamming":
            top = omegas
        elif self._strategy == "binary":
            top = 2**omegas - 1
        elif self._strategy == "ternary":
            top = int(np.floor(3**omegas / 2))
        else:
            top = omegas * self._golomb_max()
        return np.arange(-top, top + 1)

    # -- strategies: wrap a per-qubit gate --------------------------------------
    def hamming(self, enc):
        return enc

    def binary(self, enc):
        def _enc(inputs, wires, **kwargs):
            return enc(inputs * (2**wires), wires, **kwargs)

        return _enc

    def ternary(self, enc):
        def _enc(inputs, wires, **kwargs):
            return enc(inputs * (3**wires), wires, **kwargs)

        return _enc

    def golomb(self, enc):
        def _enc(inputs, wires, **kwargs):
            Gates.GolombEncoding(w=inputs, wires=wires, **kwargs)

        return _enc
