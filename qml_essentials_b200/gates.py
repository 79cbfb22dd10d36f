"""``Gates.<NAME>(...)`` dispatcher (host side).

Mirror of the reference's ``qml_essentials/gates.py``: attribute access on the
``Gates`` class yields a handler named after the gate (gates.py:24-33,71-161);
the handler filters keyword arguments and forwards to the unitary gate set.
Pulse-level gates (pulses.py, an ODE solve per leaf gate) are outside the B200
hot-path scope: ``gate_mode="pulse"`` raises.
"""

from __future__ import annotations

import logging
from contextlib import contextmanager
from typing import Callable, List, Union

from .operations import Barrier as BarrierOp
from .unitary import UnitaryGates

log = logging.getLogger(__name__)

_ALLOWED = ("w", "wires", "phi", "theta", "omega", "noise_params", "random_key")
_ROTATIONAL = {"RX", "RY", "RZ", "Rot", "CRX", "CRY", "CRZ", "GolombEncoding", "CPhase"}
_ENTANGLING = {"CX", "CY", "CZ", "CRX", "CRY", "CRZ", "CPhase"}


class PulseInformation:
    """Placeholder for the reference's pulse-parameter registry (pulses.py:633+).
    The unitary path needs only 'no pulse parameters'."""

    @staticmethod
    def set_envelope(name: str) -> None:
        return None

    @staticmethod
    def gate_by_name(name):
        return None

    @staticmethod
    def num_params(gate) -> int:
        return 0


def Barrier(wires: Union[int, List[int]], *args, **kwargs):
    return BarrierOp(wires)


class GatesMeta(type):
    def __getattr__(cls, gate_name):
        if gate_name.startswith("__"):
            raise AttributeError(gate_name)

        def handler(*args, **kwargs):
            return cls._inner_getattr(gate_name, *args, **kwargs)

        handler.__name__ = gate_name
        return handler


class Gates(metaclass=GatesMeta):
    """Dynamic accessor: ``Gates.RX(w, wires, noise_params=..., random_key=...)``."""

    @classmethod
    def _inner_getattr(cls, gate_name, *args, **kwargs):
        if gate_name == "Barrier":
            return Barrier(*args, **kwargs)
        gate_mode = kwargs.pop("gate_mode", "unitary")
        if gate_mode == "pulse":
            raise NotImplementedError(
                "gate_mode='pulse' (pulse-level ODE gates) is outside the scope of the "
                "B200 circuit-execution backend"
            )
        if gate_mode != "unitary":
            raise ValueError(f"Unknown gate mode: {gate_mode}. Use 'unitary' or 'pulse'.")
        dropped = [k for k in kwargs if k not in _ALLOWED]
        if dropped:
            log.debug(f"Unsupported keyword arguments: {dropped}")
        kwargs = {k: v for k, v in kwargs.items() if k in _ALLOWED}
        gate = getattr(UnitaryGates, gate_name, None)
        if gate is None:
            raise AttributeError(f"'UnitaryGates' object has no attribute '{gate_name}'")
        return gate(*args, **kwargs)

    @classmethod
    @contextmanager
    def pulse_manager_context(cls, pulse_params):
        yield

    @classmethod
    def parse_gates(cls, gates, set_of_gates=None):
        """str | callable | list of both | None -> list of callables (gates.py:173-207)."""
        source = set_of_gates or cls
        if gates is None:
            return [lambda *a, **k: None]
        if isinstance(gates, str):
            return [getattr(source, gates)]
        if isinstance(gates, list):
            parsed = []
            for g in gates:
                if isinstance(g, str):
                    parsed.append(getattr(source, g))
                elif callable(g):
                    parsed.append(g)
                else:
                    raise ValueError(
                        f"Operation {g} is not a valid gate or callable. Got {type(g)}"
                    )
            return parsed
        if callable(gates):
            return [gates]
        raise ValueError(
            f"Operation {gates} is not a valid gate or callable or list of both."
        )

    @classmethod
    def is_rotational(cls, gate) -> bool:
        return gate.__name__ in _ROTATIONAL

    @classmethod
    def is_entangling(cls, gate) -> bool:
        return gate.__name__ in _ENTANGLING
