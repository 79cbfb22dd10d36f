"""Gate / channel / observable objects (host side, NumPy).

API mirror of the reference's ``qml_essentials/operations.py`` for the
circuit-execution path: instantiating an operation inside a circuit function
appends it to the active tape (operations.py:112-155); ``name``, ``wires``,
``parameters``, ``matrix``, ``dagger()/power()/*/@/+`` and the Kraus channels keep
their meaning.  What differs is the engine underneath: an operation here does
not contract itself against a state.  It describes itself to the tape-to-kernel
compiler through :meth:`Operation.spec` as one of a few *matrix sources*
(constant, trigonometric-affine, product, diagonal phase, Kraus set) whose
parameters may be affine proxies (:mod:`.symbolic`) of the batched arguments.

Conventions: wire 0 is the most significant bit; a k-qubit
matrix is indexed ``(out_0..out_{k-1}, in_0..in_{k-1})`` with ``wires[0]`` most
significant (operations.py:38-50,439).
"""

from __future__ import annotations

from functools import reduce
from itertools import product as _iproduct
from typing import Callable, List, Optional, Sequence, Tuple, Union

import numpy as np

from .symbolic import Sym, SymArray, SymbolicError, is_symbolic
from .tape import active_tape, recording  # noqa: F401  (re-export, as the reference)

C128 = np.complex128


def _cdtype():
    """Host matrices are always complex128; device precision is a Script option."""
    return C128


# ---------------------------------------------------------------------------
# matrix sources handed to the compiler
# ---------------------------------------------------------------------------
class ConstMat:
    """A fixed 2^k x 2^k matrix."""

    __slots__ = ("matrix",)

    def __init__(self, matrix):
        self.matrix = np.asarray(matrix, dtype=C128)

    def evaluate(self):
        return self.matrix

    def dagger(self):
        return ConstMat(self.matrix.conj().T)

    def scaled(self, x):
        return ConstMat(self.matrix * x)

    @property
    def symbolic(self):
        return False


class TrigMat:
    """M(theta) = C0 + cos(kappa*theta) * A + sin(kappa*theta) * B.

    Covers every parametrised gate of the reference: R_P(theta) = cos(theta/2) I
    - i sin(theta/2) P (operations.py:1029-1031,1289-1295), controlled Pauli
    rotations (operations.py:1400-1411) and CPhase (operations.py:1199-1200).
    """

    __slots__ = ("C0", "A", "B", "kappa", "theta")

    def __init__(self, C0, A, B, kappa, theta):
        self.C0 = np.asarray(C0, dtype=C128)
        self.A = np.asarray(A, dtype=C128)
        self.B = np.asarray(B, dtype=C128)
        self.kappa = float(kappa)
        self.theta = theta

    @property
    def symbolic(self):
        return is_symbolic(self.theta)

    def evaluate(self):
        t = float(self.theta) * self.kappa
        return self.C0 + np.cos(t) * self.A + np.sin(t) * self.B

    def dagger(self):
        return TrigMat(self.C0.conj().T, self.A.conj().T, self.B.conj().T, self.kappa,
                       self.theta)

    def scaled(self, x):
        return TrigMat(self.C0 * x, self.A * x, self.B * x, self.kappa, self.theta)


class ProductMat:
    """Product of same-wire sources; ``factors[0]`` is applied to the state first."""

    __slots__ = ("factors",)

    def __init__(self, factors):
        self.factors = list(factors)

    @property
    def symbolic(self):
        return any(f.symbolic for f in self.factors)

    def evaluate(self):
        m = self.factors[0].evaluate()
        for f in self.factors[1:]:
            m = f.evaluate() @ m
        return m

    def dagger(self):
        return ProductMat([f.dagger() for f in reversed(self.factors)])

    def scaled(self, x):
        return ProductMat([self.factors[0].scaled(x)] + self.factors[1:])


class DiagPhaseMat:
    """diag(exp(-i * marks_j * theta)): the Golomb encoding (unitary.py:690-700)."""

    __slots__ = ("marks", "theta")

    def __init__(self, marks, theta):
        self.marks = np.asarray(marks, dtype=np.float64)
        self.theta = theta

    @property
    def symbolic(self):
        return is_symbolic(self.theta)

    def evaluate(self):
        return np.diag(np.exp(-1j * self.marks * float(self.theta))).astype(C128)

    def dagger(self):
        return DiagPhaseMat(-self.marks, self.theta)

    def scaled(self, x):
        raise SymbolicError("scaling a diagonal phase gate needs concrete values")


class KrausSet:
    """Kraus operators of a noise channel (always concrete)."""

    __slots__ = ("ops",)

    def __init__(self, ops):
        self.ops = [np.asarray(K, dtype=C128) for K in ops]

    symbolic = False


# ---------------------------------------------------------------------------
# host-side contraction helper (operator algebra only - never used by Script)
# ---------------------------------------------------------------------------
def _apply_matrix(mat: np.ndarray, tensor: np.ndarray, axes: Sequence[int]) -> np.ndarray:
    """Contract a (2^k, 2^k) matrix against ``axes`` of a rank-N tensor of 2s and
    put the output indices back in place (semantics of operations.py:19-77)."""
    k = len(axes)
    gt = mat.reshape((2,) * (2 * k))
    moved = np.tensordot(gt, tensor, axes=(list(range(k, 2 * k)), list(axes)))
    return np.moveaxis(moved, list(range(k)), list(axes))


class Operation:
    """Base class of gates, channels and observables (operations.py:80-512)."""

    is_controlled = False
    is_clifford = False

    _matrix: Optional[np.ndarray] = None
    _num_wires: Optional[int] = None
    _param_names: Tuple[str, ...] = ()

    def __init__(
        self,
        wires: Union[int, List[int]] = 0,
        matrix: Optional[np.ndarray] = None,
        record: bool = True,
        name: Optional[str] = None,
        _source=None,
    ) -> None:
        self.name = name or self.__class__.__name__
        self.wires = wires
        if self._num_wires is not None and len(self.wires) != self._num_wires:
            raise ValueError(
                f"{self.name} expects {self._num_wires} wire(s), "
                f"got {len(self.wires)}: {self.wires}"
            )
        if len(self.wires) != len(set(self.wires)):
            raise ValueError(f"{self.name} received duplicate wires: {self.wires}")
        self._source = _source
        if matrix is not None:
            self._matrix = np.asarray(matrix, dtype=C128)
        if record:
            tape = active_tape()
            if tape is not None:
                tape.append(self)

    # -- description ---------------------------------------------------------
    @property
    def wires(self) -> List[int]:
        return self._wires

    @wires.setter
    def wires(self, wires) -> None:
        if isinstance(wires, (list, tuple, range)):
            self._wires = [int(w) for w in wires]
        else:
            self._wires = [int(wires)]

    @property
    def parameters(self) -> list:
        return [getattr(self, n) for n in self._param_names]

    def __repr__(self) -> str:
        params = self.parameters
        if params:
            txt = ", ".join(
                f"{float(v):.4f}" if not is_symbolic(v) else repr(v) for v in params
            )
            return f"{self.name}({txt}, wires={self.wires})"
        return f"{self.name}(wires={self.wires})"

    def spec(self):
        """Matrix source for the compiler."""
        if self._source is not None:
            return self._source
        if self._matrix is None:
            raise NotImplementedError(f"{self.__class__.__name__} does not define a matrix.")
        return ConstMat(self._matrix)

    @property
    def matrix(self) -> np.ndarray:
        """Concrete matrix (operations.py:189-203)."""
        return self.spec().evaluate()

    def decompose(self) -> List["Operation"]:
        raise NotImplementedError(
            f"{self.__class__.__name__} does not define a decomposition."
        )

    # -- tape-replacing algebra (operations.py:245-320) -----------------------
    def _update_tape_operation(self, op: "Operation") -> None:
        tape = active_tape()
        if tape is not None:
            if tape and tape[-1] is self:
                tape[-1] = op
            else:
                tape.append(op)

    def dagger(self) -> "Operation":
        op = Operation(wires=self.wires, record=False, _source=self.spec().dagger())
        self._update_tape_operation(op)
        return op

    def power(self, power) -> "Operation":
        mat = np.linalg.matrix_power(self.matrix, power)
        op = Operation(wires=self.wires, matrix=mat, record=False)
        self._update_tape_operation(op)
        return op

    def __mul__(self, other):
        if isinstance(other, Operation):
            return self.__matmul__(other)
        if is_symbolic(other):
            raise SymbolicError("scaling an operation by a batched value")
        op = Operation(wires=self.wires, record=False,
                       _source=self.spec().scaled(complex(other)))
        self._update_tape_operation(op)
        return op

    __rmul__ = __mul__

    def __add__(self, other: "Operation") -> "Operation":
        if sorted(self.wires) != sorted(other.wires):
            raise ValueError(
                "Can only add operations acting on the same set of wires, "
                f"got {self.wires} and {other.wires}"
            )
        return Operation(wires=self.wires, matrix=self.matrix + other.matrix, record=False)

    def prod(self, *ops: "Operation") -> "Operation":
        """Generalised product on the union of the wire sets (operations.py:344-384)."""
        if not ops:
            return self
        all_ops = (self,) + ops
        all_wires: List[int] = []
        for o in all_ops:
            for w in o.wires:
                if w not in all_wires:
                    all_wires.append(w)
        n = len(all_wires)
        mat = _embed_matrix(all_ops[0].matrix, all_ops[0].wires, all_wires, n)
        for o in all_ops[1:]:
            mat = mat @ _embed_matrix(o.matrix, o.wires, all_wires, n)
        return Operation(
            wires=all_wires,
            matrix=mat,
            name="Prod(" + "*".join(o.name for o in all_ops) + ")",
            record=False,
        )

    def __matmul__(self, other):
        if not isinstance(other, Operation):
            return NotImplemented
        return self.prod(other)

    # -- host operator algebra (NOT the execution path) ------------------------
    def lifted_matrix(self, n_qubits: int) -> np.ndarray:
        """Full 2^n x 2^n embedding (operations.py:402-419)."""
        dim = 2**n_qubits
        eye = np.eye(dim, dtype=C128).reshape((2,) * n_qubits + (dim,))
        out = _apply_matrix(self.matrix, eye, self.wires)
        return out.reshape(dim, dim)

    def apply_to_state(self, state: np.ndarray, n_qubits: int) -> np.ndarray:
        """Host helper for operator-algebra checks (operations.py:421-442).
        ``Script.execute`` never calls this - circuits run on the GPU only."""
        psi = np.asarray(state, dtype=C128).reshape((2,) * n_qubits)
        return _apply_matrix(self.matrix, psi, self.wires).reshape(2**n_qubits)

    def apply_to_density(self, rho: np.ndarray, n_qubits: int) -> np.ndarray:
        """Host helper, rho -> U rho U^dagger (operations.py:485-512)."""
        t = np.asarray(rho, dtype=C128).reshape((2,) * (2 * n_qubits))
        U = self.matrix
        t = _apply_matrix(U, t, self.wires)
        t = _apply_matrix(U.conj(), t, [w + n_qubits for w in self.wires])
        return t.reshape(2**n_qubits, 2**n_qubits)


class Hermitian(Operation):
    """Generic Hermitian observable / gate from a matrix (operations.py:515-578)."""

    def __init__(self, matrix, wires=0, record: bool = True) -> None:
        super().__init__(wires=wires, matrix=np.asarray(matrix, dtype=C128), record=record)

    def evolve(self, name=None, **kw):
        raise NotImplementedError(
            "Hamiltonian evolution (evolution.py) is outside the B200 hot-path scope"
        )


class QubitUnitary(Operation):
    """Arbitrary k-qubit matrix gate (generic ``Operation(matrix=...)`` shorthand)."""

    def __init__(self, matrix, wires=0, record: bool = True, name=None) -> None:
        super().__init__(wires=wires, matrix=matrix, record=record, name=name)


_I2 = np.eye(2, dtype=C128)
_X = np.array([[0, 1], [1, 0]], dtype=C128)
_Y = np.array([[0, -1j], [1j, 0]], dtype=C128)
_Z = np.array([[1, 0], [0, -1]], dtype=C128)
_P0 = np.array([[1, 0], [0, 0]], dtype=C128)
_P1 = np.array([[0, 0], [0, 1]], dtype=C128)


class Id(Operation):
    """Identity on any number of wires (operations.py:719-743)."""

    _matrix = _I2
    _num_wires = None
    is_clifford = True

    def __init__(self, wires=0, **kwargs) -> None:
        k = len(wires) if isinstance(wires, (list, tuple, range)) else 1
        if k > 1:
            kwargs["matrix"] = np.eye(2**k, dtype=C128)
        super().__init__(wires=wires, **kwargs)


def _fixed_gate(name, matrix, n_wires, clifford=True, controlled=False, doc=""):
    cls = type(
        name,
        (Operation,),
        {
            "_matrix": np.asarray(matrix, dtype=C128),
            "_num_wires": n_wires,
            "is_clifford": clifford,
            "is_controlled": controlled,
            "__doc__": doc,
            "__init__": lambda self, wires=list(range(n_wires)) if n_wires > 1 else 0,
            **kw: Operation.__init__(self, wires=wires, **kw),
        },
    )
    return cls


PauliX = _fixed_gate("PauliX", _X, 1, doc="Pauli-X (operations.py:746-759).")
PauliY = _fixed_gate("PauliY", _Y, 1, doc="Pauli-Y (operations.py:762-775).")
PauliZ = _fixed_gate("PauliZ", _Z, 1, doc="Pauli-Z (operations.py:778-791).")
H = _fixed_gate("H", np.array([[1, 1], [1, -1]]) / np.sqrt(2.0), 1,
                doc="Hadamard (operations.py:794-807); exact double constants.")
S = _fixed_gate("S", np.diag([1, 1j]), 1, doc="Phase gate (operations.py:810-827).")
SWAP = _fixed_gate(
    "SWAP", np.eye(4)[[0, 2, 1, 3]], 2, doc="SWAP (operations.py:830-845)."
)


def _controlled(target):
    return np.kron(_P0, _I2) + np.kron(_P1, target)


def _make_cz_decompose():
    def decompose(self):
        c, t = self.wires
        return [H(wires=t, record=False), CX(wires=[c, t], record=False),
                H(wires=t, record=False)]

    return decompose


CX = _fixed_gate("CX", _controlled(_X), 2, controlled=True,
                 doc="Controlled-X, wires=[control, target] (operations.py:1074,1098).")
CY = _fixed_gate("CY", _controlled(_Y), 2, controlled=True,
                 doc="Controlled-Y (operations.py:1099).")
CZ = _fixed_gate("CZ", _controlled(_Z), 2, controlled=True,
                 doc="Controlled-Z (operations.py:1100).")
CZ.decompose = _make_cz_decompose()
CCX = _fixed_gate("CCX", np.eye(8)[[0, 1, 2, 3, 4, 5, 7, 6]], 3, clifford=False,
                  controlled=True, doc="Toffoli (operations.py:1103-1134).")
CSWAP = _fixed_gate("CSWAP", np.eye(8)[[0, 1, 2, 3, 4, 6, 5, 7]], 3, clifford=False,
                    controlled=True, doc="Fredkin (operations.py:1137-1168).")

_PAULI_LABELS = ["I", "X", "Y", "Z"]
_PAULI_CLASSES = [Id, PauliX, PauliY, PauliZ]
_PAULI_MATRICES = {"I": _I2, "X": _X, "Y": _Y, "Z": _Z}
_PAULI_MATS = [_PAULI_MATRICES[c] for c in _PAULI_LABELS]


def _word_matrix(word: str) -> np.ndarray:
    return reduce(np.kron, [_PAULI_MATRICES[c] for c in word])


def _rot_source(P: np.ndarray, theta) -> TrigMat:
    """cos(theta/2) I - i sin(theta/2) P (operations.py:1029-1031)."""
    d = P.shape[0]
    return TrigMat(np.zeros((d, d)), np.eye(d), -1j * P, 0.5, theta)


def _make_rotation_gate(pauli_class, name):
    P = pauli_class._matrix

    class _RotationGate(Operation):
        __doc__ = f"{name}(theta) = exp(-i theta/2 {name[1]}) (operations.py:1002-1045)."
        _num_wires = 1
        _param_names = ("theta",)

        def __init__(self, theta, wires=0, **kwargs) -> None:
            self.theta = theta
            super().__init__(wires=wires, **kwargs)

        def spec(self):
            return _rot_source(P, self.theta)

        def generator(self):
            return pauli_class(wires=self.wires[0], record=False)

    _RotationGate.__name__ = _RotationGate.__qualname__ = name
    return _RotationGate


RX = _make_rotation_gate(PauliX, "RX")
RY = _make_rotation_gate(PauliY, "RY")
RZ = _make_rotation_gate(PauliZ, "RZ")


class ControlledPhaseShift(Operation):
    """CPhase(phi) = diag(1, 1, 1, e^{i phi}) (operations.py:1171-1201)."""

    _num_wires = 2
    _param_names = ("phi",)
    is_controlled = True

    def __init__(self, phi, wires=[0, 1], **kwargs) -> None:
        self.phi = phi
        super().__init__(wires=wires, **kwargs)

    def spec(self):
        e3 = np.zeros((4, 4), dtype=C128)
        e3[3, 3] = 1.0
        return TrigMat(np.diag([1, 1, 1, 0]), e3, 1j * e3, 1.0, self.phi)


class Rot(Operation):
    """Rot(phi, theta, omega) = RZ(omega) RY(theta) RZ(phi) (operations.py:1204-1252)."""

    _num_wires = 1
    _param_names = ("phi", "theta", "omega")

    def __init__(self, phi, theta, omega, wires=0, **kwargs) -> None:
        self.phi, self.theta, self.omega = phi, theta, omega
        super().__init__(wires=wires, **kwargs)

    def spec(self):
        return ProductMat(
            [_rot_source(_Z, self.phi), _rot_source(_Y, self.theta),
             _rot_source(_Z, self.omega)]
        )

    def decompose(self):
        w = self.wires[0]
        return [RZ(self.phi, wires=w, record=False), RY(self.theta, wires=w, record=False),
                RZ(self.omega, wires=w, record=False)]


class PauliRot(Operation):
    """exp(-i theta/2 P) for a Pauli word P (operations.py:1255-1312)."""

    _param_names = ("theta",)
    _PAULI_MAP = _PAULI_MATRICES

    def __init__(self, theta, pauli_word: str, wires=0, **kwargs) -> None:
        self.theta = theta
        self.pauli_word = pauli_word
        super().__init__(wires=wires, **kwargs)
        if len(self.wires) != len(pauli_word):
            raise ValueError(
                f"PauliRot word {pauli_word!r} needs {len(pauli_word)} wires, "
                f"got {self.wires}"
            )

    def spec(self):
        return _rot_source(_word_matrix(self.pauli_word), self.theta)

    def generator(self):
        return Hermitian(matrix=_word_matrix(self.pauli_word), wires=self.wires,
                         record=False)


def _make_pauli_rotation_subclass(name: str, word: str):
    class _PauliRotSubclass(PauliRot):
        __doc__ = f"{name}(theta) = exp(-i theta/2 {word}) (operations.py:1315-1351)."
        _num_wires = len(word)

        def __init__(self, theta, wires=None, **kwargs) -> None:
            if wires is None:
                wires = list(range(len(word)))
            super().__init__(theta, word, wires=wires, **kwargs)

    _PauliRotSubclass.__name__ = _PauliRotSubclass.__qualname__ = name
    return _PauliRotSubclass


RXX = _make_pauli_rotation_subclass("RXX", "XX")
RYY = _make_pauli_rotation_subclass("RYY", "YY")
RZZ = _make_pauli_rotation_subclass("RZZ", "ZZ")
RZX = _make_pauli_rotation_subclass("RZX", "ZX")


class ControlledPauliRot(Operation):
    """PauliRot on the targets conditioned on all controls being |1>
    (operations.py:1357-1427): identity with R in the last block."""

    _param_names = ("theta",)
    is_controlled = True

    def __init__(self, theta, pauli_word: str, wires, n_controls: int = 1, **kwargs):
        self.theta = theta
        self.pauli_word = pauli_word
        self.n_controls = n_controls
        wl = [wires] if isinstance(wires, (int, np.integer)) else list(wires)
        if len(wl) != n_controls + len(pauli_word):
            raise ValueError(
                f"ControlledPauliRot expects {n_controls + len(pauli_word)} wires "
                f"({n_controls} control + {len(pauli_word)} target), got {len(wl)}."
            )
        super().__init__(wires=wl, **kwargs)

    def _blocks(self):
        P = _word_matrix(self.pauli_word)
        d_t = P.shape[0]
        dim = (2**self.n_controls) * d_t
        start = dim - d_t
        return P, d_t, dim, start

    def spec(self):
        P, d_t, dim, start = self._blocks()
        C0 = np.eye(dim, dtype=C128)
        C0[start:, start:] = 0
        A = np.zeros((dim, dim), dtype=C128)
        A[start:, start:] = np.eye(d_t)
        B = np.zeros((dim, dim), dtype=C128)
        B[start:, start:] = -1j * P
        return TrigMat(C0, A, B, 0.5, self.theta)

    def generator(self):
        P, d_t, dim, start = self._blocks()
        gen = np.zeros((dim, dim), dtype=C128)
        gen[start:, start:] = P
        return Hermitian(matrix=gen, wires=self.wires, record=False)


def _make_controlled_rotation_subclass(name: str, axis: str):
    class _CRotation(ControlledPauliRot):
        __doc__ = (f"{name}(theta) = |0><0| (x) I + |1><1| (x) R{axis}(theta) "
                   "(operations.py:1430-1487).")
        _num_wires = 2

        def __init__(self, theta, wires=[0, 1], **kwargs) -> None:
            super().__init__(theta, axis, wires=wires, n_controls=1, **kwargs)

        def decompose(self):
            c, t = self.wires
            th = self.theta
            core = [RZ(th / 2, wires=t, record=False), CX(wires=[c, t], record=False),
                    RZ(-th / 2, wires=t, record=False), CX(wires=[c, t], record=False)]
            if axis == "Z":
                return core
            if axis == "X":
                return [H(wires=t, record=False)] + core + [H(wires=t, record=False)]
            return [RX(-np.pi / 2, wires=t, record=False)] + core[:3] + [
                RX(np.pi / 2, wires=t, record=False)]

    _CRotation.__name__ = _CRotation.__qualname__ = name
    return _CRotation


CRX = _make_controlled_rotation_subclass("CRX", "X")
CRY = _make_controlled_rotation_subclass("CRY", "Y")
CRZ = _make_controlled_rotation_subclass("CRZ", "Z")


class DiagonalQubitUnitary(Operation):
    """diag(d_0 .. d_{2^k-1}) (operations.py:881-961)."""

    _param_names = ()

    def __init__(self, diag, wires=0, **kwargs) -> None:
        wl = list(wires) if isinstance(wires, (list, tuple, range)) else [wires]
        self.diag = np.asarray(diag, dtype=C128)
        if self.diag.shape != (2 ** len(wl),):
            raise ValueError(
                f"DiagonalQubitUnitary expects {2 ** len(wl)} diagonal entries "
                f"for {len(wl)} wire(s), got shape {self.diag.shape}"
            )
        kwargs.setdefault("name", "DiagU")
        super().__init__(wires=wires, matrix=np.diag(self.diag), **kwargs)

    @classmethod
    def from_phase(cls, marks, theta, wires, **kwargs):
        """diag(exp(-i marks theta)) with ``theta`` possibly a batched proxy."""
        if not is_symbolic(theta):
            return cls(np.exp(-1j * np.asarray(marks, float) * float(theta)), wires, **kwargs)
        self = cls.__new__(cls)
        self.diag = None
        kwargs.setdefault("name", "DiagU")
        Operation.__init__(self, wires=wires, _source=DiagPhaseMat(marks, theta), **kwargs)
        return self


class Barrier(Operation):
    """No-op separator; still counted on the tape (operations.py:964-991)."""

    _matrix = None

    def __init__(self, wires=0) -> None:
        super().__init__(wires=wires)

    def spec(self):
        return None

    def apply_to_state(self, state, n_qubits):
        return state

    def apply_to_density(self, rho, n_qubits):
        return rho


class RandomUnitary(Operation):
    """Random Hermitian matrix as a gate (operations.py:848-878; parity of the
    random stream is unpinned, see rng.py)."""

    def __init__(self, wires, key, scale: float = 1.0, record: bool = True) -> None:
        from . import rng

        wl = list(wires) if isinstance(wires, (list, tuple, range)) else [wires]
        dim = 2 ** len(wl)
        ka, kb = rng.split(key)
        A = rng.normal(ka, (dim, dim)) + 1j * rng.normal(kb, (dim, dim))
        Hm = (A + A.conj().T) / 2.0
        Hm = Hm * (scale / np.linalg.norm(Hm, ord="fro"))
        super().__init__(wl, matrix=Hm, record=record)


# ---------------------------------------------------------------------------
# noise channels (operations.py:1490-1929)
# ---------------------------------------------------------------------------
class KrausChannel(Operation):
    """Base class: phi(rho) = sum_k K_k rho K_k^dagger."""

    def kraus_matrices(self) -> List[np.ndarray]:
        raise NotImplementedError

    def spec(self):
        return KrausSet(self.kraus_matrices())

    @property
    def matrix(self):
        raise TypeError(
            f"{self.__class__.__name__} is a noise channel and has no single "
            "unitary matrix. Use apply_to_density() instead."
        )

    def apply_to_state(self, state, n_qubits):
        raise TypeError(
            f"{self.__class__.__name__} is a noise channel and cannot be "
            "applied to a pure statevector. Use execute(type='density') instead."
        )

    def apply_to_density(self, rho, n_qubits):
        t = np.asarray(rho, dtype=C128).reshape((2,) * (2 * n_qubits))
        bra = [w + n_qubits for w in self.wires]
        acc = np.zeros_like(t)
        for K in self.kraus_matrices():
            acc = acc + _apply_matrix(K.conj(), _apply_matrix(K, t, self.wires), bra)
        return acc.reshape(2**n_qubits, 2**n_qubits)


def _check_prob(p, name="p"):
    if is_symbolic(p):
        raise SymbolicError("channel probabilities must not be batched")
    if not 0.0 <= p <= 1.0:
        raise ValueError(f"{name} must be in [0, 1].")


class BitFlip(KrausChannel):
    """sqrt(1-p) I, sqrt(p) X (operations.py:1581-1617)."""

    _num_wires = 1
    _param_names = ("p",)

    def __init__(self, p, wires=0) -> None:
        _check_prob(p)
        self.p = p
        super().__init__(wires=wires)

    def kraus_matrices(self):
        return [np.sqrt(1 - self.p) * _I2, np.sqrt(self.p) * _X]


class PhaseFlip(KrausChannel):
    """sqrt(1-p) I, sqrt(p) Z (operations.py:1620-1656)."""

    _num_wires = 1
    _param_names = ("p",)

    def __init__(self, p, wires=0) -> None:
        _check_prob(p)
        self.p = p
        super().__init__(wires=wires)

    def kraus_matrices(self):
        return [np.sqrt(1 - self.p) * _I2, np.sqrt(self.p) * _Z]


class DepolarizingChannel(KrausChannel):
    """sqrt(1-p) I, sqrt(p/3) X, Y, Z (operations.py:1659-1698)."""

    _num_wires = 1
    _param_names = ("p",)

    def __init__(self, p, wires=0) -> None:
        _check_prob(p)
        self.p = p
        super().__init__(wires=wires)

    def kraus_matrices(self):
        p = self.p
        return [np.sqrt(1 - p) * _I2, np.sqrt(p / 3) * _X, np.sqrt(p / 3) * _Y,
                np.sqrt(p / 3) * _Z]


class AmplitudeDamping(KrausChannel):
    """diag(1, sqrt(1-g)), sqrt(g)|0><1| (operations.py:1701-1739)."""

    _num_wires = 1
    _param_names = ("gamma",)

    def __init__(self, gamma, wires=0) -> None:
        _check_prob(gamma, "gamma")
        self.gamma = gamma
        super().__init__(wires=wires)

    def kraus_matrices(self):
        g = self.gamma
        return [np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=C128),
                np.array([[0, np.sqrt(g)], [0, 0]], dtype=C128)]


class PhaseDamping(KrausChannel):
    """diag(1, sqrt(1-g)), diag(0, sqrt(g)) (operations.py:1742-1779)."""

    _num_wires = 1
    _param_names = ("gamma",)

    def __init__(self, gamma, wires=0) -> None:
        _check_prob(gamma, "gamma")
        self.gamma = gamma
        super().__init__(wires=wires)

    def kraus_matrices(self):
        g = self.gamma
        return [np.array([[1, 0], [0, np.sqrt(1 - g)]], dtype=C128),
                np.array([[0, 0], [0, np.sqrt(g)]], dtype=C128)]


class ThermalRelaxationError(KrausChannel):
    """T1/T2 relaxation (operations.py:1782-1895), both regimes."""

    _num_wires = 1
    _param_names = ("pe", "t1", "t2", "tg")

    def __init__(self, pe, t1, t2, tg, wires=0) -> None:
        if not 0.0 <= pe <= 1.0:
            raise ValueError("pe must be in [0, 1].")
        if t1 <= 0:
            raise ValueError("t1 must be > 0.")
        if t2 <= 0:
            raise ValueError("t2 must be > 0.")
        if t2 > 2 * t1:
            raise ValueError("t2 must be <= 2·t1.")
        if tg < 0:
            raise ValueError("tg must be >= 0.")
        self.pe, self.t1, self.t2, self.tg = pe, t1, t2, tg
        super().__init__(wires=wires)

    def kraus_matrices(self):
        pe, t1, t2, tg = self.pe, self.t1, self.t2, self.tg
        e1 = np.exp(-tg / t1)
        p_reset = 1.0 - e1
        e2 = np.exp(-tg / t2)
        if t2 <= t1:
            pz = (1.0 - p_reset) * (1.0 - e2 / e1) / 2.0
            pr0 = (1.0 - pe) * p_reset
            pr1 = pe * p_reset
            pid = 1.0 - pz - pr0 - pr1
            E = lambda r, c: np.array(  # noqa: E731
                [[1.0 if (i, j) == (r, c) else 0.0 for j in range(2)] for i in range(2)],
                dtype=C128)
            return [np.sqrt(pid) * _I2, np.sqrt(pz) * _Z, np.sqrt(pr0) * E(0, 0),
                    np.sqrt(pr0) * E(0, 1), np.sqrt(pr1) * E(1, 0), np.sqrt(pr1) * E(1, 1)]
        choi = np.array(
            [[1 - pe * p_reset, 0, 0, e2], [0, pe * p_reset, 0, 0],
             [0, 0, (1 - pe) * p_reset, 0], [e2, 0, 0, 1 - (1 - pe) * p_reset]],
            dtype=C128)
        lam, vec = np.linalg.eigh(choi)
        return [np.sqrt(abs(lam[i])) * vec[:, i].reshape(2, 2, order="F") for i in range(4)]


class QubitChannel(KrausChannel):
    """Generic channel from user Kraus operators (operations.py:1898-1929)."""

    def __init__(self, kraus_ops, wires=0) -> None:
        self._kraus_ops = [np.asarray(K, dtype=C128) for K in kraus_ops]
        super().__init__(wires=wires)

    def kraus_matrices(self):
        return self._kraus_ops


# ---------------------------------------------------------------------------
# small matrix helpers (operations.py:1932-2164)
# ---------------------------------------------------------------------------
def _permute_matrix(mat: np.ndarray, perm: list, n_qubits: int) -> np.ndarray:
    dim = 2**n_qubits
    t = mat.reshape([2] * (2 * n_qubits))
    t = np.transpose(t, list(perm) + [p + n_qubits for p in perm])
    return t.reshape(dim, dim)


def _embed_matrix(mat, op_wires, all_wires, n_total) -> np.ndarray:
    if len(op_wires) == n_total and list(op_wires) == list(all_wires):
        return mat
    missing = [w for w in all_wires if w not in op_wires]
    full = mat
    for _ in missing:
        full = np.kron(full, _I2)
    current = list(op_wires) + missing
    if current != list(all_wires):
        full = _permute_matrix(full, [current.index(w) for w in all_wires], n_total)
    return full


def evolve_pauli_with_clifford(clifford, pauli, adjoint_left: bool = True):
    all_wires = sorted(set(clifford.wires) | set(pauli.wires))
    n = len(all_wires)
    Cm = _embed_matrix(clifford.matrix, clifford.wires, all_wires, n)
    Pm = _embed_matrix(pauli.matrix, pauli.wires, all_wires, n)
    res = Cm.conj().T @ Pm @ Cm if adjoint_left else Cm @ Pm @ Cm.conj().T
    return Hermitian(matrix=res, wires=all_wires, record=False)


def prod(*ops: Operation) -> Operation:
    if not ops:
        raise ValueError("At least one operation must be provided to prod().")
    return ops[0].prod(*ops[1:])
