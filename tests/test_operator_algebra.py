"""Operator algebra of `operations.py` (dagger / power / scalar and operator products / sums
/ tensor products): the analytic expectations of the reference's own tests
(tests/test_jaqsi.py:1384-1617), restated against the host mirror."""

import numpy as np
import pytest

from qml_essentials_b200 import operations as op
from qml_essentials_b200.script import Script
from qml_essentials_b200.tape import recording

X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]])
Z = np.diag([1.0 + 0j, -1.0])
I2 = np.eye(2, dtype=complex)
CXM = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=complex)


def _ev(circuit, obs, n=None):
    return np.asarray(Script(circuit, n_qubits=n).execute(type="expval", obs=obs))


def test_dagger_and_power_in_circuits():
    def undo():
        op.RX(0.5, wires=0)
        op.RX(0.5, wires=0).dagger()

    assert np.allclose(_ev(undo, [op.PauliZ(0, record=False)]), 1)

    def squared():
        op.PauliX(wires=0).power(2)

    assert np.allclose(_ev(squared, [op.PauliZ(0, record=False)]), 1)


def test_scalar_multiplication_both_sides_and_on_the_tape():
    x = op.PauliX(wires=0, record=False)
    assert np.allclose((x * 2.0).matrix, 2 * X) and (x * 2.0).wires == [0]
    assert np.allclose((2.0 * x).matrix, 2 * X) and (2.0 * x).wires == [0]
    with recording() as tape:
        op.PauliX(wires=0) * 3.0
    assert len(tape) == 1 and np.allclose(tape[0].matrix, 3 * X)


def test_scaled_operator_inside_a_circuit():
    """H, then (Z * 1): <Z> stays 0."""

    def circuit():
        op.H(wires=0)
        op.PauliZ(wires=0) * 1.0

    assert np.allclose(_ev(circuit, [op.PauliZ(0, record=False)]), [0.0], atol=1e-10)


def test_addition():
    x, y = op.PauliX(wires=0, record=False), op.PauliY(wires=0, record=False)
    s = x + y
    assert np.allclose(s.matrix, X + Y) and s.wires == [0]
    assert np.allclose(s.matrix, s.matrix.conj().T)  # Hermitian stays Hermitian
    assert np.allclose((x + x).matrix, 2 * X)
    assert np.allclose((x + y).matrix, (y + x).matrix)
    with pytest.raises(ValueError):
        x + op.PauliX(wires=1, record=False)


def test_tensor_and_matrix_products():
    x0, z0 = op.PauliX(wires=0, record=False), op.PauliZ(wires=0, record=False)
    z1, y1 = op.PauliZ(wires=1, record=False), op.PauliY(wires=1, record=False)
    t = x0 @ z1  # disjoint wires: Kronecker product
    assert np.allclose(t.matrix, np.kron(X, Z)) and t.wires == [0, 1]
    assert t.matrix.shape == (4, 4)
    m = x0 @ z0  # same wire: matrix product
    assert np.allclose(m.matrix, X @ Z) and m.wires == [0]
    assert np.allclose((x0 * z0).matrix, X @ Z)  # '*' of two operations composes
    cx01, cx12 = op.CX(wires=[0, 1], record=False), op.CX(wires=[1, 2], record=False)
    p = cx01 @ cx12  # partial overlap: embed both, multiply
    assert np.allclose(p.matrix, np.kron(CXM, I2) @ np.kron(I2, CXM)) and p.wires == [0, 1, 2]
    i1 = op.Id(wires=1, record=False)
    assert np.allclose((x0 @ i1).matrix, np.kron(X, I2))
    three = x0 @ y1 @ op.PauliZ(wires=2, record=False)
    assert np.allclose(three.matrix, np.kron(np.kron(X, Y), Z)) and three.matrix.shape == (8, 8)


def test_prod_function_and_method():
    x, y, z = (op.PauliX(wires=0, record=False), op.PauliY(wires=1, record=False),
               op.PauliZ(wires=0, record=False))
    want = np.kron(X @ Z, Y)  # X(0) Z(0) (x) Y(1)
    for r in (op.prod(x, y, z), x.prod(y, z)):
        assert np.allclose(r.matrix, want) and r.wires == [0, 1]
        assert r.name == "Prod(PauliX*PauliY*PauliZ)"


def test_product_observable_execution():
    """<Z (x) Z> on a Bell pair is +1."""

    def bell():
        op.H(wires=0)
        op.CX(wires=[0, 1])

    zz = op.PauliZ(wires=0, record=False) @ op.PauliZ(wires=1, record=False)
    assert np.allclose(_ev(bell, [zz]), [1.0], atol=1e-10)
