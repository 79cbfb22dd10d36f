"""Operator-algebra expressions shared by `tools/gen_golden_reference_analysis.py` (evaluated
on the reference's `operations` module) and `tests/test_reference_golden.py` (evaluated on
the drop-in's): each entry maps an `operations`-like module to an Operation."""


def _h(op, m, w):
    return op.Hermitian(m, wires=w, record=False)


CASES = {
    "rx_dagger": lambda op: op.RX(0.3, wires=0, record=False).dagger(),
    "rot_power2": lambda op: op.Rot(0.1, 0.2, 0.3, wires=1, record=False).power(2),
    "crx_dagger": lambda op: op.CRX(0.7, wires=[1, 0], record=False).dagger(),
    "x_times_scalar": lambda op: op.PauliX(wires=0, record=False) * 2.5,
    "scalar_times_y": lambda op: -0.5 * op.PauliY(wires=2, record=False),
    "x_plus_z": lambda op: op.PauliX(wires=0, record=False) + op.PauliZ(wires=0, record=False),
    "x_matmul_y_other_wire": lambda op: op.PauliX(wires=0, record=False) @ op.PauliY(wires=1, record=False),
    "z_matmul_x_same_wire": lambda op: op.PauliZ(wires=0, record=False) @ op.PauliX(wires=0, record=False),
    "cx_matmul_rz": lambda op: op.CX(wires=[0, 1], record=False) @ op.RZ(0.4, wires=1, record=False),
    "sum_of_products": lambda op: (op.PauliZ(wires=0, record=False) @ op.PauliZ(wires=1, record=False))
    + 0.3 * (op.PauliX(wires=0, record=False) @ op.PauliX(wires=1, record=False)),
    "h_power3": lambda op: op.H(wires=0, record=False).power(3),
    "s_dagger": lambda op: op.S(wires=0).dagger(),
}

# callers replayed on both sides: name -> (n_qubits, n_layers, circuit_type, parameter sets)
CALLER_MODELS = {
    "he3": (3, 2, "Hardware_Efficient", 12),
    "c19": (2, 1, "Circuit_19", 1),
    "se4": (4, 1, "Strongly_Entangling", 9),
}
SPECTRUM_SETTINGS = {
    "default": dict(),
    "mfs2_shift_trim": dict(mfs=2, shift=True, trim=True),
    "mts2": dict(mfs=1, mts=2),
    "shift": dict(shift=True),
}
FCC_METHODS = ("pearson", "complex_pearson", "covariance", "spearman")
