"""Pins the oracle: reference known-answer tests, golden topologies generated from
the reference's own topologies.py, and the independent dense cross-check."""

import json
import os

import numpy as np
import pytest

import kat_cases
from oracle import circuits as oc
from oracle import dense, gates as og, sim as osim
from qml_essentials_b200 import operations as op
from qml_essentials_b200.tape import recording
from parity_cases import _tape_of

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "topologies.json")
_LAMBDAS = {"n-1": lambda n: n - 1, "n-2": lambda n: n - 2, "n//2": lambda n: n // 2}


def _oracle_run(circuit, n, typ, obs, args=()):
    """Record with the product's op classes (concrete values), evaluate with the
    ORACLE only (matrices of fixed gates are re-derived from oracle.gates by name)."""
    with recording() as tape:
        circuit(*[float(a) for a in args])
    spec = []
    for o in tape:
        name = type(o).__name__
        if isinstance(o, op.KrausChannel):
            params = [float(p) for p in o.parameters]
            extra = o.kraus_matrices() if name == "QubitChannel" else None
            spec.append((name, list(o.wires), params, extra))
        else:
            spec.append((name, list(o.wires), [float(p) for p in o.parameters], None))
    ospec = [(type(o).__name__, list(o.wires), []) for o in obs]
    return osim.simulate_and_measure(spec, n, typ, ospec)


def test_oracle_passes_reference_kats():
    kat_cases.run_all(_oracle_run)


def test_oracle_passes_pennylane_convention_closed_forms():
    """reference tests/test_jaqsi.py:494-661 compare with PennyLane; its documented gate and
    channel definitions give closed forms the oracle must reproduce (extra pin)."""
    kat_cases.run_pennylane_conventions(_oracle_run, atol=1e-12)


def test_golden_topologies_oracle_and_product():
    from qml_essentials_b200.topologies import Topology

    golden = json.load(open(GOLDEN))
    assert len(golden) >= 14
    for name, case in golden.items():
        kw = {k: _LAMBDAS.get(v, v) if isinstance(v, str) else v
              for k, v in case["kwargs"].items()}
        for n, pairs in case["pairs"].items():
            n = int(n)
            assert [list(p) for p in oc.TOPO[case["topology"]](n, **kw)] == pairs, name
            got = getattr(Topology, case["topology"])(n_qubits=n, **kw)
            assert [list(p) for p in got] == pairs, name


def test_survey_op_counts():
    """Tape sizes of the BASELINE configs."""
    cases = [(2, 1, "Circuit_19", None, False, 14), (4, 4, "Hardware_Efficient", None, False, 96),
             (6, 3, "Circuit_15", None, True, 96),
             (8, 4, "Strongly_Entangling", {"Depolarizing": 0.01, "AmplitudeDamping": 0.02},
              False, 472)]
    for n, L, ct, noise, skip, want in cases:
        P = oc.n_params_per_layer(ct, n)
        tape = oc.variational_tape(n, L, ct, np.zeros((L + 1, P)), [0.1], noise_params=noise,
                                   skip_encoding=skip)
        assert sum(e[0] != "Barrier" for e in tape) == want


@pytest.mark.parametrize("seed", range(4))
def test_einsum_oracle_matches_dense_oracle(seed):
    rng = np.random.default_rng(seed)
    n = 4
    noise = {"BitFlip": 0.02, "PhaseFlip": 0.03, "Depolarizing": 0.04,
             "MultiQubitDepolarizing": 0.05, "AmplitudeDamping": 0.06, "PhaseDamping": 0.07,
             "StatePreparation": 0.01, "Measurement": 0.02}
    for ct in ("Circuit_19", "Strongly_Entangling", "Circuit_6", "Circuit_9"):
        P = oc.n_params_per_layer(ct, n)
        params = rng.uniform(0, 2 * np.pi, (3, P))
        tape = oc.variational_tape(n, 2, ct, params, [0.37])
        assert np.allclose(osim.simulate_pure(tape, n), dense.run(tape, n), atol=1e-12)
    P = oc.n_params_per_layer("Circuit_19", 3)
    tape = oc.variational_tape(3, 1, "Circuit_19", rng.uniform(0, 6, (2, P)), [0.2],
                               noise_params=noise)
    assert np.allclose(osim.simulate_mixed(tape, 3), dense.run(tape, 3), atol=1e-12)


def test_oracle_gate_definitions():
    """Unitarity, Kraus completeness and a few textbook identities."""
    rng = np.random.default_rng(0)
    for name, k, npar, extra in [("RX", 1, 1, None), ("RY", 1, 1, None), ("RZ", 1, 1, None),
                                 ("Rot", 1, 3, None), ("CRX", 2, 1, None), ("CRY", 2, 1, None),
                                 ("CRZ", 2, 1, None), ("ControlledPhaseShift", 2, 1, None),
                                 ("RXX", 2, 1, None), ("RZX", 2, 1, None),
                                 ("PauliRot", 3, 1, "XYZ"),
                                 ("ControlledPauliRot", 3, 1, ("Y", 2))]:
        U = og.unitary_matrix(name, list(range(k)), list(rng.uniform(0, 6, npar)), extra)
        assert np.allclose(U @ U.conj().T, np.eye(2**k), atol=1e-12), name
    for name, params in [("BitFlip", [0.3]), ("PhaseFlip", [0.2]), ("DepolarizingChannel", [0.4]),
                         ("AmplitudeDamping", [0.25]), ("PhaseDamping", [0.35]),
                         ("ThermalRelaxationError", [0.1, 1.5, 1.0, 0.4]),
                         ("ThermalRelaxationError", [0.2, 1.0, 1.8, 0.3])]:
        Ks = og.kraus_matrices(name, params)
        assert np.allclose(sum(K.conj().T @ K for K in Ks), np.eye(2), atol=1e-12), name
    assert np.allclose(og.unitary_matrix("ControlledPhaseShift", [0, 1], [np.pi]),
                       og.unitary_matrix("CZ", [0, 1], []))
    assert np.allclose(og.unitary_matrix("Rot", [0], [0.3, 0.0, 0.4]),
                       og.unitary_matrix("RZ", [0], [0.7]))
    Ks = og.n_qubit_depolarizing_kraus(0.3, 2)
    assert len(Ks) == 16
    assert np.allclose(sum(K.conj().T @ K for K in Ks), np.eye(4), atol=1e-12)


def test_product_matrices_equal_oracle_matrices():
    """Host gate algebra (operations.py mirror) agrees with the oracle's."""
    th = [0.31, 1.7, -2.2]
    pairs = [
        (op.RX(th[0], wires=0, record=False), ("RX", [0], th[:1])),
        (op.Rot(*th, wires=0, record=False), ("Rot", [0], th)),
        (op.CRY(th[1], wires=[0, 1], record=False), ("CRY", [0, 1], th[1:2])),
        (op.ControlledPhaseShift(th[2], wires=[0, 1], record=False),
         ("ControlledPhaseShift", [0, 1], th[2:])),
        (op.RZX(th[0], wires=[0, 1], record=False), ("RZX", [0, 1], th[:1])),
        (op.PauliRot(th[1], "YXZ", wires=[0, 1, 2], record=False),
         ("PauliRot", [0, 1, 2], th[1:2], "YXZ")),
        (op.ControlledPauliRot(th[2], "Z", wires=[0, 1, 2], n_controls=2, record=False),
         ("ControlledPauliRot", [0, 1, 2], th[2:], ("Z", 2))),
        (op.H(0, record=False), ("H", [0], [])), (op.S(0, record=False), ("S", [0], [])),
        (op.CY(wires=[0, 1], record=False), ("CY", [0, 1], [])),
        (op.CCX(wires=[0, 1, 2], record=False), ("CCX", [0, 1, 2], [])),
        (op.CSWAP(wires=[0, 1, 2], record=False), ("CSWAP", [0, 1, 2], [])),
        (op.SWAP(wires=[0, 1], record=False), ("SWAP", [0, 1], [])),
    ]
    for o, spec in pairs:
        extra = spec[3] if len(spec) > 3 else None
        assert np.allclose(o.matrix, og.unitary_matrix(spec[0], spec[1], spec[2], extra),
                           atol=1e-14), spec[0]
    d = op.RX(th[0], wires=0, record=False).dagger()
    assert np.allclose(d.matrix, og.unitary_matrix("RX", [0], [-th[0]]))
    lifted = op.CX(wires=[0, 2], record=False).lifted_matrix(3)
    assert np.allclose(lifted, dense.lift(og.controlled(og.X), [0, 2], 3))
    for ch, spec in [
        (op.ThermalRelaxationError(0.1, 1.5, 1.0, 0.4, wires=0), ("ThermalRelaxationError",
                                                                 [0.1, 1.5, 1.0, 0.4])),
        (op.ThermalRelaxationError(0.2, 1.0, 1.8, 0.3, wires=0), ("ThermalRelaxationError",
                                                                 [0.2, 1.0, 1.8, 0.3])),
        (op.AmplitudeDamping(0.3, wires=0), ("AmplitudeDamping", [0.3])),
    ]:
        from qml_essentials_b200.compiler import kraus_superop

        assert np.allclose(kraus_superop(ch.kraus_matrices()),
                           kraus_superop(og.kraus_matrices(*spec)), atol=1e-12)
