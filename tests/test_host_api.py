"""Host logic on CPU: recording, compiler, Script/Model API contracts.  Circuits run
through the oracle's program interpreter (see conftest) - the same parity cases run
through the CUDA library in test_gpu_parity.py."""

import warnings

import numpy as np
import pytest

import kat_cases
import parity_cases as pc
from qml_essentials_b200 import compiler, jaqsi as js, memory, operations as op, rng
from qml_essentials_b200.ansaetze import Ansaetze, Circuit, Encoding
from qml_essentials_b200.gates import Gates
from qml_essentials_b200.model import Model
from qml_essentials_b200.script import Script
from qml_essentials_b200.symbolic import Sym, SymArray, SymbolicError
from qml_essentials_b200.tape import copy_to_tape, recording


def _run(circuit, n, typ, obs, args=()):
    return Script(circuit, n_qubits=n).execute(typ, obs=obs, args=args)


def test_reference_kats_through_product_host_path():
    kat_cases.run_all(_run)


def test_pennylane_convention_closed_forms_through_product_host_path():
    kat_cases.run_pennylane_conventions(_run, atol=1e-12)


def test_parity_every_gate_and_channel():
    assert max(pc.case_every_gate().values()) < 1e-10
    assert max(pc.case_every_channel().values()) < 1e-10


def test_parity_baseline_configs_reduced():
    for name, err in pc.case_baseline_configs().items():
        assert err < 1e-10, name


def test_parity_all_ansaetze_and_noise_keys():
    for name, err in {**pc.case_all_ansaetze(), **pc.case_noise_keys()}.items():
        assert err < 1e-10, name


@pytest.mark.parametrize("strategy", ["binary", "ternary", "golomb"])
def test_encoding_strategies(strategy):
    assert pc.case_model(2, 2, "Circuit_19", 5, 2, strategy=strategy) < 1e-10


def test_edge_cases_through_host_path():
    for prec in ("complex128", "complex64"):
        errs = pc.case_edge_cases(prec)
        assert max(errs.values()) < pc.TOL[prec], errs


def test_shots_bookkeeping_matches_oracle_bit_for_bit():
    r = pc.case_shots()
    assert r["count_mismatch"] == 0 and r["expval_err"] < 1e-12 and r["sums"] < 1e-12


# ---- Script contract (reference tests/test_jaqsi.py:701-833) ---------------------
def _rx(theta):
    op.RX(theta, wires=0)


def test_batched_equals_sequential_and_cos():
    s = Script(_rx)
    th = np.linspace(0, np.pi, 7)
    obs = [op.PauliZ(0)]
    batched = s.execute("expval", obs=obs, args=(th,), in_axes=(0,))
    assert batched.shape == (7, 1)
    assert np.allclose(batched[:, 0], np.cos(th), atol=1e-12)
    seq = np.stack([s.execute("expval", obs=obs, args=(t,)) for t in th])
    assert np.allclose(batched, seq, atol=1e-12)


def test_in_axes_none_broadcasts_and_mismatch_raises():
    def circ(theta, phi):
        op.RX(theta, wires=0)
        op.RY(phi, wires=1)

    s = Script(circ)
    th = np.array([0.1, 0.2, 0.3])
    r = s.execute("probs", args=(th, np.array(0.5)), in_axes=(0, None))
    assert r.shape == (3, 4)
    with pytest.raises(ValueError, match="in_axes has"):
        s.execute("probs", args=(th, 0.5), in_axes=(0,))


def test_batch_axis_other_than_zero():
    def circ(w):
        op.RX(w[0], wires=0)
        op.RY(w[1], wires=1)

    s = Script(circ)
    w = np.random.default_rng(0).uniform(0, 3, (2, 5))
    r = s.execute("state", args=(w,), in_axes=(1,))
    ref = np.stack([s.execute("state", args=(w[:, b],)) for b in range(5)])
    assert np.allclose(r, ref, atol=1e-12)


def test_result_shapes_and_n_qubits_inference():
    def circ():
        op.H(wires=0)
        op.CX(wires=[0, 2])

    s = Script(circ)  # n inferred from max wire (simulation.py:25-39)
    assert s.execute("state").shape == (8,)
    assert s.execute("probs").shape == (8,)
    assert s.execute("density").shape == (8, 8)
    assert s.execute("expval", obs=[op.PauliZ(0), op.PauliZ(3)]).shape == (2,)  # obs widens n


def test_state_on_noisy_tape_raises_and_unknown_type():
    def noisy():
        op.H(wires=0)
        op.BitFlip(0.1, wires=0)

    with pytest.raises(ValueError, match="not defined for mixed"):
        Script(noisy).execute("state")
    with pytest.raises(ValueError, match="Unknown measurement type"):
        Script(noisy).execute("bogus")


def test_operation_validation_errors():
    with pytest.raises(ValueError, match="expects 2 wire"):
        op.CX(wires=[0])
    with pytest.raises(ValueError, match="duplicate wires"):
        op.CX(wires=[1, 1])
    with pytest.raises(ValueError, match=r"p must be in \[0, 1\]"):
        op.BitFlip(1.5, wires=0)
    with pytest.raises(TypeError):
        op.BitFlip(0.1, wires=0).apply_to_state(np.array([1, 0]), 1)
    with pytest.raises(TypeError):
        op.DepolarizingChannel(0.1, wires=0).matrix


def test_tape_semantics_dagger_replaces_and_copy_to_tape():
    with recording() as tape:
        op.RX(0.3, wires=0).dagger()
        op.PauliX(wires=1).power(2)
        copy_to_tape(lambda: op.CX(wires=[0, 1]), offset=2)
    assert len(tape) == 3
    assert np.allclose(tape[0].matrix, op.RX(-0.3, wires=0, record=False).matrix)
    assert np.allclose(tape[1].matrix, np.eye(2))
    assert tape[2].wires == [2, 3]
    with recording() as outer:
        op.H(wires=0)
        with recording() as inner:
            op.S(wires=0)
        op.PauliZ(wires=0)
    assert [o.name for o in outer] == ["H", "PauliZ"] and [o.name for o in inner] == ["S"]


def test_affine_proxies():
    a = SymArray.leaves(0, (2, 3))
    e = 2.0 * a[1, 2] - a[0, 0] / 4 + 1.5
    assert e.const == 1.5 and e.terms == {(0, 5): 2.0, (0, 0): -0.25}
    assert (np.float64(3.0) * a[0, 1]).terms == {(0, 1): 3.0}
    assert a[..., 1].shape == (2,)
    assert a.mean().terms[(0, 0)] == pytest.approx(1 / 6)
    with pytest.raises(SymbolicError):
        a[0, 0] * a[0, 1]
    with pytest.raises(SymbolicError):
        float(a[0, 0])
    with pytest.raises(SymbolicError):
        bool(a[0, 0] > 0)


def test_non_affine_circuit_falls_back_to_per_element_recording():
    def circ(t):
        op.RX(np.cos(t) * t, wires=0)
        op.CX(wires=[0, 1])
        op.QubitUnitary(op.RY(t**2, wires=0, record=False).matrix, wires=1)

    ts = np.linspace(0.1, 3, 6)
    s = Script(circ, 2)
    got = s.execute("probs", args=(ts,), in_axes=(0,))
    ref = np.stack([s.execute("probs", args=(t,)) for t in ts])
    assert np.allclose(got, ref, atol=1e-12)


def test_compiler_fusion_counts():
    """cfg2: 96 tape gates -> 20 fused 1-qubit chains + 20 CX index shuffles;
    cfg4-like noisy circuit: channels cost no op of their own."""
    m = Model(4, 4, "Hardware_Efficient")
    m(inputs=np.linspace(0, 1, 3), params=np.zeros((2, 5, 12)) + 0.1)
    plan = [p for p in m.script._jit_cache.values() if hasattr(p, "program")][0]
    kinds = list(plan.program.ops["kind"])
    assert plan.n_ops == 96 + 25  # incl. barriers (script.py:270)
    assert kinds.count(compiler.OP_MAT) == 20 and kinds.count(compiler.OP_PERM) == 20
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = Model(3, 2, "Strongly_Entangling")
        m(inputs=np.linspace(0, 1, 3), execution_type="density",
          noise_params={"Depolarizing": 0.01, "AmplitudeDamping": 0.02})
    plan = [p for p in m.script._jit_cache.values() if hasattr(p, "program")][0]
    assert plan.density_program and plan.program.n_bits == 6
    n_chan = sum(isinstance(o, op.KrausChannel) for o in m.script._record(
        m.params, np.array([0.3]), noise_params=m.noise_params, random_key=rng.key(0)))
    assert n_chan > 40
    # ops: one 4x4 superchain per (layer-block, wire) + two 2-bit permutations per CX
    # (ket side and bra side), none for the channels
    assert len(plan.program.ops) <= 0.65 * plan.n_ops
    kinds = list(plan.program.ops["kind"])
    assert kinds.count(compiler.OP_PERM) == 2 * 18  # 18 CX -> ket-side + bra-side shuffles


def test_plan_cache_reuse_and_value_independence():
    m = Model(2, 1, "Circuit_19")
    x = np.linspace(-1, 1, 9)
    r1 = m(inputs=x, params=np.full((1, 2, 6), 0.3))
    n_plans = sum(hasattr(p, "program") for p in m.script._jit_cache.values())
    r2 = m(inputs=x, params=np.full((1, 2, 6), 0.9))  # new values, same plan
    r3 = m(inputs=np.linspace(-1, 1, 17), params=np.full((1, 2, 6), 0.9))  # new batch size
    assert sum(hasattr(p, "program") for p in m.script._jit_cache.values()) == n_plans
    assert not np.allclose(r1, r2) and r3.shape == (17, 2)
    # zero single input drops the encoding (model.py:782) -> a different plan
    m(inputs=None)
    assert sum(hasattr(p, "program") for p in m.script._jit_cache.values()) == n_plans + 1


# ---- Model contract (reference tests/test_model.py:107-150, 928-1053) -------------
def test_model_output_shapes():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = Model(3, 1, "Circuit_19")
        x = np.linspace(0, 1, 4)
        p = np.random.default_rng(0).uniform(0, 6, (5, *m._params_shape))
        assert m(params=p, inputs=x).shape == (4, 5, 3)
        assert m(params=p, inputs=x, force_mean=True).shape == (4, 5)
        assert m(params=p[0], inputs=x[:1]).shape == (3,)
        assert m(params=p, inputs=x, execution_type="probs").shape == (4, 5, 2, 2, 2)
        assert m(params=p, inputs=x, execution_type="density").shape == (4, 5, 8, 8)
        assert m(params=p, inputs=x, execution_type="state").shape == (4, 5, 8)
        m1 = Model(3, 1, "Circuit_19", output_qubit=0)
        assert m1(params=p, inputs=x).shape == (4, 5)
        assert m1(params=p, inputs=x, execution_type="density").shape == (4, 5, 2, 2)
        assert m1(params=p, inputs=x, execution_type="probs").shape == (4, 5, 2)
        mp = Model(3, 1, "Circuit_19", output_qubit=[[0, 1], [1, 2]])
        assert mp(params=p, inputs=x).shape == (4, 5, 2)


def test_model_batched_density_equals_singles_and_repeat_axis():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for a in Ansaetze.get_available(parameterized_only=True)[:6]:
            m = Model(2, 1, a.__name__)
            m.initialize_params(rng.key(1000), repeat=3)
            p = m.params
            singles = np.stack([m(params=p[i], execution_type="density") for i in range(3)])
            assert np.allclose(singles, m(params=p, execution_type="density"), atol=1e-12)
        m = Model(2, 1, "Circuit_19", repeat_batch_axis=[False, True, True])
        key = m.initialize_params(rng.key(1000), repeat=10)
        res = m(inputs=rng.uniform(key, (10, 1)))
        assert res.shape == (10, 2)


def test_model_parity_observable_and_partial_outputs_match_oracle():
    from oracle import circuits as oc, sim as osim

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = Model(3, 1, "Circuit_19", output_qubit=[[0, 2], [1, 2]])
        x = np.array([[0.4]])
        got = m(inputs=x)
        tape = oc.variational_tape(3, 1, "Circuit_19", m.params[0], [0.4])
        zz = np.kron(np.diag([1, -1]), np.diag([1, -1]))
        ref = osim.simulate_and_measure(tape, 3, "expval", [("Hermitian", [0, 2], [], zz),
                                                            ("Hermitian", [1, 2], [], zz)])
        assert np.allclose(got, ref, atol=1e-12)
        m2 = Model(3, 1, "Circuit_19", output_qubit=[0, 2])
        rho = osim.simulate_and_measure(tape, 3, "density")
        assert np.allclose(m2(inputs=x, execution_type="density"),
                           osim.partial_trace(rho, 3, [0, 2]), atol=1e-12)
        pr = osim.simulate_and_measure(tape, 3, "probs")
        assert np.allclose(m2(inputs=x, execution_type="probs").ravel(),
                           osim.marginalize_probs(pr, 3, [0, 2]).ravel(), atol=1e-12)


def test_model_gate_error_batch_semantics():
    """GateError jitters per element when batch_gate_error (unitary.py:231-246)."""
    from qml_essentials_b200.unitary import UnitaryGates

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = Model(2, 1, "Circuit_19")
        x = np.zeros((6, 1)) + 0.3
        noisy = m(inputs=x, noise_params={"GateError": 0.2})
        clean = m(inputs=x, noise_params={"GateError": 0.0})
        assert np.ptp(clean, axis=0).max() < 1e-12  # identical inputs, no jitter
        assert np.ptp(noisy, axis=0).min() > 1e-4   # each element drew its own jitter
        UnitaryGates.batch_gate_error = False
        try:
            same = Model(2, 1, "Circuit_19")(inputs=x, noise_params={"GateError": 0.2})
            assert np.ptp(same, axis=0).max() < 1e-12  # one draw broadcast to the batch
        finally:
            UnitaryGates.batch_gate_error = True


def test_custom_circuit_subclass_and_callable_encoding():
    class MyCircuit(Circuit):
        def n_params_per_layer(self, n_qubits):
            return 2 * n_qubits

        def n_pulse_params_per_layer(self, n_qubits):
            return 0

        def get_control_indices(self, n_qubits):
            return None

        def build(self, w, n_qubits, **kwargs):
            for q in range(n_qubits):
                Gates.RY(w[2 * q], wires=q, **kwargs)
                Gates.RZ(w[2 * q + 1], wires=q, **kwargs)
            for q in range(n_qubits - 1):
                Gates.CZ(wires=[q, q + 1], **kwargs)

    m = Model(3, 2, MyCircuit, encoding=[Gates.RX, "RY"])
    x = np.random.default_rng(1).uniform(-1, 1, (4, 2))
    got = m(inputs=x)
    assert got.shape == (4, 3)
    single = np.stack([m(inputs=x[i:i + 1]) for i in range(4)])
    assert np.allclose(got, single, atol=1e-12)


# ---- memory model (reference tests/test_jaqsi.py:1620-1983) -----------------------
def test_memory_estimates_and_chunking(monkeypatch):
    # register-resident states need no workspace; big states evolve in the output when they
    # can; everything else budgets the state AND the per-element table of gate matrices
    assert memory.estimate_peak_bytes(4, 1000, "expval", False, n_obs=4) < 1000 * 4 * 8 * 1.2
    big = memory.estimate_peak_bytes(8, 4096, "density", True, n_ops=472)
    out = 4096 * 4**8 * 16
    assert out < big <= int((out + 4096 * 472 * 16 * 16) * 1.1)  # in place inside the output
    probs = memory.estimate_peak_bytes(8, 4096, "probs", True)
    assert probs > 4096 * 4**8 * 16  # needs a state workspace
    # a mid-size statevector (shared-memory regime) is NOT free - the evolved
    # state or the matrix table lives in the workspace
    assert memory.estimate_peak_bytes(12, 4_000_000, "expval", False, n_obs=12, n_ops=100) > \
        4_000_000 * 2**12 * 16
    monkeypatch.setattr(memory, "available_memory_bytes", lambda: 7 * 1024**3)
    assert memory.compute_chunk_size(4, 10**5, "expval", False, 4) == 10**5
    c = memory.compute_chunk_size(8, 16384, "probs", True)
    assert 1 <= c < 16384
    assert memory.estimate_peak_bytes(8, c, "probs", True) <= 0.8 * 7 * 1024**3 + 16384 * 256 * 8
    out = memory.execute_chunked(lambda a: a * 2, (np.arange(10.0),), (0,), 10, 3)
    assert np.allclose(out, np.arange(10.0) * 2)


def test_chunked_equals_full(monkeypatch):
    def circ(t):
        op.RX(t, wires=0)
        op.CX(wires=[0, 1])
        op.RY(t * 0.5, wires=1)

    th = np.linspace(0, 3, 11)
    s = Script(circ, 2)
    full = s.execute("density", args=(th,), in_axes=(0,))
    s2 = Script(circ, 2)
    monkeypatch.setattr(memory, "compute_chunk_size", lambda *a, **k: 4)
    chunked = s2.execute("density", args=(th,), in_axes=(0,))
    assert np.allclose(full, chunked, atol=1e-10)


def test_qi_helpers():
    rho = Script(kat_cases.bell).execute("density")
    for keep in ([0], [1]):
        assert np.allclose(js.partial_trace(rho, 2, keep), 0.5 * np.eye(2), atol=1e-12)
    assert np.allclose(js.partial_trace(rho, 2, [0, 1]), rho)
    assert js.partial_trace(np.stack([rho, rho]), 2, [1]).shape == (2, 2, 2)
    pr = Script(kat_cases.bell).execute("probs")
    assert np.allclose(js.marginalize_probs(pr, 2, [0]), [[0.5, 0.5]])
    ob = js.build_parity_observable([0, 1])
    assert ob._pauli_label == "ZZ" and np.allclose(np.diag(ob.matrix), [1, -1, -1, 1])


def test_rng_streams_are_deterministic_and_split_independent():
    k = rng.key(1000)
    a, b = rng.split(k)
    assert not np.array_equal(a.data, b.data)
    assert np.array_equal(rng.split(k).data, rng.split(rng.key(1000)).data)
    u = rng.uniform(a, (1000,), 0, 2 * np.pi)
    assert 0 <= u.min() and u.max() < 2 * np.pi and abs(u.mean() - np.pi) < 0.2
    z = rng.normal(b, (4000,))
    assert abs(z.mean()) < 0.06 and abs(z.std() - 1) < 0.05
    m1, m2 = Model(2, 1, "Circuit_19", random_seed=7), Model(2, 1, "Circuit_19", random_seed=7)
    assert np.array_equal(m1.params, m2.params)
