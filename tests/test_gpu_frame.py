"""On-chip frame engine (strategy 3) on a B200 through the C ABI: parity with the oracle
(1e-10 complex128 / 1e-5 complex64) for single-CTA states, several states per CTA and
cluster-resident states (DSMEM relayouts), and agreement with the streamed strategy."""

import os
import warnings

import numpy as np
import pytest

import parity_cases as pc
from qml_essentials_b200.model import Model
from qml_essentials_b200.script import get_executor

pytestmark = pytest.mark.gpu
NOISE = {"Depolarizing": 0.01, "AmplitudeDamping": 0.02}


@pytest.mark.parametrize("precision", ["complex128", "complex64"])
@pytest.mark.parametrize("n,L,ct,B_I,B_P,typ,noise", [
    (6, 3, "Circuit_15", 0, 37, "density", None),
    (6, 3, "Circuit_15", 0, 300, "state", None),
    (6, 2, "Circuit_19", 5, 3, "expval", None),
    (7, 2, "Hardware_Efficient", 3, 5, "probs", None),
    (9, 2, "Strongly_Entangling", 2, 3, "expval", None),
    (11, 1, "Circuit_19", 2, 2, "expval", None),
    (12, 1, "Hardware_Efficient", 1, 2, "state", None),
    (13, 1, "Strongly_Entangling", 2, 1, "probs", None),
    (3, 2, "Strongly_Entangling", 4, 2, "density", NOISE),
    (4, 2, "Strongly_Entangling", 3, 2, "expval", NOISE),
    (5, 1, "Hardware_Efficient", 2, 2, "probs", NOISE),
    (6, 1, "Circuit_19", 2, 1, "expval", {"BitFlip": 0.05, "PhaseDamping": 0.1}),
    (4, 1, "Circuit_6", 2, 1, "probs", {"BitFlip": 0.1, "MultiQubitDepolarizing": 0.05}),
])
def test_frame_single_cta_parity(precision, n, L, ct, B_I, B_P, typ, noise):
    err = pc.case_model(n, L, ct, B_I, B_P, typ, noise, precision=precision)
    assert err < pc.TOL[precision]


@pytest.mark.parametrize("precision", ["complex128", "complex64"])
@pytest.mark.parametrize("n,L,ct,B_I,B_P,typ,noise", [
    (14, 1, "Hardware_Efficient", 1, 2, "expval", None),
    (15, 1, "Circuit_19", 2, 1, "probs", None),
    (7, 1, "Strongly_Entangling", 2, 1, "probs", {"Depolarizing": 0.02}),
    (7, 1, "Strongly_Entangling", 1, 2, "density", NOISE),
    (8, 1, "Strongly_Entangling", 3, 1, "expval", NOISE),
])
def test_frame_cluster_parity(precision, n, L, ct, B_I, B_P, typ, noise):
    err = pc.case_model(n, L, ct, B_I, B_P, typ, noise, precision=precision)
    assert err < pc.TOL[precision]


@pytest.mark.parametrize("typ", ["expval", "probs", "density"])
def test_config4_full_size_vs_oracle(typ):
    """BASELINE config 4 itself - Model(8, 4, 'Strongly_Entangling'), depolarizing +
    amplitude damping - on 2 inputs against the oracle at 1e-10."""
    err = pc.case_model(8, 4, "Strongly_Entangling", 2, 1, typ, NOISE, precision="complex128")
    assert err < 1e-10


@pytest.mark.parametrize("precision", ["complex128", "complex64"])
@pytest.mark.parametrize("n,L,ct,typ", [
    (7, 2, "Strongly_Entangling", "probs"), (7, 1, "Circuit_10", "density"),
    (8, 1, "Circuit_19", "probs"), (6, 2, "Circuit_5", "density"),
])
def test_two_qubit_channels_on_cluster_sized_states(precision, n, L, ct, typ):
    """16x16 superoperators (MultiQubitDepolarizing) run the HEAVY kernel variant, whose
    relayout holds at most 32 amplitudes per thread: complex64 must not take the
    2^14-amplitude tile (found by tools/fuzz_gpu.py: the relayout was silently skipped)."""
    noise = {"BitFlip": 0.01, "MultiQubitDepolarizing": 0.02}
    err = pc.case_model(n, L, ct, 2, 2, typ, noise, precision=precision)
    assert err < pc.TOL[precision]


def test_frame_equals_streamed_strategy(monkeypatch):
    """The same circuits through the HBM-streaming kernels (QMLB_FRAME=0) and the frame
    engine: independent schedules, same numbers."""
    rng = np.random.default_rng(3)

    def run(n, L, ct, typ, noise, B):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = Model(n, L, ct)
            params = rng.uniform(0, 2 * np.pi, (B, *m._params_shape))
            inputs = np.linspace(-1, 1, 3).reshape(3, 1)
            return np.asarray(m(params=params, inputs=inputs, execution_type=typ,
                                noise_params=dict(noise) if noise else None)), params

    for n, L, ct, typ, noise in ((8, 2, "Strongly_Entangling", "expval", NOISE),
                                 (14, 2, "Hardware_Efficient", "expval", None),
                                 (10, 2, "Circuit_15", "probs", None)):
        state = rng.bit_generator.state
        a, _ = run(n, L, ct, typ, noise, 2)
        rng.bit_generator.state = state
        monkeypatch.setenv("QMLB_FRAME", "0")
        monkeypatch.setenv("QMLB_PTM", "0")
        b, _ = run(n, L, ct, typ, noise, 2)
        monkeypatch.delenv("QMLB_FRAME")
        monkeypatch.delenv("QMLB_PTM")
        assert np.abs(a - b).max() < 1e-10, (n, ct)


def test_frame_engine_is_selected():
    """Config 3 is planned as strategy 3 on the device; config 4 as strategy 5 (Pauli basis)
    (the density matrix is rebuilt from the coefficients by a phased Walsh-Hadamard transform)."""
    ex = get_executor()
    from qml_essentials_b200 import backend
    import test_cabi

    for n, L, ct, typ, noise, want in ((6, 3, "Circuit_15", "expval", None, 3),
                                       (8, 4, "Strongly_Entangling", "expval", NOISE, 5),
                                       (8, 4, "Strongly_Entangling", "probs", NOISE, 5),
                                       (8, 4, "Strongly_Entangling", "density", NOISE, 5)):
        plan = test_cabi._plan_of(n, L, ct, "complex128", typ, noise)
        h = backend.ProgramHandle(ex.lib, plan.program, plan.out_type, plan.obs_recs,
                                  plan.obs_pool, "complex128")
        assert h.strategy == want, (ct, typ)


@pytest.mark.parametrize("precision", ["complex128", "complex64"])
@pytest.mark.parametrize("n,L,typ,B_I,B_P", [
    (8, 4, "expval", 3, 2), (8, 2, "probs", 2, 2), (7, 2, "expval", 5, 3), (5, 3, "probs", 9, 4),
    (3, 2, "expval", 17, 5), (2, 2, "probs", 4, 4),
    (8, 2, "density", 2, 1), (7, 1, "density", 3, 1), (6, 2, "density", 2, 3), (5, 2, "density", 5, 2),
    (4, 2, "density", 7, 3), (3, 3, "density", 9, 5),
])
def test_pauli_basis_equals_complex_engine(monkeypatch, precision, n, L, typ, B_I, B_P):
    """Strategy 5 (real Pauli coefficients, CX folded + sign op) against the complex (ket,
    bra) evolution of strategy 3 on the same noisy circuits - two independent algorithms."""
    rng = np.random.default_rng(11)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = Model(n, L, "Strongly_Entangling", precision=precision)
        params = rng.uniform(0, 2 * np.pi, (B_P, *m._params_shape))
        inputs = rng.uniform(-1, 1, (B_I, 1))
        noise = {"Depolarizing": 0.01, "AmplitudeDamping": 0.02, "PhaseFlip": 0.03}
        call = lambda: np.asarray(m(params=params, inputs=inputs, execution_type=typ,
                                    noise_params=dict(noise)))
        a = call()
        monkeypatch.setenv("QMLB_PTM", "0")
        b = call()
    assert np.abs(a - b).max() < (1e-11 if precision == "complex128" else 2e-6)


@pytest.mark.parametrize("precision", ["complex128", "complex64"])
@pytest.mark.parametrize("n,L,ct,B_I,B_P", [
    (18, 2, "Hardware_Efficient", 1, 1),
    (18, 1, "Hardware_Efficient", 2, 2),
    (19, 1, "Circuit_19", 1, 1),
    (18, 1, "Strongly_Entangling", 1, 2),
])
def test_streamed_tiles_parity(precision, n, L, ct, B_I, B_P):
    """Strategy 4: HBM-resident state, tile passes through shared memory moved by bulk (TMA)
    copies - against the oracle."""
    err = pc.case_model(n, L, ct, B_I, B_P, "expval", precision=precision)
    assert err < pc.TOL[precision]


def test_streamed_tiles_equal_register_group_stream(monkeypatch):
    """n = 24: the tile passes (strategy 4) and round 1's register-group passes
    (QMLB_FSTREAM=0 -> strategy 2) are independent schedules of the same circuit."""
    from qml_essentials_b200 import backend
    import test_cabi

    ex = get_executor()
    outs = {}
    for fs in ("1", "0"):
        monkeypatch.setenv("QMLB_FSTREAM", fs)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = Model(24, 3, "Hardware_Efficient", precision="complex64")
            params = np.random.default_rng(3).uniform(0, 2 * np.pi, (1, *m._params_shape))
            outs[fs] = np.asarray(m(params=params, inputs=np.array([[0.5]])))
        plan = [p for p in m.script._jit_cache.values() if hasattr(p, "program") and p.device][-1]
        assert list(plan.device.values())[0].strategy == (4 if fs == "1" else 2)
    assert np.abs(outs["1"] - outs["0"]).max() < 2e-5
    assert np.all(np.abs(outs["1"]) <= 1 + 1e-5)
