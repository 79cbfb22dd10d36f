"""Parity through the C ABI on a B200: the CUDA kernels against the oracle on the
same seeded inputs (1e-10 complex128, 1e-5 complex64), the reference's KATs, the
three execution strategies against each other, bit-exact shot bookkeeping, and
size-independent invariants at BASELINE sizes."""

import os
import subprocess
import sys
import warnings

import numpy as np
import pytest

import kat_cases
import parity_cases as pc
from qml_essentials_b200 import operations as op
from qml_essentials_b200.model import Model
from qml_essentials_b200.script import Script, get_executor

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worst(d):
    return max(v for v in d.values() if isinstance(v, (int, float)))


def test_native_library_is_the_executor():
    ex = get_executor()
    assert ex.name == "cuda-sm100a"
    n0 = ex.launch_count()
    Script(kat_cases.bell, 2).execute("probs")
    assert ex.launch_count() > n0


def test_reference_kats_on_gpu():
    kat_cases.run_all(lambda c, n, t, o, args=(): Script(c, n_qubits=n).execute(
        t, obs=o, args=args))
    kat_cases.run_all(lambda c, n, t, o, args=(): Script(
        c, n_qubits=n, precision="complex64").execute(t, obs=o, args=args), atol=1e-5)


def test_pennylane_convention_closed_forms_on_gpu():
    """Closed forms of the circuits the reference compares with PennyLane
    (tests/test_jaqsi.py:494-661): controlled gates, Rot, every 1-qubit channel."""
    kat_cases.run_pennylane_conventions(lambda c, n, t, o, args=(): Script(
        c, n_qubits=n).execute(t, obs=o, args=args), atol=1e-10)
    kat_cases.run_pennylane_conventions(lambda c, n, t, o, args=(): Script(
        c, n_qubits=n, precision="complex64").execute(t, obs=o, args=args), atol=1e-5)


@pytest.mark.parametrize("precision", ["complex128", "complex64"])
def test_every_gate_and_channel(precision):
    tol = pc.TOL[precision]
    assert _worst(pc.case_every_gate(precision)) < tol
    assert _worst(pc.case_every_channel(precision)) < tol


@pytest.mark.parametrize("precision", ["complex128", "complex64"])
def test_baseline_configs_reduced(precision):
    for name, err in pc.case_baseline_configs(precision).items():
        assert err < pc.TOL[precision], name


@pytest.mark.parametrize("precision", ["complex128", "complex64"])
def test_all_ansaetze_and_noise_keys(precision):
    for name, err in {**pc.case_all_ansaetze(precision), **pc.case_noise_keys(precision)}.items():
        assert err < pc.TOL[precision], name


@pytest.mark.parametrize("precision", ["complex128", "complex64"])
def test_shots_counts_bit_exact(precision):
    r = pc.case_shots(precision)
    assert r["count_mismatch"] == 0
    assert r["expval_err"] < 1e-12 and r["sums"] < 1e-12


@pytest.mark.parametrize("precision", ["complex128", "complex64"])
def test_edge_cases(precision):
    """Empty and permutation-only circuits, one qubit, ragged last CTA, one shot."""
    errs = pc.case_edge_cases(precision)
    assert max(errs.values()) < pc.TOL[precision], errs


@pytest.mark.parametrize("n", [1, 2, 3, 5, 6, 7, 9, 10, 12])
def test_statevector_sizes_across_kernel_regimes(n):
    """n <= 5 register kernel, 6..7 warp teams, 8..13 CTA teams (complex128)."""
    L = 1 if n > 8 else 2
    assert pc.case_model(n, L, "Circuit_19" if n > 1 else "No_Entangling", 3, 2, "probs") < 1e-10
    if n >= 2:
        assert pc.case_model(n, L, "Hardware_Efficient", 2, 2, "expval") < 1e-10


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5])
def test_noisy_density_sizes(n):
    noise = {"Depolarizing": 0.01, "AmplitudeDamping": 0.02, "PhaseFlip": 0.005}
    assert pc.case_model(n, 2, "Strongly_Entangling" if n > 1 else "No_Entangling", 2, 2,
                         "density", noise=noise) < 1e-10


def test_streamed_strategy_matches_resident_strategy():
    """Force the HBM-streaming register-group passes (k_stream) on small states, both
    precisions (separate processes: the strategy is an environment switch read at
    program creation).  Covers lean passes (1-qubit chains, CRX, CX, 4x4 superoperators)
    and heavy ones (16x16 two-qubit channel, CCX / CSWAP permutations)."""
    code = ("import sys; sys.path.insert(0, 'tests'); import parity_cases as pc; "
            "P = sys.argv[1]; tol = pc.TOL[P]; "
            "e = max(pc.case_model(7, 2, 'Hardware_Efficient', 2, 3, 'state', precision=P),"
            " pc.case_model(8, 1, 'Circuit_19', 2, 2, 'expval', precision=P),"
            " pc.case_model(3, 2, 'Strongly_Entangling', 2, 2, 'density', precision=P,"
            " noise={'Depolarizing': 0.01, 'AmplitudeDamping': 0.02}),"
            " pc.case_model(4, 1, 'Circuit_6', 1, 2, 'probs', precision=P, noise={'BitFlip': 0.1,"
            " 'MultiQubitDepolarizing': 0.05}),"
            " max(pc.case_every_gate(P, n=6).values())); print('ERR', e); assert e < tol")
    for prec in ("complex128", "complex64"):
        env = dict(os.environ, QMLB_FORCE_STRATEGY="2")
        r = subprocess.run([sys.executable, "-c", code, prec], cwd=ROOT, env=env,
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr


def test_chunked_equals_full_on_device(monkeypatch):
    from qml_essentials_b200 import memory

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = Model(4, 2, "Circuit_19")
        x = np.linspace(-1, 1, 37)
        full = m(inputs=x, execution_type="density")
        m2 = Model(4, 2, "Circuit_19")
        # Script chunks on the library's exact workspace figure (CudaExecutor.peak_bytes):
        # pretend the device has 60 kB free -> about 10 elements per chunk
        monkeypatch.setattr(memory, "available_memory_bytes", lambda: 60_000)
        assert np.array_equal(full, m2(inputs=x, execution_type="density"))
        chunks = [v for k, v in m2.script._jit_cache.items() if k[0] == "_mem"]
        assert chunks and 1 <= chunks[0] < 37


def test_full_size_cfg2_invariants_and_sample_parity():
    """BASELINE config 2 at full size (270 336 evals): |<Z>| <= 1, batch-order
    consistency against single-parameter calls, and oracle parity on a sample."""
    import bench

    model, params, inputs = bench.workload()
    res = model(params=params, inputs=inputs)
    assert res.shape == (264, 1024, 4) and np.all(np.abs(res) <= 1 + 1e-12)
    one = model(params=params[17], inputs=inputs)
    assert np.allclose(res[:, 17], one, atol=1e-12)
    _, cpu_ev, _ = bench.oracle_cpu_evals_per_s(params, inputs, 4, os.cpu_count(), repeats=1)
    assert np.abs(res[:, :4] - cpu_ev).max() < 1e-10
    m32 = bench.workload()[0]
    m32.script.precision = "complex64"
    assert np.abs(m32(params=params[:64], inputs=inputs) - res[:, :64]).max() < 1e-5


def test_full_size_cfg3_density_invariants():
    """Config 3 circuit (n=6, Circuit_15): trace 1, Hermitian, idempotent (pure)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = Model(6, 3, "Circuit_15")
        p = np.random.default_rng(1000).uniform(0, 2 * np.pi, (2000, *m._params_shape))
        rho = m(params=p, execution_type="density")
    assert rho.shape == (2000, 64, 64)
    assert np.allclose(np.trace(rho, axis1=1, axis2=2), 1, atol=1e-10)
    assert np.allclose(rho, np.conj(np.swapaxes(rho, 1, 2)), atol=1e-12)
    assert np.allclose(rho[:50] @ rho[:50], rho[:50], atol=1e-10)


def test_noisy_density_n7_streamed_invariants():
    """Noisy density beyond shared memory (n = 7: 4^7 x 16 B = 256 KiB per state)
    runs as streamed tile passes: valid state + agreement with probs output."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = Model(7, 2, "Strongly_Entangling")
        x = np.linspace(-1, 1, 6)
        noise = {"Depolarizing": 0.01, "AmplitudeDamping": 0.02}
        rho = m(inputs=x, execution_type="density", noise_params=dict(noise))
        pr = m(inputs=x, execution_type="probs", noise_params=dict(noise))
        ev = m(inputs=x, execution_type="expval", noise_params=dict(noise))
    assert np.allclose(np.trace(rho, axis1=1, axis2=2), 1, atol=1e-10)
    assert np.allclose(rho, np.conj(np.swapaxes(rho, 1, 2)), atol=1e-12)
    pur = np.real(np.einsum("bij,bji->b", rho, rho))
    assert np.all(pur < 1 - 1e-3) and np.all(pur > 1 / 128)
    assert np.allclose(np.real(np.einsum("bii->bi", rho)), pr.reshape(6, -1), atol=1e-12)
    z0 = np.array([1 - 2 * ((i >> 6) & 1) for i in range(128)])
    assert np.allclose(ev[:, 0], pr.reshape(6, -1) @ z0, atol=1e-12)
    assert pc.case_model(7, 1, "Strongly_Entangling", 1, 1, "density", noise=noise) < 1e-10


def test_device_side_purity_and_overlap_reductions():
    import torch

    from oracle import sim as osim

    ex = get_executor()
    rng = np.random.default_rng(3)
    B, n = 10, 5
    st = rng.normal(size=(B, 2**n)) + 1j * rng.normal(size=(B, 2**n))
    st /= np.linalg.norm(st, axis=1, keepdims=True)
    t = torch.from_numpy(st).to(ex.device)
    pur = ex.purities(t, n, False).cpu().numpy()
    fid = ex.overlap_fidelities(t, n).cpu().numpy()
    for b in range(B):
        rho = np.outer(st[b], st[b].conj())
        for q in range(n):
            r = osim.partial_trace(rho, n, [q])
            assert abs(pur[b, q] - np.real(np.trace(r @ r))) < 1e-12
    assert np.allclose(fid, np.abs(np.einsum("bi,bi->b", st[:5].conj(), st[5:])) ** 2,
                       atol=1e-12)
    rho_b = torch.from_numpy(np.einsum("bi,bj->bij", st, st.conj())).to(ex.device)
    assert np.allclose(ex.purities(rho_b, n, True).cpu().numpy(), pur, atol=1e-12)


def test_analysis_callers_through_device_reductions():
    """Coefficients / Expressibility / Entanglement on the GPU: states stay in HBM,
    purities and pair fidelities are reduced by qmlb_purity / qmlb_overlap_fidelity."""
    from qml_essentials_b200.coefficients import FCC, Coefficients
    from qml_essentials_b200.entanglement import Entanglement
    from qml_essentials_b200.expressibility import Expressibility

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = Model(4, 2, "Hardware_Efficient")
        coeffs, freqs = Coefficients.get_spectrum(m, shift=True)
        x = np.array([0.4, 2.2, 5.1])
        want = np.asarray(m(inputs=x.reshape(-1, 1), force_mean=True)).reshape(-1)
        assert np.allclose(Coefficients.evaluate_Fourier_series(coeffs, freqs, x), want,
                           atol=1e-10)
        assert 0 <= FCC.get_fcc(model=m, n_samples=64) <= 1

        m6 = Model(6, 3, "Circuit_15")
        mw = Entanglement.meyer_wallach(m6, n_samples=200)
        rho = m6(params=m6.params, inputs=None, execution_type="density")
        assert abs(mw - Entanglement._compute_meyer_wallach_meas(rho, 6).mean()) < 1e-10
        noise = {"Depolarizing": 0.02}
        m3 = Model(3, 2, "Strongly_Entangling")
        mwn = Entanglement.meyer_wallach(m3, n_samples=20, noise_params=dict(noise))
        rho = m3(params=m3.params, inputs=None, execution_type="density",
                 noise_params=dict(noise))
        assert abs(mwn - Entanglement._compute_meyer_wallach_meas(rho, 3).mean()) < 1e-10

        fid = Expressibility._sample_state_fidelities(m6, 500, kwargs={})
        rho = m6(params=m6.params, execution_type="density")
        ref = np.abs(np.einsum("bij,bji->b", rho[:500], rho[500:]))
        assert np.allclose(fid, ref, atol=1e-10)
        kl = Expressibility.kl_divergence_to_haar(model=m6, n_samples=2000, n_bins=75)
        assert np.isfinite(kl).all() and kl.mean() >= 0


def test_streamed_statevector_n18_matches_oracle():
    """Natural strategy 2 (state in HBM, k_stream passes + one-sweep <Z_q>): 2^18
    amplitudes against the oracle, both precisions."""
    assert pc.case_model(18, 2, "Hardware_Efficient", 1, 1, "expval") < 1e-10
    assert pc.case_model(18, 1, "Circuit_19", 2, 1, "expval", precision="complex64") < 1e-5
    assert pc.case_model(15, 2, "Strongly_Entangling", 1, 2, "probs") < 1e-10


def test_large_n_product_state_invariant():
    """n = 27 (1 GiB state, complex64): a circuit without entanglers leaves a product
    state, so <Z_q> of qubits 3g..3g+2 equals the 3-qubit model run with the same
    parameters (register kernel, oracle-checked above) - a size-independent check of
    pass scheduling, 32-bit index arithmetic and the one-sweep reduction."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        n = 27
        m = Model(n, 1, "No_Entangling", precision="complex64")
        p = np.random.default_rng(5).uniform(0, 2 * np.pi, (1, *m._params_shape))
        x = np.array([[0.7]])
        ev = np.asarray(m(params=p, inputs=x)).reshape(-1)
        per = p.reshape(m._params_shape[0], n, -1)  # (layers', qubits, params per qubit)
        small = Model(3, 1, "No_Entangling")
        want = np.concatenate([
            np.asarray(small(params=per[:, g:g + 3].reshape(1, per.shape[0], -1),
                             inputs=x)).reshape(-1) for g in range(0, n, 3)])
    assert np.allclose(ev, want, atol=2e-5)
