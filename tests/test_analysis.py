"""Callers of the hot path: Coefficients / FCC, Expressibility, Entanglement
(coefficients.py, expressibility.py, entanglement.py of the reference).  CPU suite:
the circuits run through the oracle-backed interpreter; multi-rank sharding is
exercised with gloo, world size 2."""

import os
import socket
import subprocess
import sys
import warnings

import numpy as np
import pytest

from qml_essentials_b200.coefficients import FCC, Coefficients
from qml_essentials_b200.entanglement import Entanglement
from qml_essentials_b200.expressibility import Expressibility
from qml_essentials_b200.model import Model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(*a, **k):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return Model(*a, **k)


# ---------------------------------------------------------------- Coefficients
def test_spectrum_reconstructs_model_output():
    """evaluate_Fourier_series(get_spectrum(model)) == model(x) at off-grid points: an
    end-to-end check that does not share code with the transform."""
    m = _model(n_qubits=2, n_layers=2, circuit_type="Circuit_19")
    coeffs, freqs = Coefficients.get_spectrum(m, shift=True)
    assert coeffs.shape[0] == freqs.shape[0] == m.degree[0]
    x = np.array([0.3, 1.1, 2.9, 4.2])
    want = m(inputs=x.reshape(-1, 1), force_mean=True)
    got = Coefficients.evaluate_Fourier_series(coeffs, freqs, x)
    assert np.allclose(got, np.asarray(want).reshape(got.shape), atol=1e-10)
    # real signal: c(-k) = conj(c(k)); trim/shift keep the zero frequency centred
    assert np.allclose(coeffs, np.conj(coeffs[::-1]), atol=1e-12)


def test_spectrum_oversampling_trim_and_cap():
    m = _model(n_qubits=2, n_layers=1, circuit_type="Circuit_19")
    c1, f1 = Coefficients.get_spectrum(m, mfs=1, shift=True)
    c2, f2 = Coefficients.get_spectrum(m, mfs=2, shift=True, trim=True)
    assert len(f2) % 2 == 1 and len(f2) > len(f1)
    keep = np.isin(f2, f1)
    assert np.allclose(c2[keep], c1, atol=1e-10) and np.allclose(c2[~keep], 0, atol=1e-10)
    c3, f3 = Coefficients.get_spectrum(m, mfs=2, shift=True, trim=True, numerical_cap=1e-9)
    assert len(f3) <= len(f1) and np.all(np.abs(c3) > 0)
    psd = Coefficients.get_psd(c1)
    assert psd.shape == c1.shape and np.all(psd >= 0)


def test_spectrum_batched_params_keeps_sample_axis():
    m = _model(n_qubits=2, n_layers=1, circuit_type="Hardware_Efficient")
    m.initialize_params(repeat=5)
    c, f = Coefficients.get_spectrum(m, shift=True, trim=True)
    assert c.shape == (len(f), 5)
    one = _model(n_qubits=2, n_layers=1, circuit_type="Hardware_Efficient")
    one.params = m.params[3:4]
    c1, _ = Coefficients.get_spectrum(one, shift=True, trim=True)
    assert np.allclose(c[:, 3], c1, atol=1e-12)


# ---------------------------------------------------------------- correlation estimators
def test_correlation_estimators_against_numpy_and_scipy():
    from scipy.stats import spearmanr

    rng = np.random.default_rng(3)
    X = rng.normal(size=(200, 5))
    X[:, 1] += 0.5 * X[:, 0]
    assert np.allclose(FCC._pearson(X), np.corrcoef(X.T), atol=1e-12)
    assert np.allclose(FCC._covariance(X), np.cov(X.T), atol=1e-12)
    assert np.allclose(FCC._spearman(X), spearmanr(X).statistic, atol=1e-12)
    Z = X[:, :3] + 1j * rng.normal(size=(200, 3))
    Z = np.concatenate([Z, (np.exp(0.7j) * Z[:, :1])], axis=1)
    cp = FCC._complex_pearson(Z)
    assert np.isclose(abs(cp[0, 3]), 1.0) and np.isclose(np.angle(cp[0, 3]), 0.7)
    assert np.allclose(FCC._pearson(Z), np.corrcoef(np.concatenate([Z.real, Z.imag]).T),
                       atol=1e-12)
    # missing values: pairwise deletion
    Y = X.copy()
    Y[::7, 2] = np.nan
    ok = np.isfinite(Y[:, 2])
    # complex_pearson normalises pairwise; pearson (like the reference, coefficients.py:
    # 1486-1492) divides the pairwise covariance by the per-column standard deviations
    assert np.isclose(FCC._complex_pearson(Y)[0, 2].real, np.corrcoef(Y[ok, 0], Y[ok, 2])[0, 1])
    cov = FCC._covariance(Y)
    assert np.isclose(cov[0, 2], np.cov(Y[ok, 0], Y[ok, 2])[0, 1])
    assert np.isclose(FCC._pearson(Y)[0, 2], cov[0, 2] / np.sqrt(cov[0, 0] * cov[2, 2]))
    with pytest.raises(ValueError):
        FCC._correlate(X, "kendall")


def test_weighting_mean_matches_reference_unit_test():
    """tests/test_coefficients.py:939-952 of the reference."""
    fp = np.arange(16, dtype=float).reshape(4, 4)
    coeffs = np.array([[[1.0, 3.0], [-2.0, 4.0]], [[5.0, 7.0], [8.0, 10.0]]])
    w = np.abs(np.mean(coeffs, axis=-1)).T.reshape(-1)
    assert np.allclose(FCC._weighting_mean(fp, coeffs), fp * np.outer(w, w))


def test_fcc_bounded_and_fingerprint_shapes():
    m = _model(n_qubits=2, n_layers=2, circuit_type="Circuit_2", output_qubit=-1)
    for method in ("pearson", "complex_pearson", "spearman", "covariance"):
        v = FCC.get_fcc(model=m, n_samples=24, method=method)
        assert np.isfinite(v) and v >= 0 and (method == "covariance" or v <= 1)
    fp, (rf, cf) = FCC.get_fourier_fingerprint(model=m, n_samples=24)
    assert fp.shape == (len(rf), len(cf))
    fpw, _ = FCC.get_fourier_fingerprint(model=m, n_samples=24, weight=True)
    assert fpw.shape == fp.shape
    full, freqs = FCC.get_fourier_fingerprint(model=m, n_samples=24, trim_redundant=False)
    assert full.shape == (len(freqs), len(freqs))
    # the fast path equals "fingerprint then nanmean" (coefficients.py:1012-1031)
    a = FCC.get_fcc(model=m, n_samples=16, random_key=None)
    assert 0 <= a <= 1


def test_fcc_paper_value_hardware_efficient_vs_circuit_19():
    """Rank order of Fig. 3a (arXiv:2508.20868), reduced size: HE correlates its
    coefficients far more than Circuit_19 (reference: 0.080 vs 0.010 at 6 qubits)."""
    vals = {}
    for ct in ("Circuit_19", "Hardware_Efficient"):
        m = _model(n_qubits=3, n_layers=1, circuit_type=ct, output_qubit=-1, encoding=["RY"])
        vals[ct] = FCC.get_fcc(model=m, n_samples=400)
    assert vals["Hardware_Efficient"] > vals["Circuit_19"]


# ---------------------------------------------------------------- Expressibility
def test_haar_integral_closed_form_matches_quadrature(tmp_path, monkeypatch):
    from scipy import integrate

    monkeypatch.chdir(tmp_path)
    x, y = Expressibility.haar_integral(n_qubits=3, n_bins=20, cache=True)
    _, y2 = Expressibility.haar_integral(n_qubits=3, n_bins=20, cache=True)  # from cache
    q = [integrate.quad(Expressibility._haar_probability, i / 20, (i + 1) / 20, args=(3,))[0]
         for i in range(20)]
    assert np.allclose(y, q, atol=1e-12) and np.array_equal(y, y2) and len(x) == 20
    assert np.isclose(y.sum(), 1.0)
    assert abs(Expressibility.kullback_leibler_divergence(y, y2).mean()) < 1e-12


def test_pure_state_fidelity_equals_uhlmann_formula():
    """The device reduction |<a|b>|^2 against the reference's sqrtm formula
    (expressibility.py:48-66) on the same density matrices."""
    from scipy.linalg import sqrtm

    m = _model(n_qubits=3, n_layers=1, circuit_type="Circuit_19")
    fid = Expressibility._sample_state_fidelities(m, 6, kwargs={})
    rho = np.asarray(m(params=m.params, execution_type="density"))
    root = np.array([sqrtm(r) for r in rho[:6]])
    inner = root @ rho[6:] @ root
    ref = np.abs(np.trace(np.array([sqrtm(r) for r in inner]), axis1=1, axis2=2) ** 2)
    assert np.allclose(fid, ref, atol=1e-6)  # sqrtm of a rank-1 matrix is ill-conditioned
    assert np.allclose(fid, np.abs(np.einsum("bii->b", rho[:6] @ rho[6:])), atol=1e-12)


def test_expressibility_ranks_circuits_like_sim_et_al(tmp_path, monkeypatch):
    """Circuit_9 (0.6773) is far less expressive than Circuit_6 (0.0061) at one layer,
    4 qubits (Sim et al., Adv. Quantum Technol. 2019; reference tolerance 40 %)."""
    monkeypatch.chdir(tmp_path)
    kl = {}
    for ct in ("Circuit_9", "Circuit_6"):
        m = _model(n_qubits=4, n_layers=1, circuit_type=ct,
                   initialization_domain=[0, 4 * np.pi], data_reupload=False)
        kl[ct] = float(Expressibility.kl_divergence_to_haar(model=m, n_samples=3000,
                                                            n_bins=75).mean())
    assert abs(kl["Circuit_9"] - 0.6773) / 0.6773 < 0.4
    assert kl["Circuit_6"] < 0.05 < kl["Circuit_9"]


def test_noisy_fidelities_use_density_path():
    m = _model(n_qubits=2, n_layers=1, circuit_type="Circuit_19")
    f = Expressibility._sample_state_fidelities(
        m, 4, kwargs={"noise_params": {"Depolarizing": 0.05}})
    assert f.shape == (4,) and np.all((f >= 0) & (f <= 1 + 1e-9))


# ---------------------------------------------------------------- Entanglement
def test_meyer_wallach_limits_and_host_formula():
    m0 = _model(n_qubits=3, n_layers=1, circuit_type="No_Entangling")
    assert abs(Entanglement.meyer_wallach(m0, n_samples=10)) < 1e-12
    m = _model(n_qubits=3, n_layers=2, circuit_type="Strongly_Entangling")
    v = Entanglement.meyer_wallach(m, n_samples=20)
    rho = np.asarray(m(params=m.params, inputs=None, execution_type="density"))
    ref = Entanglement._compute_meyer_wallach_meas(rho, 3)
    assert np.isclose(v, ref.mean(), atol=1e-12) and 0 < v <= 1
    # current parameters, no resampling
    m.initialize_params(repeat=1)
    assert 0 <= Entanglement.meyer_wallach(m, n_samples=None) <= 1


def test_meyer_wallach_equals_bell_measurement():
    """tests/test_entanglement.py:288-313 of the reference (1e-5 at a fixed key)."""
    from qml_essentials_b200 import rng

    m = _model(n_qubits=2, n_layers=1, circuit_type="Strongly_Entangling")
    mw = Entanglement.meyer_wallach(m, n_samples=30, random_key=rng.key(7))
    bell = Entanglement.bell_measurements(m, n_samples=30, random_key=rng.key(7))
    assert abs(mw - bell) < 1e-5


def test_meyer_wallach_noisy_uses_complement_purity():
    m = _model(n_qubits=2, n_layers=1, circuit_type="Strongly_Entangling")
    noise = {"Depolarizing": 0.05}
    v = Entanglement.meyer_wallach(m, n_samples=5, noise_params=dict(noise))
    rho = np.asarray(m(params=m.params, inputs=None, execution_type="density",
                       noise_params=dict(noise)))
    assert np.isclose(v, Entanglement._compute_meyer_wallach_meas(rho, 2).mean(), atol=1e-12)


# ---------------------------------------------------------------- two ranks, gloo
_RANK_CODE = r"""
import os, sys, json, warnings
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch.distributed as dist
from qml_essentials_b200 import script, rng, parallel
from _interp_executor import InterpExecutor
script._set_executor_for_testing(InterpExecutor())
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    dist.init_process_group("gloo")
from qml_essentials_b200.model import Model
from qml_essentials_b200.coefficients import FCC
from qml_essentials_b200.entanglement import Entanglement
from qml_essentials_b200.expressibility import Expressibility
warnings.simplefilter("ignore")
out = {{}}
m = Model(n_qubits=2, n_layers=2, circuit_type="Circuit_19", output_qubit=-1)
for meth in ("pearson", "complex_pearson", "spearman", "covariance"):
    out["fcc_" + meth] = float(FCC.get_fcc(model=m, n_samples=21, random_key=rng.key(5), method=meth))
fp, _ = FCC.get_fourier_fingerprint(model=m, n_samples=21, random_key=rng.key(5), weight=True)
out["fp_weighted"] = np.nan_to_num(fp).tolist()
m3 = Model(n_qubits=3, n_layers=1, circuit_type="Strongly_Entangling")
out["mw"] = Entanglement.meyer_wallach(m3, n_samples=13, random_key=rng.key(9))
out["bell"] = Entanglement.bell_measurements(m3, n_samples=13, random_key=rng.key(9))
_, z = Expressibility.state_fidelities(n_samples=37, n_bins=10, model=m3, random_key=rng.key(11))
out["hist"] = z.tolist()
out["shard"] = list(parallel.shard_bounds(13))
if parallel.world()[0] == 0:
    print("RESULT " + json.dumps(out))
if dist.is_initialized():
    dist.destroy_process_group()
"""


def _run(world):
    code = _RANK_CODE.format(root=ROOT)
    env = dict(os.environ)
    if world == 1:
        cmd = [sys.executable, "-c", code]
    else:
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), "--no-python", sys.executable, "-c", code]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    import json

    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def test_batch_sharding_two_ranks_equals_single_process():
    """World size 2 over gloo: sample axis sharded, sufficient statistics all-reduced;
    the averages equal the single-process values."""
    one, two = _run(1), _run(2)
    assert one["shard"] == [0, 13] and two["shard"] == [0, 7]
    for k in one:
        if k == "shard":
            continue
        assert np.allclose(np.asarray(one[k]), np.asarray(two[k]), atol=1e-10), k


def test_concentratable_entanglement_matches_subset_purity_formula():
    """CE = 1 - 2^-n sum over qubit subsets of Tr(rho_subset^2) (arXiv:2104.06923, eq. 1),
    evaluated from the model's own state; the product path measures it with the 3n-qubit
    swap-test circuit (entanglement.py:471-577)."""
    import itertools

    from qml_essentials_b200 import jaqsi as js
    from qml_essentials_b200 import rng

    n = 2
    m = _model(n_qubits=n, n_layers=1, circuit_type="Strongly_Entangling")
    ce = Entanglement.concentratable_entanglement(m, n_samples=6, random_key=rng.key(3))
    rho = np.asarray(m(params=m.params, inputs=None, execution_type="density"))
    want = []
    for r in rho:
        tot = 0.0
        for k in range(n + 1):
            for sub in itertools.combinations(range(n), k):
                if not sub:
                    tot += 1.0
                    continue
                red = js.partial_trace(r[None], n, list(sub))[0]
                tot += float(np.real(np.trace(red @ red)))
        want.append(1 - tot / 2 ** n)
    assert abs(ce - np.mean(want)) < 1e-10
    m0 = _model(n_qubits=2, n_layers=1, circuit_type="No_Entangling")
    assert abs(Entanglement.concentratable_entanglement(m0, n_samples=3)) < 1e-12


def test_two_feature_spectrum_reconstructs_model_output():
    """n_input_feat = 2 (encoding ['RX', 'RY']): N-D FFT grid, per-axis frequencies, and the
    series evaluated at off-grid points equals the model (coefficients.py:109-150,172-236)."""
    m = _model(n_qubits=2, n_layers=1, circuit_type="Circuit_19", encoding=["RX", "RY"])
    assert m.n_input_feat == 2
    coeffs, freqs = Coefficients.get_spectrum(m, shift=True)
    assert coeffs.ndim == 2 and len(freqs) == 2
    assert coeffs.shape == (len(freqs[0]), len(freqs[1]))
    pts = np.array([[0.3, 1.7], [2.9, 0.2], [4.1, 5.5]])
    want = np.asarray(m(inputs=pts, force_mean=True)).reshape(-1)
    got = Coefficients.evaluate_Fourier_series(coeffs, list(freqs), pts)
    assert np.allclose(got, want, atol=1e-10)
    fcc = FCC.get_fcc(model=m, n_samples=12)
    assert 0 <= fcc <= 1


def test_stats_from_device_moments_reproduce_every_estimator():
    """The K + K + K^2 moments the GPU reduces (qmlb_coef_moments) carry everything the
    Pearson / complex-Pearson / covariance estimators of coefficients.py:1346-1498 need."""
    from qml_essentials_b200.coefficients import FCC, _Stats

    g = np.random.default_rng(5)
    c = g.normal(size=(200, 9)) + 1j * g.normal(size=(200, 9))  # (samples, coefficients)
    c[:, 3] = 0.5 * c[:, 1] + 0.1 * c[:, 3]
    s1, s2, cc = c.sum(axis=0), (np.abs(c) ** 2).sum(axis=0), c.conj().T @ c
    st = _Stats.from_moments(200, s1, s2, cc)
    assert np.allclose(st.covariance(1), FCC._covariance(c), atol=1e-12)
    assert np.allclose(st.complex_pearson(1), FCC._complex_pearson(c), atol=1e-12)
    # Pearson: real and imaginary parts are separate observations (2N of them)
    sp = _Stats.from_moments(400, s1.real + s1.imag, s2, cc.real)
    cov = sp.covariance(1)
    std = np.sqrt(np.diagonal(cov))
    assert np.allclose(np.clip(np.real(cov / (std[:, None] * std[None, :])), -1, 1),
                       FCC._pearson(c), atol=1e-12)
