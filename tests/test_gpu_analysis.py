"""Device post-processing kernels: grid DFT against numpy.fft, FCC
sufficient statistics against the host estimator, and Coefficients.get_spectrum / FCC through
the device path against the host path."""

import warnings

import numpy as np
import pytest

from qml_essentials_b200.script import get_executor

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision,tol", [("complex128", 1e-12), ("complex64", 2e-6)])
@pytest.mark.parametrize("n_x,n_p,n_obs", [(33, 50, 4), (264, 1024, 4), (7, 3, 1), (128, 37, 6)])
def test_grid_dft_matches_numpy(precision, tol, n_x, n_p, n_obs):
    import torch

    ex = get_executor()
    rng = np.random.default_rng(0)
    real = np.float64 if precision == "complex128" else np.float32
    ev = rng.uniform(-1, 1, (n_x, n_p, n_obs)).astype(real)
    got = ex.grid_dft(torch.from_numpy(ev).to(ex.device)).cpu().numpy()
    want = np.fft.fft(ev.astype(np.float64).mean(axis=2), axis=0) / n_x  # coefficients.py:135-150
    assert got.shape == (n_x, n_p)
    assert np.abs(got - want).max() < tol


@pytest.mark.parametrize("precision,tol", [("complex128", 1e-11), ("complex64", 1e-4)])
def test_coef_moments_match_host_statistics(precision, tol):
    import torch

    from qml_essentials_b200.coefficients import _Stats

    ex = get_executor()
    rng = np.random.default_rng(1)
    cdt = np.complex128 if precision == "complex128" else np.complex64
    coef = (rng.normal(size=(40, 300)) + 1j * rng.normal(size=(40, 300))).astype(cdt)
    rows = [0, 1, 2, 5, 9, 17, 39]
    s1, s2, cc = ex.coef_moments(torch.from_numpy(coef).to(ex.device), rows)
    sel = coef[rows].astype(np.complex128)       # (K, N)
    assert np.abs(s1.cpu().numpy() - sel.sum(axis=1)).max() < tol * 300
    assert np.abs(s2.cpu().numpy() - (np.abs(sel) ** 2).sum(axis=1)).max() < tol * 300
    assert np.abs(cc.cpu().numpy() - sel.conj() @ sel.T).max() < tol * 300
    # the statistics object built from the device moments equals the host one
    host = _Stats(sel.T)
    dev = _Stats.from_moments(300, s1.cpu().numpy(), s2.cpu().numpy(), cc.cpu().numpy())
    for f in _Stats.FIELDS:
        assert np.abs(getattr(host, f) - getattr(dev, f)).max() < tol * 300, f
    assert np.nanmax(np.abs(host.complex_pearson(2) - dev.complex_pearson(2))) < 1e-6


def test_get_spectrum_device_path_equals_host_path(monkeypatch):
    from qml_essentials_b200 import rng
    from qml_essentials_b200.coefficients import Coefficients
    from qml_essentials_b200.model import Model

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = Model(4, 2, "Hardware_Efficient")
        m.initialize_params(rng.key(7), repeat=64)
        dev_c, dev_f = Coefficients.get_spectrum(m, mfs=3, shift=True, trim=True)
        monkeypatch.setenv("QMLB_HOST_FFT", "1")
        host_c, host_f = Coefficients.get_spectrum(m, mfs=3, shift=True, trim=True)
        monkeypatch.delenv("QMLB_HOST_FFT")
        assert np.array_equal(np.asarray(dev_f), np.asarray(host_f))
        assert dev_c.shape == host_c.shape and np.abs(dev_c - host_c).max() < 1e-12


@pytest.mark.parametrize("method", ["pearson", "complex_pearson", "covariance"])
def test_fcc_device_moments_equal_host_route(monkeypatch, method):
    """FCC through circuit kernel -> grid DFT -> moments on the GPU (only K + K + K^2 numbers
    leave the device) against the reference's host route over the full coefficient array."""
    from qml_essentials_b200 import rng
    from qml_essentials_b200.coefficients import FCC
    from qml_essentials_b200.model import Model

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = Model(3, 2, "Circuit_19")
        a = FCC.get_fcc(m, n_samples=300, random_key=rng.key(3), method=method)
        monkeypatch.setenv("QMLB_HOST_FFT", "1")
        b = FCC.get_fcc(m, n_samples=300, random_key=rng.key(3), method=method)
    assert np.isfinite(a) and abs(a - b) < 1e-9
