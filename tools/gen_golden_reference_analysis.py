"""Golden vectors from the reference's OWN analysis helpers (array in, array out).

    python tools/gen_golden_reference_analysis.py   # writes tests/golden/reference_analysis.npz

Third companion of `tools/gen_golden_reference.py` (read its header first): the unmodified
`coefficients.py` (`FCC._correlate` and its four methods, `_calculate_mask`,
`_flat_frequencies`, the two weightings, `calculate_fcc`, `Coefficients.get_psd`,
`evaluate_Fourier_series`) and `expressibility.py` (`_haar_probability`, `haar_integral`,
`kullback_leibler_divergence`) run on the NumPy stand-in for JAX with seeded random arrays.
These are the host-side formulas behind the "next" rows of SURVEY.md 8(f); the drop-in's own
versions (`qml_essentials_b200/coefficients.py`, `expressibility.py`) are held to them in
`tests/test_reference_golden.py`.
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "jax_numpy_shim"))
sys.path.insert(1, "/root/reference")
import stub_missing  # noqa: E402

stub_missing.install()

import numpy as np  # noqa: E402

import qml_essentials.coefficients as rc  # noqa: E402
import qml_essentials.expressibility as rx  # noqa: E402

assert rc.__file__.startswith("/root/reference/"), rc.__file__


def main():
    rng = np.random.default_rng(20260003)
    store = {}
    warnings.simplefilter("ignore")

    # correlation of Fourier coefficients over parameter samples: rows = samples
    mat = rng.normal(size=(40, 7)) + 1j * rng.normal(size=(40, 7))
    mat[:, 3] = 0.6 * mat[:, 1] + 0.4 * mat[:, 3]          # a correlated pair
    mat[:, 5] = 0.0                                          # a dead coefficient (zero variance)
    holes = mat.copy()
    holes[rng.integers(0, 40, 6), rng.integers(0, 7, 6)] = np.nan   # missing values
    store["corr_mat"], store["corr_mat_holes"] = mat, holes
    for method in ("pearson", "complex_pearson", "spearman", "covariance"):
        store[f"corr_{method}"] = np.asarray(rc.FCC._correlate(mat, method=method))
        store[f"corr_holes_{method}"] = np.asarray(rc.FCC._correlate(holes, method=method))

    # frequency masks / labels, one and two input features
    f1 = np.fft.fftshift(np.fft.fftfreq(9, 1 / 9))
    store["freqs1"] = f1
    store["mask1"] = np.asarray(rc.FCC._calculate_mask(f1))
    store["flat1"] = np.asarray(rc.FCC._flat_frequencies(f1))
    f2 = np.stack([np.fft.fftshift(np.fft.fftfreq(5, 1 / 5))] * 2)
    store["freqs2"] = f2
    store["mask2"] = np.asarray(rc.FCC._calculate_mask(f2))
    store["flat2"] = np.asarray(rc.FCC._flat_frequencies(f2))

    fp = np.abs(np.asarray(rc.FCC._correlate(mat, method="pearson")))
    store["fp"] = fp
    store["weight_linear"] = np.asarray(rc.FCC._weighting_linear(fp))
    store["weight_mean"] = np.asarray(rc.FCC._weighting_mean(fp, mat.transpose()))
    store["fcc"] = np.asarray(rc.FCC.calculate_fcc(fp))

    coeffs = rng.normal(size=9) + 1j * rng.normal(size=9)
    store["psd_in"] = coeffs
    store["psd"] = np.asarray(rc.Coefficients.get_psd(coeffs))
    # a real-valued series: c_{-k} = conj(c_k), frequencies in fftshift order
    half = rng.normal(size=4) + 1j * rng.normal(size=4)
    series = np.concatenate([np.conj(half[::-1]), [0.7], half])
    xs = np.linspace(-np.pi, np.pi, 11)
    store["series_c"], store["series_f"], store["series_x"] = series, f1, xs
    store["series_y"] = np.asarray([rc.Coefficients.evaluate_Fourier_series(series, f1, float(x))
                                    for x in xs])

    # expressibility
    fid = np.linspace(0.0, 1.0, 13)
    for nq in (1, 2, 4):
        store[f"haar_prob_{nq}"] = np.asarray([rx.Expressibility._haar_probability(float(f), nq)
                                               for f in fid])
        x, y = rx.Expressibility.haar_integral(nq, 20, cache=False)
        store[f"haar_int_x_{nq}"], store[f"haar_int_y_{nq}"] = np.asarray(x), np.asarray(y)
    store["haar_fid"] = fid
    p = rng.uniform(0.01, 1.0, (3, 20))
    p /= p.sum(axis=1, keepdims=True)
    q = rng.uniform(0.01, 1.0, 20)
    q /= q.sum()
    store["kl_p"], store["kl_q"] = p, q
    store["kl"] = np.asarray(rx.Expressibility.kullback_leibler_divergence(p, q))

    # jaqsi helpers and the Meyer-Wallach measure on random states (batched and single)
    import qml_essentials.entanglement as ren
    import qml_essentials.jaqsi as rjs

    n = 3
    psi = rng.normal(size=(4, 2 ** n)) + 1j * rng.normal(size=(4, 2 ** n))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    rhos = np.einsum("bi,bj->bij", psi, psi.conj())
    mixed = 0.7 * rhos + 0.3 * np.eye(2 ** n) / 2 ** n
    store["mw_rhos"], store["mw_mixed"] = rhos, mixed
    store["mw_pure"] = np.asarray(ren.Entanglement._compute_meyer_wallach_meas(rhos, n))
    store["mw_mix"] = np.asarray(ren.Entanglement._compute_meyer_wallach_meas(mixed, n))
    for keep in ([0], [2], [0, 2], [1, 2], [0, 1, 2]):
        tag = "".join(map(str, keep))
        store[f"ptrace_{tag}"] = np.asarray(rjs.partial_trace(rhos, n, keep))
        store[f"ptrace1_{tag}"] = np.asarray(rjs.partial_trace(rhos[1], n, keep))
        probs = np.abs(psi) ** 2
        store[f"marg_{tag}"] = np.asarray(rjs.marginalize_probs(probs, n, tuple(keep)))
        store[f"marg1_{tag}"] = np.asarray(rjs.marginalize_probs(probs[2], n, tuple(keep)))
    store["marg_probs"] = np.abs(psi) ** 2

    # operator algebra (operations.py:112-400): the same expressions on both modules
    sys.path.insert(2, os.path.join(HERE, "..", "tests"))
    import golden_algebra_cases as gac
    import qml_essentials.operations as rop

    for name, fn in gac.CASES.items():
        o = fn(rop)
        store[f"alg_{name}_matrix"] = np.asarray(o.matrix)
        store[f"alg_{name}_wires"] = np.asarray(list(o.wires), dtype=np.int64)

    # the callers themselves, deterministic routes (model.params given, no sampling):
    # Coefficients.get_spectrum (coefficients.py:25-150) and Entanglement.meyer_wallach with
    # n_samples=None (entanglement.py:17-105), through the reference's batched Model call
    import qml_essentials.model as rmodel

    sys.path.insert(2, os.path.join(HERE, "..", "tests"))
    import golden_algebra_cases as gac  # noqa: F811

    for name, (n, L, ct, B_P) in gac.CALLER_MODELS.items():
        m = rmodel.Model(n_qubits=n, n_layers=L, circuit_type=ct)
        m.params = rng.uniform(0, 2 * np.pi, (B_P,) + tuple(np.shape(m.params))[1:])
        store[f"call_{name}_params"] = np.asarray(m.params)
        for tag, kw in gac.SPECTRUM_SETTINGS.items():
            c, f = rc.Coefficients.get_spectrum(m, **kw)
            store[f"call_{name}_{tag}_coeffs"] = np.asarray(c)
            store[f"call_{name}_{tag}_freqs"] = np.asarray(f)
        store[f"call_{name}_mw"] = np.asarray(ren.Entanglement.meyer_wallach(m, n_samples=None))
        # FCC on the given parameter samples (n_samples=0: no re-initialisation,
        # coefficients.py:1100-1160)
        if B_P >= 3:
            for method in gac.FCC_METHODS:
                store[f"call_{name}_fcc_{method}"] = np.asarray(
                    rc.FCC.get_fcc(m, n_samples=0, method=method))
            fp, fr = rc.FCC.get_fourier_fingerprint(m, n_samples=0)
            store[f"call_{name}_fp"], store[f"call_{name}_fp_freqs"] = np.asarray(fp), np.asarray(fr)

    out = os.path.join(HERE, "..", "tests", "golden", "reference_analysis.npz")
    np.savez_compressed(out, **store)
    print(f"wrote {os.path.normpath(out)}: {len(store)} arrays, {os.path.getsize(out)} bytes")
    for k, v in store.items():
        print(" ", k, v.shape, v.dtype)


if __name__ == "__main__":
    main()
