"""The oracle against outputs of the reference's OWN source files.

`tests/golden/reference_sim.npz` was written by `tools/gen_golden_reference.py`, which runs
the unmodified `/root/reference/qml_essentials/{operations,simulation,tape}.py` in the build
container on a NumPy stand-in for the slice of the JAX API they use (JAX is not installed;
`tools/jax_numpy_shim`).  It holds, for seeded random inputs, what the reference's gate /
channel classes return as matrices and Kraus sets and what its `simulate_pure`,
`simulate_mixed`, `measure_state`, `measure_density` and `simulate_and_measure` return for
random circuits.  Here the oracle (`oracle/gates.py`, `oracle/sim.py`) replays the same tape
entries: 1e-12 absolute, i.e. inside the 1e-10 the north star asks of the CUDA path - and the
`-m gpu` suite holds the CUDA path to this oracle.  Nothing here reads `/root/reference`.
"""
import json
import os

import numpy as np
import pytest

from oracle import gates as G
from oracle import sim as osim

TOL = 1e-12
PATH = os.path.join(os.path.dirname(__file__), "golden", "reference_sim.npz")


@pytest.fixture(scope="module")
def golden():
    z = np.load(PATH)
    index = json.loads(bytes(z["index_json"]).decode())
    return z, index


def _entry(rec, z):
    extra = rec["extra"]
    if extra is not None:
        if "str" in extra:
            extra = extra["str"]
        elif "tuple" in extra:
            extra = (extra["tuple"][0], extra["tuple"][1])
        else:
            extra = z[extra["array"]]
            if rec["name"] == "QubitChannel":
                extra = list(extra)
    return (rec["name"], rec["wires"], rec["params"], extra)


def test_fixture_covers_every_gate_and_channel_of_the_path(golden):
    z, index = golden
    names = {r["entry"]["name"] for r in index if r["kind"] == "single"}
    assert names >= {"Id", "PauliX", "PauliY", "PauliZ", "H", "S", "SWAP", "CX", "CY", "CZ", "CCX",
                     "CSWAP", "RX", "RY", "RZ", "CRX", "CRY", "CRZ", "RXX", "RYY", "RZZ", "RZX",
                     "ControlledPhaseShift", "Rot", "PauliRot", "ControlledPauliRot",
                     "DiagonalQubitUnitary", "BitFlip", "PhaseFlip", "DepolarizingChannel",
                     "AmplitudeDamping", "PhaseDamping", "ThermalRelaxationError", "QubitChannel"}
    circuits = [r for r in index if r["kind"] == "circuit"]
    assert sum(1 for c in circuits if c["noisy"]) >= 8 and sum(1 for c in circuits if not c["noisy"]) >= 10


def test_gate_matrices_and_kraus_sets_equal_the_reference(golden):
    """operations.py:719-1929 as executed, against oracle/gates.py."""
    z, index = golden
    worst = 0.0
    for r in index:
        if r["kind"] != "single":
            continue
        name, wires, params, extra = _entry(r["entry"], z)
        key = f"single{r['id']}"
        if G.is_channel(name):
            want = z[key + "_kraus"]
            got = np.stack(G.kraus_matrices(name, params, extra))
            assert got.shape == want.shape, name
            # a Kraus set is defined up to order only where the reference fixes one: same order
        else:
            want = z[key + "_matrix"]
            got = G.unitary_matrix(name, wires, params, extra)
        err = float(np.abs(got - want).max())
        assert err < TOL, (name, params, err)
        worst = max(worst, err)
    assert worst < TOL


def test_thermal_relaxation_fixture_has_both_regimes(golden):
    """t2 <= t1 (six Kraus matrices, operations.py:1854-1876) and t2 > t1 (Choi route,
    operations.py:1877-1895) are both in the fixture."""
    z, index = golden
    counts = set()
    for r in index:
        if r["kind"] == "single" and r["entry"]["name"] == "ThermalRelaxationError":
            _, t1, t2, _ = r["entry"]["params"]
            counts.add(t2 <= t1)
    for c in (r for r in index if r["kind"] == "circuit"):
        for e in c["tape"]:
            if e["name"] == "ThermalRelaxationError":
                counts.add(e["params"][2] <= e["params"][1])
    assert counts == {True, False}


def test_circuits_equal_the_reference_simulator(golden):
    """simulation.py:65-128 (evolution) and :204-317 (measurement) as executed."""
    z, index = golden
    for c in index:
        if c["kind"] != "circuit":
            continue
        n, tag = c["n"], f"case{c['id']}"
        tape = [_entry(e, z) for e in c["tape"]]
        obs = [_entry(o, z) for o in c["obs"]]
        zobs = [o for o in obs if o[0] == "PauliZ"]
        if c["noisy"]:
            rho = osim.simulate_mixed(tape, n)
            assert np.abs(rho - z[tag + "_density"]).max() < TOL
            assert np.abs(osim.measure_density(rho, n, "probs", obs) - z[tag + "_probs"]).max() < TOL
            assert np.abs(osim.measure_density(rho, n, "expval", obs) - z[tag + "_expval"]).max() < 10 * TOL
            assert np.abs(osim.measure_density(rho, n, "expval", zobs) - z[tag + "_expval_z"]).max() < TOL
            # routing (simulation.py:42-57, 176-194): a tape with a channel is a density run
            assert osim.uses_density(tape, "expval")
            out = osim.simulate_and_measure(tape, n, "expval", obs)
            assert np.abs(out - z[tag + "_expval"]).max() < 10 * TOL
        else:
            psi = osim.simulate_pure(tape, n)
            assert np.abs(psi - z[tag + "_state"]).max() < TOL
            assert np.abs(osim.measure_state(psi, n, "probs", obs) - z[tag + "_probs"]).max() < TOL
            assert np.abs(osim.measure_state(psi, n, "expval", obs) - z[tag + "_expval"]).max() < 10 * TOL
            assert np.abs(osim.measure_state(psi, n, "expval", zobs) - z[tag + "_expval_z"]).max() < TOL
            dens = osim.simulate_and_measure(tape, n, "density", obs)
            assert np.abs(dens - z[tag + "_density"]).max() < TOL
            assert np.abs(osim.simulate_mixed(tape, n) - z[tag + "_density_mixed"]).max() < TOL


def test_reference_outputs_are_physical(golden):
    """Sanity of the fixture itself: unit norm, unit trace, Hermitian, probabilities sum to 1."""
    z, index = golden
    for c in index:
        if c["kind"] != "circuit":
            continue
        tag = f"case{c['id']}"
        rho = z[tag + "_density"]
        assert abs(np.trace(rho) - 1) < 1e-12 and np.abs(rho - rho.conj().T).max() < 1e-12
        assert abs(z[tag + "_probs"].sum() - 1) < 1e-12
        if not c["noisy"]:
            assert abs(np.vdot(z[tag + "_state"], z[tag + "_state"]) - 1) < 1e-12


# ---- Model.__call__ of the reference (tools/gen_golden_reference_model.py) -----------------
MODEL_PATH = os.path.join(os.path.dirname(__file__), "golden", "reference_model.npz")


@pytest.fixture(scope="module")
def golden_model():
    z = np.load(MODEL_PATH)
    return z, json.loads(bytes(z["index_json"]).decode())


def _oracle_model_rows(c, params, inputs):
    from oracle import circuits as oc

    n, rows = c["n"], []
    obs = [("PauliZ", [q], [], None) for q in range(n)]
    for i in range(c["B_I"]):
        for p in range(c["B_P"]):
            tape = oc.variational_tape(n, c["L"], c["ct"], params[p], [inputs[i, 0]],
                                       noise_params=c["noise"])
            rows.append(np.asarray(osim.simulate_and_measure(tape, n, c["typ"], obs)).reshape(-1))
    return np.stack(rows)


def test_model_fixture_covers_baseline_configs_and_every_ansatz(golden_model):
    z, index = golden_model
    from oracle import circuits as oc  # noqa: F401

    names = {c["ct"] for c in index}
    assert len(names) >= 23 and {"Circuit_19", "Hardware_Efficient", "Circuit_15",
                                 "Strongly_Entangling", "GHZ", "No_Ansatz"} <= names
    assert {c["typ"] for c in index} == {"expval", "probs", "state", "density"}
    assert any(c["noise"] for c in index)
    opts = [c["kw"] for c in index if c.get("kw")]
    assert {"binary", "ternary", "golomb"} <= {k["encoding"]["strategy"] for k in opts
                                                if isinstance(k.get("encoding"), dict)}
    assert any("output_qubit" in k for k in opts) and any("data_reupload" in k for k in opts)


def test_model_calls_equal_the_reference(golden_model):
    """model.py:746-1064, 1512-1737 + ansaetze.py + unitary.py noise insertion as executed by
    the reference, against oracle/circuits.py + oracle/sim.py: parameter layout, ansatz
    structure, encoding and re-uploading, noise channel placement, output shapes."""
    z, index = golden_model
    worst = 0.0
    for c in index:
        if c.get("kw") or c.get("api_only"):
            continue  # constructor options: replayed through the drop-in API below
        tag = f"model{c['id']}"
        params, inputs, want = z[tag + "_params"], z[tag + "_inputs"], z[tag + "_out"]
        got = _oracle_model_rows(c, params, inputs)
        assert got.shape == want.shape, (c, got.shape, want.shape)
        err = float(np.abs(got - want).max())
        assert err < 1e-11, (c["ct"], c["typ"], c["n"], err)
        worst = max(worst, err)
    assert worst < 1e-11


def _api_model_rows(c, params, inputs, precision="complex128"):
    """The same cases through the drop-in `Model.__call__` (batched over inputs x parameter
    sets, flat order b = i * B_P + p as `model.py:1414-1483` assimilates them)."""
    import warnings

    from qml_essentials_b200.ansaetze import Encoding
    from qml_essentials_b200.model import Model

    kw = dict(c.get("kw") or {})
    if isinstance(kw.get("encoding"), dict):
        kw["encoding"] = Encoding(kw["encoding"]["strategy"], kw["encoding"]["gates"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = Model(n_qubits=c["n"], n_layers=c["L"], circuit_type=c["ct"], precision=precision,
                  **kw)
        assert tuple(m._params_shape) == tuple(c["params_shape"]), (c["ct"], m._params_shape)
        out = np.asarray(m(params=params, inputs=inputs, execution_type=c["typ"],
                           noise_params=dict(c["noise"]) if c["noise"] else None))
    return out


def _check_api_against_fixture(golden_model, tol, precision="complex128", options=None):
    """options: None = every case, False = the plain cases, True = the constructor options."""
    z, index = golden_model
    worst = 0.0
    for c in index:
        if options is not None and bool(c.get("kw") or c.get("api_only")) != options:
            continue
        tag = f"model{c['id']}"
        got = _api_model_rows(c, z[tag + "_params"], z[tag + "_inputs"], precision)
        # ONE batched call of the drop-in against ONE batched call of the reference: same
        # result array, same shape (squeezed axes included)
        want = z[tag + "_batched"]
        assert got.shape == want.shape, (c, got.shape, want.shape)
        err = float(np.abs(got - want).max())
        assert err < tol, (c["ct"], c["typ"], c["n"], c.get("kw"), err)
        worst = max(worst, err)
    return worst


def test_drop_in_model_equals_the_reference_through_the_interpreter(golden_model):
    """Host logic of the drop-in (recording, tape -> program compiler, batch factors, encodings,
    output-qubit post-processing, output shapes) on the CPU program interpreter, against the
    reference's own results: all 73 cases."""
    assert _check_api_against_fixture(golden_model, 1e-10) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("complex128", 1e-10), ("complex64", 1e-5)])
def test_cuda_model_equals_the_reference(golden_model, precision, tol):
    """The CUDA path (ctypes -> libqmlb200.so) against results of the reference's own
    `Model.__call__` on identical parameters, inputs and noise, at the north star's tolerance:
    the BASELINE configurations (reduced), every ansatz, noisy density / probs / expval."""
    assert _check_api_against_fixture(golden_model, tol, precision, options=False) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("complex128", 1e-10), ("complex64", 1e-5)])
def test_cuda_model_options_equal_the_reference(golden_model, precision, tol):
    """Same, for the constructor options: two input features, RY / binary / ternary / golomb
    encodings, no re-uploading, output-qubit subsets (expval, marginal probs, partial trace),
    state preparation, and all nine noise keys at once (thermal relaxation in both regimes)."""
    assert _check_api_against_fixture(golden_model, tol, precision, options=True) < tol


# ---- analysis helpers of the reference (tools/gen_golden_reference_analysis.py) ------------
ANALYSIS_PATH = os.path.join(os.path.dirname(__file__), "golden", "reference_analysis.npz")


def _close(a, b, tol=1e-12):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    return bool(np.all(np.abs(np.nan_to_num(a) - np.nan_to_num(b)) < tol))


def test_fcc_statistics_equal_the_reference():
    """coefficients.py:1165-1650 as executed (`FCC._correlate` with its four methods, with and
    without missing values, masks, flat frequency labels, weightings, `calculate_fcc`)
    against the drop-in's host versions."""
    import warnings

    from qml_essentials_b200.coefficients import FCC

    z = np.load(ANALYSIS_PATH)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for method in ("pearson", "complex_pearson", "spearman", "covariance"):
            assert _close(FCC._correlate(z["corr_mat"], method=method), z[f"corr_{method}"]), method
            assert _close(FCC._correlate(z["corr_mat_holes"], method=method),
                          z[f"corr_holes_{method}"]), method
        assert np.array_equal(np.asarray(FCC._calculate_mask(z["freqs1"])), z["mask1"])
        assert np.array_equal(np.asarray(FCC._calculate_mask(z["freqs2"])), z["mask2"])
        assert _close(FCC._flat_frequencies(z["freqs1"]), z["flat1"])
        assert _close(FCC._flat_frequencies(z["freqs2"]), z["flat2"])
        assert _close(FCC._weighting_linear(z["fp"]), z["weight_linear"])
        assert _close(FCC._weighting_mean(z["fp"], z["corr_mat"].transpose()), z["weight_mean"])
        assert abs(float(FCC.calculate_fcc(z["fp"])) - float(z["fcc"])) < 1e-14


def test_spectrum_helpers_equal_the_reference():
    """coefficients.py:152-238 (`get_psd`, `evaluate_Fourier_series`)."""
    from qml_essentials_b200.coefficients import Coefficients

    z = np.load(ANALYSIS_PATH)
    assert _close(Coefficients.get_psd(z["psd_in"]), z["psd"])
    ys = [Coefficients.evaluate_Fourier_series(z["series_c"], z["series_f"], float(x))
          for x in z["series_x"]]
    assert _close(np.real(np.asarray(ys)).reshape(-1), np.real(z["series_y"]).reshape(-1), 1e-11)


def test_expressibility_helpers_equal_the_reference():
    """expressibility.py (`_haar_probability`, `haar_integral`, `kullback_leibler_divergence`)."""
    from qml_essentials_b200.expressibility import Expressibility

    z = np.load(ANALYSIS_PATH)
    for nq in (1, 2, 4):
        got = [Expressibility._haar_probability(float(f), nq) for f in z["haar_fid"]]
        assert _close(got, z[f"haar_prob_{nq}"])
        x, y = Expressibility.haar_integral(nq, 20, cache=False)
        assert _close(x, z[f"haar_int_x_{nq}"]) and _close(y, z[f"haar_int_y_{nq}"], 1e-10)
    assert _close(Expressibility.kullback_leibler_divergence(z["kl_p"], z["kl_q"]), z["kl"])


def test_jaqsi_helpers_and_meyer_wallach_equal_the_reference():
    """jaqsi.py:79-160 (`partial_trace`, `marginalize_probs`, batched and single) and
    entanglement.py:69-105 (`_compute_meyer_wallach_meas`) as executed."""
    from qml_essentials_b200 import jaqsi as js
    from qml_essentials_b200.entanglement import Entanglement

    z = np.load(ANALYSIS_PATH)
    n = 3
    assert _close(Entanglement._compute_meyer_wallach_meas(z["mw_rhos"], n), z["mw_pure"])
    assert _close(Entanglement._compute_meyer_wallach_meas(z["mw_mixed"], n), z["mw_mix"])
    for keep in ([0], [2], [0, 2], [1, 2], [0, 1, 2]):
        tag = "".join(map(str, keep))
        assert _close(js.partial_trace(z["mw_rhos"], n, keep), z[f"ptrace_{tag}"])
        assert _close(js.partial_trace(z["mw_rhos"][1], n, keep), z[f"ptrace1_{tag}"])
        assert _close(js.marginalize_probs(z["marg_probs"], n, tuple(keep)), z[f"marg_{tag}"])
        assert _close(js.marginalize_probs(z["marg_probs"][2], n, tuple(keep)), z[f"marg1_{tag}"])


def test_operator_algebra_equals_the_reference():
    """operations.py:112-400 (dagger / power / scalar and operator products / sums): the same
    expressions (`tests/golden_algebra_cases.py`) on the drop-in's `operations` module give
    the matrices and wires the reference's module gave."""
    import golden_algebra_cases as gac

    from qml_essentials_b200 import operations as op

    z = np.load(ANALYSIS_PATH)
    for name, fn in gac.CASES.items():
        o = fn(op)
        assert list(o.wires) == [int(w) for w in z[f"alg_{name}_wires"]], name
        assert _close(np.asarray(o.matrix), z[f"alg_{name}_matrix"]), name


def _check_callers(tol, fcc=True):
    """`Coefficients.get_spectrum` (coefficients.py:25-150: grid, batched model call, FFT,
    trim / shift) and `Entanglement.meyer_wallach(n_samples=None)` (entanglement.py:17-105)
    of the reference, run end to end on given parameters, against the drop-in's callers (host
    route on the CPU interpreter; the GPU suite holds the device route to the host route)."""
    import warnings

    import golden_algebra_cases as gac

    from qml_essentials_b200.coefficients import Coefficients
    from qml_essentials_b200.entanglement import Entanglement
    from qml_essentials_b200.model import Model

    z = np.load(ANALYSIS_PATH)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name, (n, L, ct, B_P) in gac.CALLER_MODELS.items():
            m = Model(n_qubits=n, n_layers=L, circuit_type=ct)
            m.params = z[f"call_{name}_params"]
            for tag, kw in gac.SPECTRUM_SETTINGS.items():
                c, f = Coefficients.get_spectrum(m, **kw)
                assert _close(np.asarray(c), z[f"call_{name}_{tag}_coeffs"], tol), (name, tag)
                assert _close(np.asarray(f), z[f"call_{name}_{tag}_freqs"]), (name, tag)
            mw = Entanglement.meyer_wallach(m, n_samples=None)
            assert abs(float(mw) - float(z[f"call_{name}_mw"])) < tol, name
            if fcc and B_P >= 3:  # FCC on the given samples (n_samples=0), every correlation method
                from qml_essentials_b200.coefficients import FCC

                for method in gac.FCC_METHODS:
                    got = float(FCC.get_fcc(m, n_samples=0, method=method))
                    assert abs(got - float(z[f"call_{name}_fcc_{method}"])) < 10 * tol, (name, method)
                fp, fr = FCC.get_fourier_fingerprint(m, n_samples=0)
                assert _close(np.asarray(fp), z[f"call_{name}_fp"], 10 * tol), name
                assert _close(np.asarray(fr), z[f"call_{name}_fp_freqs"]), name


def test_spectrum_fcc_and_meyer_wallach_callers_equal_the_reference():
    """Host routes of the callers on the CPU program interpreter."""
    _check_callers(1e-10)


@pytest.mark.gpu
def test_cuda_callers_equal_the_reference():
    """`get_spectrum` (circuit kernels + device grid DFT with the trim / shift folded in) and
    `meyer_wallach` (device purities) on the CUDA library against the reference's end-to-end
    outputs.  The FCC's device-moments route is held to the host route by
    `tests/test_gpu_analysis.py`; with the handful of samples of this fixture the two routes
    may disagree on which near-dead coefficients count as zero-variance, so it is not
    replayed here."""
    _check_callers(1e-9, fcc=False)
