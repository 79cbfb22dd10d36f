"""Parameter-shift gradients: exact derivative of Model outputs with
respect to the variational parameters, against central differences of the ORACLE's
independent restatement (1e-7) and the reference's analytic case d<Z>/dtheta of RX
(tests/test_jaqsi.py:131-141: -sin(theta))."""

import warnings

import numpy as np
import pytest

from oracle import circuits as oc, sim as osim
from qml_essentials_b200 import gradients
from qml_essentials_b200.model import Model


def _oracle_value(n, L, ct, params, x, typ, noise):
    tape = oc.variational_tape(n, L, ct, params, [x], noise_params=noise)
    return np.asarray(osim.simulate_and_measure(tape, n, typ, [("PauliZ", [q], []) for q in range(n)]))


def _central(n, L, ct, params, x, typ, noise, h=1e-5):
    g = np.zeros(params.shape + _oracle_value(n, L, ct, params, x, typ, noise).shape)
    for idx in np.ndindex(*params.shape):
        a, b = params.copy(), params.copy()
        a[idx] += h
        b[idx] -= h
        g[idx] = (_oracle_value(n, L, ct, a, x, typ, noise)
                  - _oracle_value(n, L, ct, b, x, typ, noise)) / (2 * h)
    return g


CASES = [
    (2, 1, "Circuit_19", "expval", None),          # RX, RZ, CRX: both generator spectra
    (3, 1, "Strongly_Entangling", "expval", None),  # Rot: three angles per gate
    (3, 1, "Hardware_Efficient", "probs", None),
    (2, 1, "Circuit_19", "expval", {"Depolarizing": 0.05, "AmplitudeDamping": 0.1}),
]


def _check(n, L, ct, typ, noise):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = Model(n, L, ct)
        p = np.random.default_rng(4).uniform(0, 2 * np.pi, m._params_shape)
        got = gradients.param_shift(m, p, inputs=np.array([[0.37]]), execution_type=typ,
                                    noise_params=dict(noise) if noise else None)
    want = _central(n, L, ct, p, 0.37, typ, noise)
    if typ == "probs":
        want = want.reshape(got.shape)
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 1e-7


@pytest.mark.parametrize("n,L,ct,typ,noise", CASES)
def test_param_shift_matches_oracle_central_differences(n, L, ct, typ, noise):
    _check(n, L, ct, typ, noise)


def test_param_shift_batched_inputs_and_validation():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = Model(2, 1, "Circuit_19")
        p = np.random.default_rng(1).uniform(0, 2 * np.pi, m._params_shape)
        x = np.array([[0.1], [0.9], [-0.4]])
        g = gradients.param_shift(m, p, inputs=x)
        assert g.shape == (*m._params_shape, 3, 2)
        for i in range(3):
            gi = gradients.param_shift(m, p, inputs=x[i:i + 1])
            assert np.abs(g[..., i, :] - gi).max() < 1e-12
        with pytest.raises(ValueError, match="expval"):
            gradients.param_shift(m, p, inputs=x, execution_type="density")
        with pytest.raises(ValueError, match="GateError"):
            gradients.param_shift(m, p, inputs=x, noise_params={"GateError": 0.1})


@pytest.mark.gpu
@pytest.mark.parametrize("n,L,ct,typ,noise", CASES + [
    (6, 1, "Circuit_19", "expval", None),   # frame engine
    (4, 4, "Hardware_Efficient", "expval", None),  # config 2 circuit: 60 parameters
])
def test_param_shift_on_device(n, L, ct, typ, noise):
    _check(n, L, ct, typ, noise)
