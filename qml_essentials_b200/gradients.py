"""Gradients of ``Model`` outputs with respect to the variational parameters.

The reference differentiates through its simulator with ``jax.grad``
(script.py:221-235,463-467; tests/test_jaqsi.py:131-141,764-786; tests/test_model.py:1082-1147)
- reverse-mode autodiff of the einsum chain.  This backend has no tracer; it uses the
PARAMETER-SHIFT rule instead, which is exact (not a finite difference) for every
parametrised gate of the reference's gate set and maps onto what the kernels are good at:
one batched launch over shifted parameter sets.

Every ansatz parameter enters exactly one gate, ``exp(-i theta G)`` with generator spectrum
``{+-1/2}`` (RX / RY / RZ / Rot angles / PauliRot / RXX ...) or ``{0, +-1/2}`` (CRX / CRY /
CRZ / controlled Pauli rotations).  The four-term rule

    dE/dtheta = c+ [E(theta + pi/2) - E(theta - pi/2)] - c- [E(theta + 3pi/2) - E(theta - 3pi/2)],
    c+- = (sqrt(2) +- 1) / (4 sqrt(2))

is exact for both spectra (Anselmetti et al. 2021; Wierichs et al. 2022: the expectation
value is a trigonometric polynomial with frequencies 1/2 and 1), including noisy circuits -
channels are linear maps - and GateError-free runs.  Cost: ``4 * n_params`` circuit
evaluations per gradient, evaluated as ONE batch by the same kernels as the forward pass.

Not covered: gradients with respect to ``inputs`` (data re-uploading feeds one input into
several gates, for which the rule above does not hold) and pulse parameters.
"""

from __future__ import annotations

from typing import Any, Optional

import numpy as np

_SHIFTS = np.array([np.pi / 2, -np.pi / 2, 3 * np.pi / 2, -3 * np.pi / 2])
_CP = (np.sqrt(2.0) + 1.0) / (4.0 * np.sqrt(2.0))
_CM = (np.sqrt(2.0) - 1.0) / (4.0 * np.sqrt(2.0))
_WEIGHTS = np.array([_CP, -_CP, -_CM, _CM])


def param_shift(model, params: Optional[np.ndarray] = None, inputs: Optional[np.ndarray] = None,
                **call_kwargs: Any) -> np.ndarray:
    """``d model(params, inputs) / d params`` by the four-term parameter-shift rule.

    ``params``: one parameter set ``(L', P)`` / ``(1, L', P)`` (default: ``model.params``).
    Returns an array of shape ``(L', P, *out)`` where ``out`` is the shape of
    ``model(params=params, inputs=inputs, **call_kwargs)`` (the input batch axis, if any,
    comes first in ``out``).  ``execution_type`` may be ``expval`` or ``probs``.
    """
    p = np.asarray(model.params if params is None else params, dtype=np.float64)
    if p.ndim == 3:
        if p.shape[0] != 1:
            raise ValueError("param_shift differentiates one parameter set at a time")
        p = p[0]
    if tuple(p.shape) != tuple(model._params_shape):
        raise ValueError(f"params must have shape {tuple(model._params_shape)}, got {p.shape}")
    et = call_kwargs.get("execution_type", model.execution_type)
    if et not in ("expval", "probs"):
        raise ValueError("parameter-shift gradients are defined for 'expval' and 'probs'")
    noise = call_kwargs.get("noise_params", model.noise_params)
    if noise and noise.get("GateError"):
        raise ValueError("GateError draws fresh angles per evaluation; its gradient is undefined")

    n_par = p.size
    shifted = np.repeat(p.reshape(1, -1), 4 * n_par, axis=0)
    idx = np.arange(n_par)
    for s in range(4):
        shifted[4 * idx + s, idx] += _SHIFTS[s]
    shifted = shifted.reshape(4 * n_par, *p.shape)

    base = np.asarray(model(params=p[None], inputs=inputs, **call_kwargs))
    vals = np.asarray(model(params=shifted, inputs=inputs, **call_kwargs))
    # the parameter-batch axis of `vals`: position of the axis that `base` lacks
    b_in = 1 if inputs is None else int(np.asarray(inputs).reshape(
        -1, model.n_input_feat).shape[0])
    axis = 1 if (b_in > 1 and vals.ndim > base.ndim and vals.shape[0] == b_in) else 0
    vals = np.moveaxis(vals, axis, 0).reshape(n_par, 4, *base.shape)
    grad = np.tensordot(_WEIGHTS, vals, axes=([0], [1]))
    return grad.reshape(*p.shape, *base.shape)
