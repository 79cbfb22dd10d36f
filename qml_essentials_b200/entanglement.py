"""Entangling capability measures.  Mirror of the state-evolution based parts of the
reference's ``qml_essentials/entanglement.py``: ``meyer_wallach`` (lines 17-103) and
``bell_measurements`` (106-219).

Meyer-Wallach needs, per sample and qubit, the purity of the state with that qubit
traced out.  The reference pulls B full density matrices to the host and loops
``partial_trace`` over the qubits (entanglement.py:86-101).  Here the states stay in
HBM and ``qmlb_purity`` reduces them to a (B, n) table on the GPU; for noise-free
circuits the 2^n statevector is used instead of the 4^n density matrix (same
purities by the Schmidt decomposition).

Multi-GPU: samples are sharded across ranks; ``sum`` and ``sum of squares`` of the
per-sample measure cross ranks in one all-reduce.

The entropy / eigen-decomposition based measures (relative entropy, entanglement of
formation, entanglement.py:222-469) are host linear algebra on top of scipy and are
outside the backend's scope.
"""

from __future__ import annotations

import logging
from typing import Any, Optional

import numpy as np

from . import jaqsi as js
from . import operations as op
from . import parallel, rng
from .model import Model

log = logging.getLogger(__name__)


def _host(x) -> np.ndarray:
    """Device tensor (or test double) -> float64 host array."""
    x = x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)
    return np.asarray(x, dtype=np.float64)


class Entanglement:
    @classmethod
    def meyer_wallach(cls, model: Model, n_samples: Optional[int], random_key=None,
                      scale: bool = False, **kwargs: Any) -> float:
        """Mean Meyer-Wallach measure over ``n_samples`` random parameter sets (or the
        model's current parameters when ``n_samples`` is None or <= 0)."""
        if "noise_params" in kwargs:
            log.warning("Meyer-Wallach measure not suitable for noisy circuits. "
                        "Consider 'concentratable entanglement' instead.")
        if scale:
            n_samples = int(2 ** model.n_qubits * n_samples)
        if n_samples is not None and n_samples > 0:
            model.initialize_params(random_key, repeat=n_samples)
        kwargs.setdefault("inputs", None)

        params = model.params
        lo, hi = parallel.shard_bounds(params.shape[0])
        if parallel.world()[1] > 1 and params.shape[0] > 1:
            params = params[lo:hi]
        ent = cls._meyer_wallach_samples(model, params, kwargs)
        stats = parallel.allreduce_sum(
            np.array([ent.sum(), (ent ** 2).sum(), float(ent.size)], dtype=np.float64))
        mean = stats[0] / stats[2]
        log.debug(f"Variance of measure: {stats[1] / stats[2] - mean ** 2}")
        return float(mean)

    @classmethod
    def _meyer_wallach_samples(cls, model: Model, params, kwargs) -> np.ndarray:
        """Per-sample measure 2 (1 - mean_q Tr[(Tr_q rho)^2]) with the reduction on the
        GPU (entanglement.py:69-103)."""
        from .script import get_executor

        if params.shape[0] == 0:
            return np.zeros(0)
        n = model.n_qubits
        noise = kwargs.get("noise_params", model.noise_params)
        noisy = bool(noise) and any(v is not None and v > 0 for k, v in noise.items()
                                    if k != "GateError")
        saved = model.output_qubit
        model.output_qubit = -1  # the measure is defined on the full register
        try:
            if noisy:
                st = model.device_result(params=params, execution_type="density", **kwargs)
                st = st.reshape(-1, 2 ** n, 2 ** n)
            else:
                st = model.device_result(params=params, execution_type="state", **kwargs)
                st = st.reshape(-1, 2 ** n)
        finally:
            model.output_qubit = saved
        pur = _host(get_executor().purities(st, n, noisy))
        return 2.0 * (1.0 - pur.mean(axis=1))

    @classmethod
    def _compute_meyer_wallach_meas(cls, rhos: np.ndarray, n_qubits: int) -> np.ndarray:
        """Host restatement for given density matrices (entanglement.py:69-103); the
        product path uses the device reduction above."""
        rhos = np.asarray(rhos).reshape(-1, 2 ** n_qubits, 2 ** n_qubits)
        qb = list(range(n_qubits))
        purity = np.zeros(rhos.shape[0])
        for j in range(n_qubits):
            red = js.partial_trace(rhos, n_qubits, qb[:j] + qb[j + 1:])
            purity += np.trace((red @ red).real, axis1=-2, axis2=-1)
        return 2 * (1 - purity / n_qubits)

    @classmethod
    def bell_measurements(cls, model: Model, n_samples: int, random_key=None,
                          scale: bool = False, **kwargs: Any) -> float:
        """Bell-measurement estimate of the Meyer-Wallach measure on a 2n-qubit circuit
        holding two copies of the model state (entanglement.py:106-219)."""
        if "noise_params" in kwargs:
            log.warning("Bell Measurements not suitable for noisy circuits. "
                        "Consider 'concentratable entanglement' instead.")
        if scale:
            n_samples = int(2 ** model.n_qubits * n_samples)
        n = model.n_qubits

        def _bell_circuit(params, inputs, pulse_params=None, random_key=None, **kw):
            from .tape import copy_to_tape

            def vari():
                model._variational(params, inputs, pulse_params=pulse_params,
                                   random_key=random_key, **kw)

            vari()                       # first copy on wires 0..n-1
            copy_to_tape(vari, offset=n)  # second copy on wires n..2n-1
            for q in range(n):
                op.CX(wires=[q, q + n])
                op.H(wires=q)

        bell = js.Script(f=_bell_circuit, n_qubits=2 * n)
        if n_samples is not None and n_samples > 0:
            model.initialize_params(random_key, repeat=n_samples)
            params = model.params
        else:
            params = model.params
            if params.ndim <= 2:
                params = params.reshape(1, *params.shape)
        total = params.shape[0]
        lo, hi = parallel.shard_bounds(total)
        if parallel.world()[1] > 1 and total > 1:
            params = params[lo:hi]
        inputs = model._inputs_validation(kwargs.pop("inputs", None))
        key = random_key if random_key is not None else model.random_key

        if params.shape[0] > 1:
            keys = rng.split(key, params.shape[0])
            probs = bell.execute(type="probs", args=(params, inputs, model.pulse_params, keys),
                                 kwargs=kwargs, in_axes=(0, None, None, 0))
        elif params.shape[0] == 1:
            probs = bell.execute(type="probs", args=(params[0], inputs, model.pulse_params, key),
                                 kwargs=kwargs)[None]
        else:
            probs = np.zeros((0, 4 ** n))
        # P(|11>) of each (q, q + n) pair
        p11 = np.stack([js.marginalize_probs(probs, 2 * n, [q, q + n])[..., -1]
                        for q in range(n)], axis=-1)
        exp = 1 - 2 * p11  # (samples, n)
        stats = parallel.allreduce_sum(
            np.concatenate([exp.sum(axis=0), [float(exp.shape[0])]]).astype(np.float64))
        measure = 2 * (1 - stats[:n] / stats[n])
        return min(max(float(measure.mean()), 0.0), 1.0)

    @classmethod
    def concentratable_entanglement(cls, model: Model, n_samples: int, random_key=None,
                                    scale: bool = False, **kwargs: Any) -> float:
        """Concentratable entanglement (arXiv:2104.06923) by a swap test on a 3n-qubit
        circuit: ancillas on wires 0..n-1, two copies of the model state on the other two
        registers, H - CSWAP - H, ``1 - P(ancillas = 0..0)`` (entanglement.py:471-577).
        The 3n-qubit circuit goes through the same ``Script`` -> CUDA path as every other
        circuit (CSWAP = a 3-bit permutation)."""
        n = model.n_qubits
        if scale:
            n_samples = int(2 ** n * n_samples)

        def _swap_test_circuit(params, inputs, pulse_params=None, random_key=None, **kw):
            from .tape import copy_to_tape

            def vari():
                model._variational(params, inputs, pulse_params=pulse_params,
                                   random_key=random_key, **kw)

            copy_to_tape(vari, offset=n)
            copy_to_tape(vari, offset=2 * n)
            for i in range(n):
                op.H(wires=i)
            for i in range(n):
                op.CSWAP(wires=[i, i + n, i + 2 * n])
            for i in range(n):
                op.H(wires=i)

        swap = js.Script(f=_swap_test_circuit, n_qubits=3 * n)
        if n_samples is not None and n_samples > 0:
            model.initialize_params(random_key, repeat=n_samples)
            params = model.params
        else:
            params = model.params
            if params.ndim <= 2:
                params = params.reshape(1, *params.shape)
        total = params.shape[0]
        lo, hi = parallel.shard_bounds(total)
        if parallel.world()[1] > 1 and total > 1:
            params = params[lo:hi]
        inputs = model._inputs_validation(kwargs.pop("inputs", None))
        key = random_key if random_key is not None else model.random_key

        if params.shape[0] > 1:
            keys = rng.split(key, params.shape[0])
            probs = swap.execute(type="probs", args=(params, inputs, model.pulse_params, keys),
                                 kwargs=kwargs, in_axes=(0, None, None, 0))
        elif params.shape[0] == 1:
            probs = swap.execute(type="probs", args=(params[0], inputs, model.pulse_params, key),
                                 kwargs=kwargs)[None]
        else:
            probs = np.zeros((0, 8 ** n))
        anc = js.marginalize_probs(probs, 3 * n, tuple(range(n)))
        ent = 1 - np.asarray(anc)[..., 0]
        stats = parallel.allreduce_sum(
            np.array([ent.sum(), float(ent.shape[0])], dtype=np.float64))
        return float(stats[0] / stats[1])

    @classmethod
    def relative_entropy(cls, *a, **k):
        raise NotImplementedError("host scipy.linalg.logm analysis: outside the backend scope")

    @classmethod
    def entanglement_of_formation(cls, *a, **k):
        raise NotImplementedError("host eigen-decomposition analysis: outside the backend scope")
