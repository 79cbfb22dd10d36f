"""Golden vectors from the reference's OWN `Model.__call__`, run in this container.

    python tools/gen_golden_reference_model.py   # writes tests/golden/reference_model.npz

Companion of `tools/gen_golden_reference.py` (read its header first).  Here the unmodified
`/root/reference/qml_essentials/model.py` (+ `ansaetze.py`, `gates.py`, `unitary.py`,
`topologies.py`, `script.py`, `simulation.py`, `operations.py`, `tape.py`) is imported on
the NumPy stand-in for JAX, with the packages it never reaches on this path replaced by
inert placeholders (`tools/jax_numpy_shim/stub_missing.py`).  Every case is evaluated twice:
one call of the reference's `Model.__call__` per (input, parameter set) - its un-batched
route (`script.py:205-219`: record the tape, `simulate_and_measure`) - and ONE call with the
whole batch - its vmapped route (`script.py:399-553`; `jax.vmap` is a loop here); the two must
agree.  Recorded per case: constructor arguments, parameters, inputs, noise parameters, the
per-sample results (`_out`, flat order b = i * B_P + p) and the batched result (`_batched`,
the reference's own output shape).
`tests/test_reference_golden.py` replays them through `oracle/circuits.py` + `oracle/sim.py`.
Not recorded: anything that draws random numbers inside the reference (GateError, shots,
initial parameters) - the stand-in's draws are not jax.random's.
"""
import json
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "jax_numpy_shim"))
sys.path.insert(1, "/root/reference")
import stub_missing  # noqa: E402

stub_missing.install()

import numpy as np  # noqa: E402

import qml_essentials.model as rmodel  # noqa: E402
from qml_essentials.ansaetze import Ansaetze  # noqa: E402

assert rmodel.__file__.startswith("/root/reference/"), rmodel.__file__

NOISE_A = {"Depolarizing": 0.01, "AmplitudeDamping": 0.02}
NOISE_B = {"BitFlip": 0.03, "PhaseFlip": 0.02, "PhaseDamping": 0.04, "Depolarizing": 0.01,
           "MultiQubitDepolarizing": 0.02}


def cases():
    out = [
        # the five BASELINE configs at sizes the reference finishes in seconds
        dict(n=2, L=1, ct="Circuit_19", typ="expval", B_I=9, B_P=1, noise=None),
        dict(n=4, L=4, ct="Hardware_Efficient", typ="expval", B_I=7, B_P=3, noise=None),
        dict(n=6, L=3, ct="Circuit_15", typ="state", B_I=1, B_P=3, noise=None),
        dict(n=6, L=3, ct="Circuit_15", typ="probs", B_I=2, B_P=2, noise=None),
        dict(n=4, L=2, ct="Strongly_Entangling", typ="density", B_I=3, B_P=1, noise=NOISE_A),
        dict(n=3, L=4, ct="Strongly_Entangling", typ="expval", B_I=2, B_P=2, noise=NOISE_A),
        dict(n=8, L=2, ct="Hardware_Efficient", typ="expval", B_I=1, B_P=2, noise=None),
        dict(n=3, L=2, ct="Circuit_6", typ="density", B_I=2, B_P=1, noise=NOISE_B),
        dict(n=2, L=2, ct="Circuit_19", typ="probs", B_I=2, B_P=2, noise=NOISE_B),
    ]
    for a in Ansaetze.get_available():
        name = a.__name__
        out.append(dict(n=4, L=2, ct=name, typ="state", B_I=2,
                        B_P=1 if name in ("GHZ", "No_Ansatz") else 2, noise=None))
        out.append(dict(n=3, L=1, ct=name, typ="expval", B_I=2, B_P=1, noise=None))
    return out


def option_cases():
    """Constructor options of model.py:26-45 (encodings, re-uploading, output qubits, state
    preparation).  `kw` is JSON: an encoding is a gate name, a list of gate names (one input
    feature each) or {"strategy", "gates"} for `Encoding(strategy, gates)`."""
    base = dict(n=3, L=2, ct="Circuit_19", B_I=2, B_P=2, noise=None)
    kws = [
        (dict(encoding=["RX", "RY"]), "expval"),
        (dict(encoding="RY"), "expval"),
        (dict(encoding={"strategy": "binary", "gates": ["RX"]}), "expval"),
        (dict(encoding={"strategy": "ternary", "gates": ["RX"]}), "expval"),
        (dict(encoding={"strategy": "golomb", "gates": ["RX"]}), "expval"),
        (dict(encoding={"strategy": "binary", "gates": ["RY"]}), "probs"),
        (dict(data_reupload=False), "expval"),
        (dict(output_qubit=0), "expval"),
        (dict(output_qubit=[0, 2]), "expval"),
        (dict(output_qubit=[0, 2]), "probs"),
        (dict(output_qubit=[0, 1]), "density"),
        (dict(state_preparation="H"), "probs"),
        (dict(state_preparation="H", encoding=["RX", "RZ"], output_qubit=[1, 2]), "probs"),
    ]
    out = [dict(base, kw=kw, typ=typ) for kw, typ in kws]
    # every noise_params key of model.py:253-265 incl. state-preparation / measurement flips
    # and depth-dependent thermal relaxation (replayed through the API only: the channel
    # strength depends on the circuit depth the Model computes)
    full = {"BitFlip": 0.01, "PhaseFlip": 0.02, "Depolarizing": 0.03,
            "MultiQubitDepolarizing": 0.04, "AmplitudeDamping": 0.05, "PhaseDamping": 0.06,
            "StatePreparation": 0.07, "Measurement": 0.08,
            "ThermalRelaxation": {"t1": 2000.0, "t2": 1000.0, "t_factor": 1.0}}
    # ONE sample per call for the depth-dependent thermal channel: the reference computes the
    # depth with zero inputs and drops the zero encodings only when batch_shape[0] == 1
    # (model.py:782, 1085-1098), and `_inputs_validation(None)` inside it flips
    # `self._zero_inputs` while the circuit is being recorded - under real `jax.vmap` the
    # circuit is traced once, under the stand-in's loop it would be re-recorded per sample
    # with the flipped flag, so a batched thermal call is NOT reproduced faithfully here
    one = dict(base, B_I=1, B_P=1)
    out.append(dict(one, kw={}, api_only=True, typ="probs", noise=full))
    out.append(dict(one, n=2, ct="Strongly_Entangling", kw={}, api_only=True, typ="density",
                    noise=full))
    out.append(dict(one, kw={}, api_only=True, typ="expval",
                    noise={"ThermalRelaxation": {"t1": 1000.0, "t2": 1800.0, "t_factor": 2.0}}))
    out.append(dict(base, kw=dict(output_qubit=[0, 1]), typ="density", noise=NOISE_A))
    out.append(dict(base, n=4, ct="Hardware_Efficient", kw=dict(encoding=["RX", "RY"]),
                    typ="expval", noise=NOISE_A))
    return out


def _ref_kwargs(kw):
    from qml_essentials.ansaetze import Encoding

    kw = dict(kw)
    if isinstance(kw.get("encoding"), dict):
        kw["encoding"] = Encoding(kw["encoding"]["strategy"], kw["encoding"]["gates"])
    return kw


def main():
    rng = np.random.default_rng(20260002)
    store, index = {}, []
    for ci, c in enumerate(cases() + option_cases()):
        kw = c.get("kw", {})
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = rmodel.Model(n_qubits=c["n"], n_layers=c["L"], circuit_type=c["ct"],
                             **_ref_kwargs(kw))
        shape = tuple(np.shape(m.params))[1:]
        params = rng.uniform(0, 2 * np.pi, (c["B_P"],) + shape)
        n_feat = len(kw["encoding"]) if isinstance(kw.get("encoding"), list) else 1
        inputs = rng.uniform(-np.pi, np.pi, (c["B_I"], n_feat))
        res = []
        for i in range(c["B_I"]):
            for p in range(c["B_P"]):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    r = m(params=params[p:p + 1], inputs=inputs[i:i + 1],
                          execution_type=c["typ"],
                          noise_params=dict(c["noise"]) if c["noise"] else None)
                res.append(np.asarray(r).reshape(-1))
        # the same batch in ONE call of the reference (its vmapped route, script.py:399-553,
        # on a fresh model: the first call fixes the cached circuit depth)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mb = rmodel.Model(n_qubits=c["n"], n_layers=c["L"], circuit_type=c["ct"],
                              **_ref_kwargs(kw))
            batched = np.asarray(mb(params=params, inputs=inputs, execution_type=c["typ"],
                                    noise_params=dict(c["noise"]) if c["noise"] else None))
        tag = f"model{ci}"
        store[tag + "_batched"] = batched
        assert np.abs(batched.reshape(len(res), -1) - np.stack(res)).max() < 1e-12, (
            "the reference's batched and per-sample routes disagree under the stand-in", c)
        c = dict(c, batched_equals_samples=bool(
            np.abs(batched.reshape(len(res), -1) - np.stack(res)).max() < 1e-12))
        store[tag + "_params"] = params
        store[tag + "_inputs"] = inputs
        store[tag + "_out"] = np.stack(res)  # row b = i * B_P + p, flattened result
        index.append(dict(c, id=ci, params_shape=list(shape)))
        print(tag, c["ct"], c["typ"], kw, store[tag + "_out"].shape, batched.shape,
              c["batched_equals_samples"])
    store["index_json"] = np.frombuffer(json.dumps(index).encode(), dtype=np.uint8)
    out = os.path.join(HERE, "..", "tests", "golden", "reference_model.npz")
    np.savez_compressed(out, **store)
    print(f"wrote {os.path.normpath(out)}: {len(index)} cases, {os.path.getsize(out)} bytes")


if __name__ == "__main__":
    main()
