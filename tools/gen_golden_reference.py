"""Golden vectors from the reference's OWN simulator source, run in this container.

    python tools/gen_golden_reference.py        # writes tests/golden/reference_sim.npz

The reference (`/root/reference/qml_essentials`) is pure Python on JAX and JAX is not
installed here.  Its hot-path files `operations.py`, `simulation.py` and `tape.py` touch only
a NumPy-shaped slice of the JAX API, so this script puts `tools/jax_numpy_shim` (a NumPy
stand-in for that slice: eager, float64 / complex128) in front of `sys.path` and imports the
UNMODIFIED reference modules from where they lie.  Every case below is built from the
reference's gate / channel classes and pushed through the reference's `simulate_pure` /
`simulate_mixed` / `measure_state` / `measure_density`; inputs (as plain tape entries) and
outputs go into one .npz that `tests/test_reference_golden.py` replays against `oracle/`.
The fixture cannot be regenerated on the GPU box (no `/root/reference` there) and nothing
reads `/root/reference` at test time.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "jax_numpy_shim"))
sys.path.insert(1, "/root/reference")

import numpy as np  # noqa: E402

import qml_essentials.operations as rop  # noqa: E402
import qml_essentials.simulation as rsim  # noqa: E402

assert rop.__file__.startswith("/root/reference/"), rop.__file__

FIXED = {"Id": 1, "PauliX": 1, "PauliY": 1, "PauliZ": 1, "H": 1, "S": 1, "SWAP": 2, "CX": 2,
         "CY": 2, "CZ": 2, "CCX": 3, "CSWAP": 3}
ONE_ANGLE = {"RX": 1, "RY": 1, "RZ": 1, "CRX": 2, "CRY": 2, "CRZ": 2, "RXX": 2, "RYY": 2,
             "RZZ": 2, "RZX": 2, "ControlledPhaseShift": 2}
CHANNELS = ("BitFlip", "PhaseFlip", "DepolarizingChannel", "AmplitudeDamping", "PhaseDamping")


def build(entry):
    """Tape entry (name, wires, params, extra) -> reference operation object."""
    name, wires, params, extra = entry
    cls = getattr(rop, name)
    w = wires[0] if len(wires) == 1 else list(wires)
    if name in FIXED:
        return cls(wires=w)
    if name in ONE_ANGLE or name in CHANNELS:
        return cls(params[0], wires=w)
    if name == "Rot":
        return cls(*params, wires=w)
    if name == "PauliRot":
        return cls(params[0], extra, wires=w)
    if name == "ControlledPauliRot":
        return cls(params[0], extra[0], wires=list(wires), n_controls=extra[1])
    if name == "ThermalRelaxationError":
        return cls(*params, wires=w)
    if name == "DiagonalQubitUnitary":
        return cls(np.asarray(extra), wires=w)
    if name == "Hermitian":
        return cls(np.asarray(extra), wires=w, record=False)
    if name == "QubitChannel":
        return cls([np.asarray(k) for k in extra], wires=w)
    raise KeyError(name)


def random_entry(rng, n, noisy):
    pool = list(FIXED) + list(ONE_ANGLE) * 2 + ["Rot", "Rot", "PauliRot", "ControlledPauliRot",
                                                "DiagonalQubitUnitary"]
    if noisy:
        pool += list(CHANNELS) * 2 + ["ThermalRelaxationError", "QubitChannel"]
    while True:
        name = pool[rng.integers(len(pool))]
        k = FIXED.get(name) or ONE_ANGLE.get(name) or {
            "Rot": 1, "PauliRot": int(rng.integers(1, 4)), "ControlledPauliRot": int(rng.integers(2, 4)),
            "DiagonalQubitUnitary": int(rng.integers(1, 3)), "ThermalRelaxationError": 1,
            "QubitChannel": 1}.get(name, 1)
        if k <= n:
            break
    wires = [int(x) for x in rng.permutation(n)[:k]]
    params, extra = [], None
    if name in ONE_ANGLE:
        params = [float(rng.uniform(-2 * np.pi, 2 * np.pi))]
    elif name in CHANNELS:
        params = [float(rng.uniform(0.0, 0.4))]
    elif name == "Rot":
        params = [float(x) for x in rng.uniform(-np.pi, np.pi, 3)]
    elif name == "PauliRot":
        params = [float(rng.uniform(-np.pi, np.pi))]
        extra = "".join(rng.choice(list("XYZI"), k))
        if set(extra) == {"I"}:
            extra = "Z" + extra[1:]
    elif name == "ControlledPauliRot":
        n_controls = int(rng.integers(1, k))
        params = [float(rng.uniform(-np.pi, np.pi))]
        extra = ("".join(rng.choice(list("XYZ"), k - n_controls)), n_controls)
    elif name == "DiagonalQubitUnitary":
        extra = np.exp(1j * rng.uniform(-np.pi, np.pi, 2 ** k))
    elif name == "ThermalRelaxationError":
        t1 = float(rng.uniform(20.0, 100.0))
        t2 = float(rng.uniform(10.0, 2 * t1)) if rng.random() < 0.5 else float(rng.uniform(5.0, t1))
        params = [float(rng.uniform(0.0, 0.3)), t1, t2, float(rng.uniform(0.5, 10.0))]
    elif name == "QubitChannel":
        # a random 1-qubit channel: Stinespring isometry cut into three Kraus matrices
        a = rng.normal(size=(6, 2)) + 1j * rng.normal(size=(6, 2))
        q, _ = np.linalg.qr(a)
        extra = [q[0:2], q[2:4], q[4:6]]
    return (name, wires, params, extra)


def observables(rng, n):
    """Z on every wire (the Model default), one dense and one two-wire Hermitian."""
    obs = [("PauliZ", [q], [], None) for q in range(n)]
    a = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    obs.append(("Hermitian", [int(rng.integers(n))], [], a + a.conj().T))
    if n >= 2:
        b = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
        obs.append(("Hermitian", [int(x) for x in rng.permutation(n)[:2]], [], b + b.conj().T))
    return obs


def main():
    rng = np.random.default_rng(20260001)
    store, index = {}, []

    def put(key, arr):
        store[key] = np.asarray(arr)

    # (1) every gate / channel alone: its matrix or Kraus set as the reference builds it
    singles = []
    for name in list(FIXED) + list(ONE_ANGLE) + ["Rot", "PauliRot", "ControlledPauliRot",
                                                 "DiagonalQubitUnitary"]:
        for _ in range(2):
            e = random_entry(rng, 3, False)
            while e[0] != name:
                e = random_entry(rng, 3, False)
            singles.append(e)
    for name in list(CHANNELS) + ["ThermalRelaxationError", "ThermalRelaxationError",
                                  "ThermalRelaxationError", "QubitChannel"]:
        e = random_entry(rng, 2, True)
        while e[0] != name:
            e = random_entry(rng, 2, True)
        singles.append(e)
    for i, e in enumerate(singles):
        op = build(e)
        if isinstance(op, rop.KrausChannel):
            put(f"single{i}_kraus", np.stack([np.asarray(k) for k in op.kraus_matrices()]))
        else:
            put(f"single{i}_matrix", np.asarray(op.matrix))
        index.append({"kind": "single", "id": i, "entry": _json_entry(e, store, f"single{i}")})

    # (2) random circuits through the reference's simulate_* and measure_*
    cases = [(n, d, False) for n in (1, 2, 3, 4, 5) for d in (6, 18)] + \
            [(n, d, True) for n in (1, 2, 3, 4) for d in (8, 16)]
    for ci, (n, depth, noisy) in enumerate(cases):
        tape = [random_entry(rng, n, noisy) for _ in range(depth)]
        if noisy and not any(e[0] in CHANNELS + ("ThermalRelaxationError", "QubitChannel")
                             for e in tape):
            tape.append(("DepolarizingChannel", [0], [0.1], None))
        obs = observables(rng, n)
        rtape = [build(e) for e in tape]
        robs = [build(o) for o in obs]
        tag = f"case{ci}"
        if noisy:
            rho = rsim.simulate_mixed(rtape, n)
            put(f"{tag}_density", rho)
            put(f"{tag}_probs", rsim.measure_density(rho, n, "probs", robs))
            put(f"{tag}_expval", rsim.measure_density(rho, n, "expval", robs))
        else:
            psi = rsim.simulate_pure(rtape, n)
            put(f"{tag}_state", psi)
            put(f"{tag}_probs", rsim.measure_state(psi, n, "probs", robs))
            put(f"{tag}_expval", rsim.measure_state(psi, n, "expval", robs))
            # pure circuit asked for its density matrix: |psi><psi| (simulation.py:176-190)
            put(f"{tag}_density", rsim.simulate_and_measure(rtape, n, "density", robs,
                                                            use_density=True))
            # the same circuit evolved gate by gate as a density matrix (simulation.py:107-128)
            put(f"{tag}_density_mixed", rsim.simulate_mixed(rtape, n))
        # Z observables alone take the reference's diagonal fast path (simulation.py:237-258)
        zobs = [o for o in robs if type(o).__name__ == "PauliZ"]
        put(f"{tag}_expval_z", rsim.simulate_and_measure(rtape, n, "expval", zobs,
                                                         use_density=noisy))
        for typ in ("probs", "expval"):
            out = rsim.simulate_and_measure(rtape, n, typ, robs, use_density=noisy)
            assert np.allclose(out, store[f"{tag}_{typ}"], atol=1e-13)
        index.append({"kind": "circuit", "id": ci, "n": n, "noisy": noisy,
                      "tape": [_json_entry(e, store, f"{tag}_t{j}") for j, e in enumerate(tape)],
                      "obs": [_json_entry(o, store, f"{tag}_o{j}") for j, o in enumerate(obs)]})

    store["index_json"] = np.frombuffer(json.dumps(index).encode(), dtype=np.uint8)
    out = os.path.join(HERE, "..", "tests", "golden", "reference_sim.npz")
    np.savez_compressed(out, **store)
    print(f"wrote {os.path.normpath(out)}: {len(singles)} single ops, {len(cases)} circuits, "
          f"{os.path.getsize(out)} bytes")


def _json_entry(e, store, key):
    name, wires, params, extra = e
    rec = {"name": name, "wires": list(wires), "params": list(params), "extra": None}
    if isinstance(extra, str):
        rec["extra"] = {"str": extra}
    elif isinstance(extra, tuple):
        rec["extra"] = {"tuple": [extra[0], int(extra[1])]}
    elif extra is not None:
        store[key + "_extra"] = np.asarray(extra)
        rec["extra"] = {"array": key + "_extra"}
    return rec


if __name__ == "__main__":
    main()
