"""Parity cases shared by the CPU suite (program interpreter) and the GPU suite
(CUDA library through the C ABI).  Each case runs the PRODUCT host path
(`Script` / `Model`) with whatever executor is installed and compares against the
reference-faithful oracle (`oracle.sim`, `oracle.circuits`) on the same seeded
inputs.  Tolerances: 1e-10 for complex128, 1e-5 for complex64 (BASELINE.json)."""

import warnings

import numpy as np

from oracle import circuits as oc
from oracle import gates as og
from oracle import sim as osim
from qml_essentials_b200 import operations as op
from qml_essentials_b200.ansaetze import Ansaetze, Encoding
from qml_essentials_b200.model import Model
from qml_essentials_b200.script import Script

TOL = {"complex128": 1e-10, "complex64": 1e-5}


def _tape_of(ops):
    """Oracle tape spec from recorded product operations (concrete values)."""
    out = []
    for o in ops:
        name = type(o).__name__
        if isinstance(o, op.KrausChannel):
            out.append(("QubitChannel", list(o.wires), [], o.kraus_matrices()))
        elif name == "Barrier":
            out.append(("Barrier", list(o.wires), []))
        else:
            out.append(("QubitUnitary", list(o.wires), [], np.asarray(o.matrix)))
    return out


def _zs(n):
    return [("PauliZ", [q], []) for q in range(n)]


# ----------------------------------------------------------------------------------
def case_every_gate(precision="complex128", n=4, batch=5, seed=0):
    """Every gate class of operations.py:719-1487 on a random state, batched angles."""
    rng = np.random.default_rng(seed)
    thetas = rng.uniform(-np.pi, np.pi, (batch, 12))

    def circuit(t):
        for q in range(n):
            op.H(wires=q)
        op.RX(t[0], wires=0); op.RY(t[1], wires=1); op.RZ(t[2], wires=2)
        op.Rot(t[3], t[4], t[5], wires=3)
        op.PauliX(wires=1); op.PauliY(wires=2); op.PauliZ(wires=0); op.S(wires=3)
        op.CX(wires=[0, 2]); op.CY(wires=[3, 1]); op.CZ(wires=[2, 3]); op.SWAP(wires=[0, 3])
        op.CRX(t[6], wires=[1, 0]); op.CRY(t[7], wires=[2, 1]); op.CRZ(t[8], wires=[0, 3])
        op.ControlledPhaseShift(t[9], wires=[3, 2])
        op.RXX(t[10], wires=[0, 1]); op.RYY(t[11], wires=[1, 3]); op.RZZ(t[0], wires=[2, 0])
        op.RZX(t[1], wires=[3, 0]); op.PauliRot(t[2], "XYZ", wires=[1, 2, 3])
        op.CCX(wires=[0, 1, 2]); op.CSWAP(wires=[3, 0, 1])
        op.ControlledPauliRot(t[3], "X", wires=[0, 1, 2], n_controls=2)
        op.Id(wires=[0, 1]); op.Barrier(wires=[0, 1, 2, 3])
        op.RX(t[4], wires=2).dagger()
        op.QubitUnitary(og.HAD, wires=1)
        op.DiagonalQubitUnitary(np.exp(1j * np.arange(4)), wires=[1, 2])

    def oracle_tape(t):
        tp = [("H", [q], []) for q in range(n)]
        tp += [("RX", [0], [t[0]]), ("RY", [1], [t[1]]), ("RZ", [2], [t[2]]),
               ("Rot", [3], [t[3], t[4], t[5]]), ("PauliX", [1], []), ("PauliY", [2], []),
               ("PauliZ", [0], []), ("S", [3], []), ("CX", [0, 2], []), ("CY", [3, 1], []),
               ("CZ", [2, 3], []), ("SWAP", [0, 3], []), ("CRX", [1, 0], [t[6]]),
               ("CRY", [2, 1], [t[7]]), ("CRZ", [0, 3], [t[8]]),
               ("ControlledPhaseShift", [3, 2], [t[9]]), ("RXX", [0, 1], [t[10]]),
               ("RYY", [1, 3], [t[11]]), ("RZZ", [2, 0], [t[0]]), ("RZX", [3, 0], [t[1]]),
               ("PauliRot", [1, 2, 3], [t[2]], "XYZ"), ("CCX", [0, 1, 2], []),
               ("CSWAP", [3, 0, 1], []),
               ("ControlledPauliRot", [0, 1, 2], [t[3]], ("X", 2)),
               ("Id", [0, 1], []), ("Barrier", [0, 1, 2, 3], []),
               ("RX", [2], [-t[4]]), ("QubitUnitary", [1], [], og.HAD),
               ("DiagonalQubitUnitary", [1, 2], [], np.exp(1j * np.arange(4)))]
        return tp

    s = Script(circuit, n_qubits=n, precision=precision)
    obs = [op.PauliZ(0, record=False), op.PauliX(1, record=False),
           op.Hermitian(np.kron(og.Z, og.Y), wires=[2, 0], record=False),
           op.Hermitian(np.diag([1.0, 2.0, -1.0, 0.5]), wires=[3, 1], record=False)]
    oobs = [("PauliZ", [0], []), ("PauliX", [1], []),
            ("Hermitian", [2, 0], [], np.kron(og.Z, og.Y)),
            ("Hermitian", [3, 1], [], np.diag([1.0, 2.0, -1.0, 0.5]))]
    errs = {}
    for typ in ("state", "probs", "expval", "density"):
        got = s.execute(typ, obs=obs if typ == "expval" else None, args=(thetas,),
                        in_axes=(0,))
        ref = np.stack([osim.simulate_and_measure(oracle_tape(t), n, typ, oobs)
                        for t in thetas])
        errs[typ] = float(np.abs(got - ref).max())
    return errs


def case_every_channel(precision="complex128", n=3, batch=4, seed=1):
    """Every Kraus channel (operations.py:1581-1929) incl. 2-qubit depolarizing."""
    from qml_essentials_b200.unitary import UnitaryGates

    rng = np.random.default_rng(seed)
    th = rng.uniform(0, 2 * np.pi, (batch, 3))
    kraus2 = og.n_qubit_depolarizing_kraus(0.07, 2)

    def circuit(t):
        op.RY(t[0], wires=0); op.BitFlip(0.1, wires=0)
        op.RX(t[1], wires=1); op.PhaseFlip(0.2, wires=1)
        op.CX(wires=[0, 1]); op.DepolarizingChannel(0.15, wires=1)
        op.Rot(t[0], t[1], t[2], wires=2); op.AmplitudeDamping(0.3, wires=2)
        op.CRZ(t[2], wires=[2, 0]); op.PhaseDamping(0.25, wires=0)
        op.ThermalRelaxationError(0.1, 1.5, 1.0, 0.4, wires=1)
        op.ThermalRelaxationError(0.2, 1.0, 1.8, 0.3, wires=2)
        UnitaryGates.NQubitDepolarizingChannel(0.07, [1, 2])
        op.CCX(wires=[2, 0, 1]); op.H(wires=0); op.BitFlip(0.05, wires=0)

    def oracle_tape(t):
        return [("RY", [0], [t[0]]), ("BitFlip", [0], [0.1]), ("RX", [1], [t[1]]),
                ("PhaseFlip", [1], [0.2]), ("CX", [0, 1], []),
                ("DepolarizingChannel", [1], [0.15]), ("Rot", [2], [t[0], t[1], t[2]]),
                ("AmplitudeDamping", [2], [0.3]), ("CRZ", [2, 0], [t[2]]),
                ("PhaseDamping", [0], [0.25]),
                ("ThermalRelaxationError", [1], [0.1, 1.5, 1.0, 0.4]),
                ("ThermalRelaxationError", [2], [0.2, 1.0, 1.8, 0.3]),
                ("QubitChannel", [1, 2], [], kraus2), ("CCX", [2, 0, 1], []),
                ("H", [0], []), ("BitFlip", [0], [0.05])]

    s = Script(circuit, n_qubits=n, precision=precision)
    obs = [op.PauliZ(q, record=False) for q in range(n)] + [
        op.Hermitian(np.kron(og.X, og.Z), wires=[0, 2], record=False)]
    oobs = _zs(n) + [("Hermitian", [0, 2], [], np.kron(og.X, og.Z))]
    errs = {}
    for typ in ("probs", "expval", "density"):
        got = s.execute(typ, obs=obs if typ == "expval" else None, args=(th,), in_axes=(0,))
        ref = np.stack([osim.simulate_and_measure(oracle_tape(t), n, typ, oobs) for t in th])
        errs[typ] = float(np.abs(got - ref).max())
    return errs


def case_model(n, L, circuit_type, B_I, B_P, typ="expval", noise=None,
               precision="complex128", seed=1000, strategy="hamming", **model_kw):
    """Model.__call__ against the oracle's independent restatement of the circuit
    program, in the reference's flat batch order b = i*B_P + p."""
    rng = np.random.default_rng(seed)
    if strategy != "hamming":
        model_kw["encoding"] = Encoding(strategy, ["RX"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = Model(n_qubits=n, n_layers=L, circuit_type=circuit_type, precision=precision,
                  **model_kw)
        params = rng.uniform(0, 2 * np.pi, (B_P, *m._params_shape))
        inputs = np.linspace(-np.pi, np.pi, B_I).reshape(B_I, 1) if B_I > 0 else None
        got = m(params=params, inputs=inputs, execution_type=typ,
                noise_params=dict(noise) if noise else None)
        depth = m._get_circuit_depth() if noise and isinstance(
            noise.get("ThermalRelaxation"), dict) else None
    bi = max(B_I, 1)
    ref = []
    for i in range(bi):
        for p in range(B_P):
            x = [inputs[i, 0]] if inputs is not None else [0.0]
            skip = (inputs is None or not inputs.any()) and bi == 1
            tape = oc.variational_tape(n, L, circuit_type, params[p], x, noise_params=noise,
                                       skip_encoding=skip, strategy=strategy,
                                       depth_for_thermal=depth)
            ref.append(osim.simulate_and_measure(tape, n, typ, _zs(n)))
    ref = np.array(ref)
    if typ == "probs":
        ref = ref.reshape(len(ref), *(2,) * n)
    ref = ref.reshape((bi, B_P) + ref.shape[1:]).squeeze()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    return float(np.abs(got - ref).max())


def case_baseline_configs(precision="complex128"):
    """Reduced-size versions of the five BASELINE.json configs (same circuits,
    smaller batches / qubit counts where the oracle would take too long)."""
    return {
        "cfg1_c19_n2": case_model(2, 1, "Circuit_19", 64, 1, precision=precision),
        "cfg2_he_n4": case_model(4, 4, "Hardware_Efficient", 33, 8, precision=precision),
        "cfg3_c15_n6_density": case_model(6, 3, "Circuit_15", 0, 6, "density",
                                          precision=precision),
        "cfg4_se_n4_noisy": case_model(
            4, 2, "Strongly_Entangling", 5, 1, "density",
            noise={"Depolarizing": 0.01, "AmplitudeDamping": 0.02}, precision=precision),
        "cfg5_he_n8_expval": case_model(8, 2, "Hardware_Efficient", 1, 2, precision=precision),
    }


def case_all_ansaetze(precision="complex128", n=4):
    out = {}
    for a in Ansaetze.get_available():
        name = a.__name__
        bp = 1 if name in ("GHZ", "No_Ansatz") else 2
        out[name] = case_model(n, 2, name, 3, bp, "state", precision=precision)
    return out


def case_noise_keys(precision="complex128"):
    """Every noise_params key of model.py:253-265 at once, plus thermal relaxation."""
    noise = {"BitFlip": 0.01, "PhaseFlip": 0.02, "Depolarizing": 0.03,
             "MultiQubitDepolarizing": 0.04, "AmplitudeDamping": 0.05, "PhaseDamping": 0.06,
             "StatePreparation": 0.07, "Measurement": 0.08,
             "ThermalRelaxation": {"t1": 2000.0, "t2": 1000.0, "t_factor": 1.0}}
    return {
        "c19_n3_probs": case_model(3, 1, "Circuit_19", 3, 2, "probs", noise=noise,
                                   precision=precision),
        "se_n2_density": case_model(2, 2, "Strongly_Entangling", 2, 2, "density",
                                    noise=noise, precision=precision),
    }


def case_shots(precision="complex128", n=3, batch=4, shots=2000, seed=5):
    """Counts must equal the oracle's bit for bit given the same uniform stream."""
    from qml_essentials_b200 import rng as qrng

    rng = np.random.default_rng(seed)
    th = rng.uniform(0, 2 * np.pi, (batch, n))

    def circuit(t):
        for q in range(n):
            op.RY(t[q], wires=q)
        op.CX(wires=[0, 1]); op.CX(wires=[1, 2])

    s = Script(circuit, n_qubits=n, precision=precision)
    key = qrng.key(42)
    est = s.execute("probs", args=(th,), in_axes=(0,), shots=shots, key=key)
    uniforms = qrng.choice_uniforms(qrng.split(key, batch), shots)
    exact = s.execute("probs", args=(th,), in_axes=(0,))
    mism = 0
    for b in range(batch):
        p = exact[b]
        u = uniforms[b]
        if precision == "complex64":
            p = p.astype(np.float32)
            cum = np.cumsum(p, dtype=np.float32)
            r = cum[-1] * (np.float32(1) - u.astype(np.float32))
            idx = np.minimum(np.searchsorted(cum, r, side="left"), 2**n - 1)
            counts = np.bincount(idx, minlength=2**n)
        else:
            _, counts = osim.sample_shots(p, n, "probs", [], shots, u)
        mism += int(np.abs(np.rint(est[b] * shots).astype(np.int64) - counts).sum())
    obs = [op.PauliZ(q, record=False) for q in range(n)]
    ev = s.execute("expval", obs=obs, args=(th,), in_axes=(0,), shots=shots, key=key)
    signs = np.array([[1 - 2 * ((i >> (n - 1 - q)) & 1) for i in range(2**n)]
                      for q in range(n)])
    ev_ref = est @ signs.T
    return {"count_mismatch": mism, "expval_err": float(np.abs(ev - ev_ref).max()),
            "sums": float(np.abs(est.sum(axis=1) - 1).max())}


def case_edge_cases(precision="complex128"):
    """Edge cases of the path (empty / permutation-only circuits, one qubit, ragged last CTA,
    single-element batches, one shot, deterministic probabilities): max |result - expected|
    with exact expectations or the oracle."""
    errs = {}
    z = lambda n: [op.PauliZ(q, record=False) for q in range(n)]  # noqa: E731

    # empty circuit (only a barrier): |000>
    s = Script(lambda: op.Barrier(wires=[0, 1, 2]), n_qubits=3, precision=precision)
    p = s.execute("probs")
    errs["empty_probs"] = float(np.abs(p - np.eye(8)[0]).max())
    errs["empty_expval"] = float(np.abs(s.execute("expval", obs=z(3)) - 1).max())
    errs["empty_density"] = float(np.abs(s.execute("density") - np.outer(np.eye(8)[0],
                                                                         np.eye(8)[0])).max())

    # permutation / diagonal ops only: X, CX, CCX, SWAP, Z, CZ  ->  a basis state with a sign
    def perm_only():
        op.PauliX(wires=0); op.CX(wires=[0, 2]); op.CCX(wires=[0, 2, 1])
        op.SWAP(wires=[1, 3]); op.PauliZ(wires=0); op.CZ(wires=[0, 3])

    s = Script(perm_only, n_qubits=4, precision=precision)
    st = s.execute("state")
    want = np.zeros(16, dtype=complex)
    want[0b1011] = 1.0  # X0 -> 1000, CX02 -> 1010, CCX(0,2;1) -> 1110, SWAP13 -> 1011, Z0 * CZ03
    errs["perm_only_state"] = float(np.abs(st - want).max())

    # one qubit, batch of one element along an axis, and a ragged batch (130 = 128 + 2)
    for B in (1, 130):
        th = np.linspace(-2.0, 2.0, B).reshape(B, 1)
        s = Script(lambda t: op.RX(t[0], wires=0), n_qubits=1, precision=precision)
        ev = s.execute("expval", obs=z(1), args=(th,), in_axes=(0,))
        errs[f"rx_cos_batch{B}"] = float(np.abs(ev[:, 0] - np.cos(th[:, 0])).max())
        assert ev.shape == (B, 1)

    # one-qubit noisy density: RX(pi) then BitFlip(0.5) -> maximally mixed diagonal
    def noisy():
        op.RX(np.pi, wires=0); op.BitFlip(0.5, wires=0)

    rho = Script(noisy, n_qubits=1, precision=precision).execute("density")
    errs["bitflip_half"] = float(np.abs(rho - np.eye(2) / 2).max())

    # shots on a deterministic distribution (|10>): every shot lands on index 2, one shot too
    def det():
        op.PauliX(wires=0)

    s = Script(det, n_qubits=2, precision=precision)
    for shots in (1, 7):
        est = s.execute("probs", shots=shots)
        errs[f"det_shots{shots}"] = float(np.abs(est - np.eye(4)[2]).max())

    # repeated / subset observables
    def bell():
        op.H(wires=0); op.CX(wires=[0, 1])

    s = Script(bell, n_qubits=3, precision=precision)
    ev = s.execute("expval", obs=[op.PauliZ(2, record=False), op.PauliZ(0, record=False),
                                  op.PauliZ(2, record=False)])
    errs["subset_obs"] = float(np.abs(ev - np.array([1.0, 0.0, 1.0])).max())
    return errs
