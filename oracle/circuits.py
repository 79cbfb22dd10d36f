"""Restatement of the reference's circuit program (oracle; test-only).

Produces the tape (list of ``(name, wires, params, extra)``) that
``Model._variational`` records for a given configuration, independently of the
product's host code: topologies (topologies.py:22-121), the 23 declarative
ansaetze (ansaetze.py:410-756), block application (ansaetze.py:216-221,323-371),
encoding (model.py:746-816, ansaetze.py:893-1000), noise insertion
(unitary.py:150-246, model.py:1000-1064) and the layer order (model.py:818-963).
"""

import itertools

import numpy as np

from . import gates as G


# ---------------------------------------------------------------- topologies
def stairs(n, offset=0, wrap=False, reverse=True, mirror=True, span=1, stride=1,
           modulo=True):
    """topologies.py:22-100."""
    off = offset(n) if callable(offset) else offset
    sp = span(n) if callable(span) else span
    pairs = []
    for q in range(0, n if wrap else n - 1, stride):
        c, t = q + off, q + off + sp
        if (t >= n or c < 0) and not modulo:
            continue
        c, t = c % n, t % n
        if c == t:
            continue
        pairs.append((c, t))
    if reverse:
        pairs = pairs[::-1]
    if mirror:
        pairs = [(t, c) for c, t in pairs]
    return pairs


def bricks(n, **kw):
    """topologies.py:102-106."""
    kw.setdefault("stride", 2)
    kw.setdefault("modulo", False)
    return stairs(n, **kw)


def all_to_all(n):
    """topologies.py:108-121."""
    return [
        (n - ql - 1, (n - q - 1) % n) for ql in range(n) for q in range(n) if q != ql
    ]


TOPO = {"stairs": stairs, "bricks": bricks, "all_to_all": all_to_all}

# ------------------------------------------------------------------ ansaetze
# (gate, topology-name or None, topology kwargs)  - ansaetze.py:410-756
_13 = dict(reverse=False, mirror=False, offset=lambda n: n - 1, span=3, wrap=True)
_W = dict(wrap=True, reverse=True, mirror=False)
ANSAETZE = {
    "No_Ansatz": [],
    "Circuit_1": [("RX",), ("RZ",)],
    "Circuit_2": [("RX",), ("RZ",), ("CX", "stairs", {})],
    "Circuit_3": [("RX",), ("RZ",), ("CRZ", "stairs", {})],
    "Circuit_4": [("RX",), ("RZ",), ("CRX", "stairs", {})],
    "Circuit_5": [("RX",), ("RZ",), ("CRZ", "all_to_all", {}), ("RX",), ("RZ",)],
    "Circuit_6": [("RX",), ("RZ",), ("CRX", "all_to_all", {}), ("RX",), ("RZ",)],
    "Circuit_7": [("RX",), ("RZ",), ("CRZ", "bricks", {}), ("RX",), ("RZ",),
                  ("CRZ", "bricks", dict(offset=1))],
    "Circuit_8": [("RX",), ("RZ",), ("CRX", "bricks", {}), ("RX",), ("RZ",),
                  ("CRX", "bricks", dict(offset=1))],
    "Circuit_9": [("H",), ("CZ", "stairs", {}), ("RX",)],
    "Circuit_10": [("RY",), ("CZ", "stairs", dict(offset=-1, wrap=True)), ("RY",)],
    "Circuit_13": [("RY",), ("CRZ", "stairs", _W), ("RY",), ("CRZ", "stairs", _13)],
    "Circuit_14": [("RY",), ("CRX", "stairs", _W), ("RY",), ("CRX", "stairs", _13)],
    "Circuit_15": [("RY",), ("CX", "stairs", _W), ("RY",), ("CX", "stairs", _13)],
    "Circuit_16": [("RX",), ("RZ",), ("CRZ", "bricks", {}),
                   ("CRZ", "bricks", dict(offset=1))],
    "Circuit_17": [("RX",), ("RZ",), ("CRX", "bricks", {}),
                   ("CRX", "bricks", dict(offset=1))],
    "Circuit_18": [("RX",), ("RZ",), ("CRZ", "stairs", dict(wrap=True, mirror=False))],
    "Circuit_19": [("RX",), ("RZ",), ("CRX", "stairs", dict(wrap=True, mirror=False))],
    "Circuit_20": [("RY",), ("CX", "stairs", _W), ("RY",),
                   ("CX", "stairs", dict(reverse=False, offset=lambda n: n - 2, span=1,
                                         wrap=True))],
    "No_Entangling": [("Rot",)],
    "Hardware_Efficient": [("RY",), ("RZ",), ("RY",),
                           ("CX", "bricks", dict(mirror=False)),
                           ("CX", "bricks", dict(offset=-1, modulo=True, wrap=True,
                                                 mirror=False))],
    "Strongly_Entangling": [("Rot",),
                            ("CX", "stairs", dict(wrap=True, reverse=False,
                                                  mirror=False)),
                            ("Rot",),
                            ("CX", "stairs", dict(reverse=False, span=lambda n: n // 2,
                                                  wrap=True, mirror=False))],
}
ROTATIONAL = {"RX", "RY", "RZ", "Rot", "CRX", "CRY", "CRZ", "CPhase"}  # gates.py:210-221
ENTANGLING = {"CX", "CY", "CZ", "CRX", "CRY", "CRZ", "CPhase"}  # gates.py:224-225


def _block_pairs(block, n):
    gate, topo, kw = block
    span = kw.get("span", 1)
    span = span(n) if callable(span) else span
    if not (n >= 2 and n > span):  # ansaetze.py:274-284
        return []
    return TOPO[topo](n, **kw)


def n_params_per_layer(circuit_type, n):
    """ansaetze.py:174-175,286-303."""
    if circuit_type == "GHZ":
        return 0
    total = 0
    for block in ANSAETZE[circuit_type]:
        gate = block[0]
        if gate not in ROTATIONAL:
            continue
        if gate in ENTANGLING:
            total += len(_block_pairs(block, n))
        else:
            total += 3 * n if gate == "Rot" else n
    return total


# ------------------------------------------------------------ noise insertion
class _Emit:
    """Collects tape entries; mirrors UnitaryGates wrappers (unitary.py:249-701)."""

    def __init__(self, noise, jitter):
        self.tape = []
        self.noise = noise
        self.jitter = jitter  # iterator of N(0,1) draws in reference draw order

    def _gate_error(self, w):
        # unitary.py:226-246: drawn whenever the key is present (also for sigma 0)
        if self.noise is not None and self.noise.get("GateError", None) is not None:
            z = next(self.jitter) if self.jitter is not None else 0.0
            return w + self.noise["GateError"] * z
        return w

    def _noise(self, wires):
        """unitary.py:175-197."""
        if self.noise is None:
            return
        for w in wires:
            for key, name in (("BitFlip", "BitFlip"), ("PhaseFlip", "PhaseFlip"),
                              ("Depolarizing", "DepolarizingChannel")):
                p = self.noise.get(key, 0.0)
                if p > 0:
                    self.tape.append((name, [w], [p], None))
        if len(wires) > 1:
            p = self.noise.get("MultiQubitDepolarizing", 0.0)
            if p > 0:
                self.tape.append(
                    ("QubitChannel", list(wires), [],
                     G.n_qubit_depolarizing_kraus(p, len(wires)))
                )

    def gate(self, name, wires, angles=()):
        wires = [wires] if isinstance(wires, (int, np.integer)) else list(wires)
        if name == "Rot":
            if self.noise is not None and "GateError" in self.noise:  # unitary.py:275
                angles = [self._gate_error(a) for a in angles]
        elif angles:
            angles = [self._gate_error(a) for a in angles]
        tape_name = {"CPhase": "ControlledPhaseShift"}.get(name, name)
        self.tape.append((tape_name, wires, list(angles), None))
        self._noise(wires)

    def barrier(self, n):
        self.tape.append(("Barrier", list(range(n)), [], None))


def _apply_ansatz(em, circuit_type, w, n):
    """DeclarativeCircuit.build (ansaetze.py:216-221) + Block.apply (:323-371)."""
    if circuit_type == "GHZ":  # ansaetze.py:423-427
        em.gate("H", 0)
        for q in range(n - 1):
            em.gate("CX", [q, q + 1])
        return
    idx = 0
    for block in ANSAETZE[circuit_type]:
        gate = block[0]
        targets = _block_pairs(block, n) if gate in ENTANGLING else range(n)
        for wires in targets:
            if gate == "Rot":
                em.gate(gate, wires, [w[idx], w[idx + 1], w[idx + 2]])
                idx += 3
            elif gate in ROTATIONAL:
                em.gate(gate, wires, [w[idx]])
                idx += 1
            else:
                em.gate(gate, wires)
        em.barrier(n)


def golomb_ruler(d):
    """unitary.py:18-50 (greedy)."""
    marks, diffs, cand = [0], set(), 1
    while len(marks) < d:
        new = {cand - m for m in marks}
        if len(new) == len(marks) and not (new & diffs):
            marks.append(cand)
            diffs |= new
        cand += 1
    return tuple(marks)


def circuit_depth(tape):
    """model.py:1102-1122 (Barriers count, channels do not)."""
    busy, depth = {}, 0
    for e in tape:
        if G.is_channel(e[0]):
            continue
        end = max((busy.get(w, 0) for w in e[1]), default=0) + 1
        for w in e[1]:
            busy[w] = end
        depth = max(depth, end)
    return depth


def variational_tape(
    n_qubits,
    n_layers,
    circuit_type,
    params,
    inputs,
    *,
    encoding=("RX",),
    strategy="hamming",
    data_reupload=True,
    enc_params=None,
    noise_params=None,
    state_preparation=(),
    skip_encoding=False,
    jitter=None,
    depth_for_thermal=None,
):
    """Tape recorded by ``Model._variational`` (model.py:818-963).

    ``params``: (L', P) where entries may be floats or (B,) arrays (batched form);
    ``inputs``: (F,) likewise.  ``skip_encoding`` is the reference's
    ``remove_zero_encoding and zero_inputs and B_I == 1`` shortcut (model.py:782).
    """
    n, L = n_qubits, n_layers
    F = 1 if strategy == "golomb" else len(encoding)
    if isinstance(data_reupload, bool):  # model.py:489-495
        dru = np.ones((L, n, F), bool) if data_reupload else np.zeros((L, n, F), bool)
        if not data_reupload:
            dru[0][0] = True
    else:
        dru = np.asarray(data_reupload).astype(bool)
        if dru.ndim == 2:
            dru = np.repeat(dru[..., None], F, axis=2)
    if enc_params is None:
        enc_params = np.ones((L, n, F))  # model.py:150

    def max_freq(count):  # ansaetze.py:872-889
        if strategy == "hamming":
            return count
        if strategy == "binary":
            return 2**count - 1
        if strategy == "ternary":
            return int(np.floor(3**count / 2))
        return count * max(golomb_ruler(2**n))

    has_dru = max(max_freq(int(np.count_nonzero(dru[..., i]))) for i in range(F)) > 1

    noise = None
    if noise_params is not None and not all(v == 0.0 for v in noise_params.values()):
        noise = dict(noise_params)  # model.py:249-267 (defaults filled in)
        for key in ("BitFlip", "PhaseFlip", "Depolarizing", "MultiQubitDepolarizing",
                    "AmplitudeDamping", "PhaseDamping", "GateError", "StatePreparation",
                    "Measurement"):
            noise.setdefault(key, 0.0)
        noise.setdefault("ThermalRelaxation", None)

    em = _Emit(noise, iter(jitter) if jitter is not None else None)
    if noise is not None and noise.get("StatePreparation", 0.0) > 0:  # model.py:1017-1020
        for q in range(n):
            em.tape.append(("BitFlip", [q], [noise["StatePreparation"]], None))
    for q in range(n):  # model.py:914-923
        for sp in state_preparation:
            em.gate(sp, q)

    def iec(layer):  # model.py:746-816
        if skip_encoding:
            return
        if strategy == "golomb":
            if dru[layer][:, 0].any():
                x = inputs[0] * np.mean(enc_params[layer][:, 0])
                x = em._gate_error(x)
                marks = np.array(golomb_ruler(2**n), dtype=float)
                em.tape.append(("DiagonalQubitUnitary", list(range(n)), [],
                                np.exp(-1j * marks * x)))
                em._noise(list(range(n)))
            return
        for q in range(n):
            for f in range(F):
                if dru[layer][q, f]:
                    x = inputs[f] * enc_params[layer][q, f]
                    if strategy == "binary":
                        x = x * (2**q)  # ansaetze.py:933-934
                    elif strategy == "ternary":
                        x = x * (3**q)  # ansaetze.py:958-959
                    em.gate(encoding[f], q, [x])

    for layer in range(L):
        _apply_ansatz(em, circuit_type, params[layer], n)
        iec(layer)
    if has_dru:  # model.py:950-959
        _apply_ansatz(em, circuit_type, params[L], n)

    if noise is not None:  # model.py:1047-1064
        tr = noise.get("ThermalRelaxation", 0.0)
        for q in range(n):
            if noise.get("AmplitudeDamping", 0.0) > 0:
                em.tape.append(("AmplitudeDamping", [q], [noise["AmplitudeDamping"]], None))
            if noise.get("PhaseDamping", 0.0) > 0:
                em.tape.append(("PhaseDamping", [q], [noise["PhaseDamping"]], None))
            if noise.get("Measurement", 0.0) > 0:
                em.tape.append(("BitFlip", [q], [noise["Measurement"]], None))
            if isinstance(tr, dict):
                tg = depth_for_thermal * tr["t_factor"]
                em.tape.append(
                    ("ThermalRelaxationError", [q], [1.0, tr["t1"], tr["t2"], tg], None)
                )
    return em.tape


def assimilate_index(B_I, B_P, B_R=1):
    """Flat batch order of model.py:1414-1483: b = (i*B_P + p)*B_R + r."""
    idx = np.array(list(itertools.product(range(B_I), range(B_P), range(B_R))))
    return idx[:, 0], idx[:, 1], idx[:, 2]
