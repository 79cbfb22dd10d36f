"""Einsum restatement of the reference simulator (oracle; test-only).

Follows ``qml_essentials/simulation.py`` and the contraction helpers of
``qml_essentials/operations.py`` one-to-one, in NumPy complex128.  The batched
variants (one einsum per tape op over a ``(B, 2, ..., 2)`` tensor) are what
``jax.vmap(simulate_pure)`` lowers to and double as the CPU baseline in
``bench.py`` (NumPy or torch-CPU back end).
"""

import string

import numpy as np

from . import gates as G

C = np.complex128


def einsum_subscript(n, k, target_axes, batched=False):
    """operations.py:19-50.  ``batched`` prepends a shared batch letter 'Z'."""
    letters = string.ascii_letters
    state_idx = list(letters[:n])
    contracted = [state_idx[ax] for ax in target_axes]
    new_out = [letters[n + i] for i in range(k)]
    gate_idx = new_out + contracted
    result_idx = list(state_idx)
    for i, ax in enumerate(target_axes):
        result_idx[ax] = new_out[i]
    b = "Z" if batched else ""
    return (
        "".join(gate_idx) + "," + b + "".join(state_idx) + "->" + b + "".join(result_idx)
    )


def infer_n_qubits(tape, obs=()):
    """simulation.py:25-39."""
    wires = set()
    for e in list(tape) + list(obs):
        wires.update(e[1])
    return max(wires) + 1 if wires else 1


def uses_density(tape, type):
    """simulation.py:42-57."""
    return type == "density" or any(G.is_channel(e[0]) for e in tape)


def _entry(e):
    name, wires, params = e[0], list(e[1]), list(e[2]) if len(e) > 2 else []
    extra = e[3] if len(e) > 3 else None
    return name, wires, params, extra


def simulate_pure(tape, n_qubits):
    """simulation.py:65-104: psi <- einsum(sub, G, psi) for every non-Barrier op."""
    dim = 2**n_qubits
    psi = np.zeros(dim, dtype=C)
    psi[0] = 1.0
    psi = psi.reshape((2,) * n_qubits)
    for e in tape:
        name, wires, params, extra = _entry(e)
        if name == "Barrier":
            continue  # simulation.py:93
        if G.is_channel(name):
            # operations.py:1539-1549
            raise TypeError(f"{name} is a noise channel; use density simulation")
        k = len(wires)
        gt = G.unitary_matrix(name, wires, params, extra).reshape((2,) * 2 * k)
        psi = np.einsum(einsum_subscript(n_qubits, k, tuple(wires)), gt, psi)
    return psi.reshape(dim)


def apply_unitary_to_density(rho_t, U, wires, n_qubits):
    """operations.py:501-512: U on ket axes, conj(U) on bra axes."""
    k = len(wires)
    Ut = U.reshape((2,) * 2 * k)
    rho_t = np.einsum(einsum_subscript(2 * n_qubits, k, tuple(wires)), Ut, rho_t)
    bra = tuple(w + n_qubits for w in wires)
    return np.einsum(einsum_subscript(2 * n_qubits, k, bra), np.conj(Ut), rho_t)


def simulate_mixed(tape, n_qubits):
    """simulation.py:107-128 with operations.py:485-512 and :1551-1578."""
    dim = 2**n_qubits
    rho = np.zeros((dim, dim), dtype=C)
    rho[0, 0] = 1.0
    rho_t = rho.reshape((2,) * 2 * n_qubits)
    for e in tape:
        name, wires, params, extra = _entry(e)
        if name == "Barrier":
            continue
        if G.is_channel(name):
            acc = np.zeros_like(rho_t)
            for K in G.kraus_matrices(name, params, extra):
                acc = acc + apply_unitary_to_density(rho_t, K, wires, n_qubits)
            rho_t = acc
        else:
            U = G.unitary_matrix(name, wires, params, extra)
            rho_t = apply_unitary_to_density(rho_t, U, wires, n_qubits)
    return rho_t.reshape(dim, dim)


def lifted_matrix(ob, n_qubits):
    """operations.py:402-419: gate applied to every column of the identity."""
    name, wires, params, extra = _entry(ob)
    k = len(wires)
    gt = G.unitary_matrix(name, wires, params, extra).reshape((2,) * 2 * k)
    dim = 2**n_qubits
    eye = np.eye(dim, dtype=C).reshape((dim,) + (2,) * n_qubits)
    sub = einsum_subscript(n_qubits, k, tuple(wires), batched=True)
    cols = np.einsum(sub, gt, eye).reshape(dim, dim)  # row j = O e_j
    return cols.T


def measure_state(state, n_qubits, type, obs=()):
    """simulation.py:204-271 (fast path and dense path give equal values)."""
    if type == "state":
        return state
    if type == "probs":
        return np.abs(state) ** 2
    if type == "expval":
        if n_qubits > 10:
            # the reference's fast path (simulation.py:251-261): 1-qubit diagonal
            # observables from marginal probabilities.  Used here only where the dense
            # lifted matrices (4^n entries) no longer fit; below that the dense path is
            # kept as the independent restatement.
            probs = (np.abs(state) ** 2).reshape((2,) * n_qubits)
            out = []
            for ob in obs:
                name, wires, params, extra = _entry(ob)
                O = G.unitary_matrix(name, wires, params, extra)
                if len(wires) != 1 or np.count_nonzero(O - np.diag(np.diag(O))):
                    raise ValueError("oracle: only 1-qubit diagonal observables beyond 10 qubits")
                marg = probs.sum(axis=tuple(a for a in range(n_qubits) if a != wires[0]))
                out.append(float(np.real(np.diag(O)) @ marg))
            return np.array(out)
        mats = [lifted_matrix(ob, n_qubits) for ob in obs]
        return np.array([np.real(np.conj(state) @ (O @ state)) for O in mats])
    raise ValueError(f"Unknown measurement type: {type!r}")


def measure_density(rho, n_qubits, type, obs=()):
    """simulation.py:274-317."""
    if type == "density":
        return rho
    if type == "probs":
        return np.real(np.diag(rho))
    if type == "expval":
        mats = [lifted_matrix(ob, n_qubits) for ob in obs]
        return np.array([np.real(np.einsum("ij,ji->", O, rho)) for O in mats])
    raise ValueError(
        "Measurement type 'state' is not defined for mixed (noisy) circuits. "
        "Use 'density' instead."
    )


def choice_indices(probs, uniforms):
    """Inverse-CDF sampling as ``jax.random.choice(key, dim, (shots,), p=probs)``
    performs it given the uniform stream ``u`` (jax 0.9 ``_src/random.py`` choice:
    ``p_cuml = cumsum(p); r = p_cuml[-1] * (1 - u); searchsorted(p_cuml, r)``;
    called from simulation.py:352).  The uniform stream is an INPUT.
    """
    p_cuml = np.cumsum(np.asarray(probs))
    r = p_cuml[-1] * (1 - np.asarray(uniforms))
    return np.searchsorted(p_cuml, r, side="left").astype(np.int64)


def sample_shots(probs, n_qubits, type, obs, shots, uniforms):
    """simulation.py:320-377 with the uniforms supplied instead of a PRNG key."""
    dim = 2**n_qubits
    assert len(uniforms) == shots
    samples = choice_indices(probs, uniforms)
    counts = np.zeros(dim, dtype=np.int32)
    np.add.at(counts, np.minimum(samples, dim - 1), 1)  # jnp scatter clamps
    est = counts / shots
    if type == "probs":
        return est, counts
    if type == "expval":
        res = [
            np.real(np.dot(np.diag(lifted_matrix(ob, n_qubits)), est)) for ob in obs
        ]
        return np.array(res), counts
    raise ValueError(
        f"Shot simulation is only supported for 'probs' and 'expval', got {type!r}."
    )


def simulate_and_measure(tape, n_qubits, type, obs=(), use_density=None):
    """simulation.py:131-201 (exact mode)."""
    if use_density is None:
        use_density = uses_density(tape, type)
    if use_density:
        if any(G.is_channel(e[0]) for e in tape):
            rho = simulate_mixed(tape, n_qubits)
        else:
            psi = simulate_pure(tape, n_qubits)
            rho = np.outer(psi, np.conj(psi))  # simulation.py:188-189
        return measure_density(rho, n_qubits, type, obs)
    return measure_state(simulate_pure(tape, n_qubits), n_qubits, type, obs)


def partial_trace(rho, n_qubits, keep):
    """jaqsi.py:60-117 (single or batched)."""
    rho = np.asarray(rho)
    dim = 2**n_qubits
    if rho.shape != (dim, dim):
        return np.stack([partial_trace(r, n_qubits, keep) for r in rho])
    rho_t = rho.reshape((2,) * (2 * n_qubits))
    for q in reversed(sorted(set(range(n_qubits)) - set(keep))):
        n_rem = rho_t.ndim // 2
        rho_t = np.trace(rho_t, axis1=q, axis2=q + n_rem)
    d = 2 ** len(keep)
    return rho_t.reshape(d, d)


def marginalize_probs(probs, n_qubits, keep):
    """jaqsi.py:120-146 (always returns a leading batch axis, like the reference)."""
    dim = 2**n_qubits
    probs = np.asarray(probs).reshape(-1, dim)
    trace_out = tuple(q for q in range(n_qubits - 1, -1, -1) if q not in keep)
    out = []
    for p in probs:
        t = p.reshape((2,) * n_qubits)
        for q in trace_out:
            t = t.sum(axis=q)
        out.append(t.ravel())
    return np.stack(out)


# --------------------------------------------------------------------------
# Batched, reference-faithful execution (the CPU baseline).
# --------------------------------------------------------------------------


def _batched_matrix(name, wires, params, extra, B, xp):
    """Gate matrices for a batch: params entries may be scalars or (B,) arrays."""
    arr = [np.asarray(p, dtype=np.float64) for p in params]
    if not any(a.ndim for a in arr):
        return None
    mats = np.empty((B, 2 ** len(wires), 2 ** len(wires)), dtype=C)
    # closed forms, vectorised over the batch (operations.py:1029-1031,1235-1242,1400-1411)
    def rp(theta, P):
        th = np.broadcast_to(theta, (B,))[:, None, None]
        d = P.shape[0]
        return np.cos(th / 2) * np.eye(d, dtype=C) - 1j * np.sin(th / 2) * P

    if name in ("RX", "RY", "RZ"):
        return rp(arr[0], G.PAULI[name[1]])
    if name == "Rot":
        return rp(arr[2], G.Z) @ rp(arr[1], G.Y) @ rp(arr[0], G.Z)
    if name in ("CRX", "CRY", "CRZ"):
        R = rp(arr[0], G.PAULI[name[2]])
        mats[:] = np.eye(4, dtype=C)
        mats[:, 2:, 2:] = R
        return mats
    if name in ("RXX", "RYY", "RZZ", "RZX"):
        return rp(arr[0], G.pauli_word(name[1:]))
    if name == "PauliRot":
        return rp(arr[0], G.pauli_word(extra))
    if name == "ControlledPhaseShift":
        mats[:] = np.eye(4, dtype=C)
        mats[:, 3, 3] = np.exp(1j * np.broadcast_to(arr[0], (B,)))
        return mats
    for b in range(B):
        pb = [float(np.broadcast_to(a, (B,))[b]) for a in arr]
        mats[b] = G.unitary_matrix(name, wires, pb, extra)
    return mats


def simulate_batched(tape, n_qubits, B, density=False, backend="numpy", threads=None):
    """One einsum per tape op over a (B, 2, ..., 2) tensor.

    This is the vmapped form of simulation.py:100-104 / :124-128: gate parameters
    may be ``(B,)`` arrays (per-element matrices, einsum with a batched gate) or
    scalars (shared matrix).  Returns ``(B, 2**n)`` states or ``(B, 2**n, 2**n)``
    density matrices.  ``backend="torch"`` runs the same contractions with
    torch-CPU complex128 on ``threads`` threads (bench CPU baseline).
    """
    if backend == "torch":
        import torch

        if threads:
            torch.set_num_threads(int(threads))
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
        einsum = torch.einsum
        conj = torch.conj
    else:
        to = lambda a: a  # noqa: E731
        einsum = np.einsum
        conj = np.conj

    rank = 2 * n_qubits if density else n_qubits
    st = np.zeros((B,) + (2,) * rank, dtype=C)
    st.reshape(B, -1)[:, 0] = 1.0
    st = to(st)

    def contract(st, mat, wires_axes, batched_gate):
        k = len(wires_axes)
        sub = einsum_subscript(rank, k, tuple(wires_axes), batched=True)
        if batched_gate:
            sub = "Z" + sub
            mat = mat.reshape((B,) + (2,) * 2 * k)
        else:
            mat = mat.reshape((2,) * 2 * k)
        return einsum(sub, mat, st)

    for e in tape:
        name, wires, params, extra = _entry(e)
        if name == "Barrier":
            continue
        if G.is_channel(name):
            assert density
            bra = [w + n_qubits for w in wires]
            acc = None
            for K in G.kraus_matrices(name, params, extra):
                Kt = to(K)
                t = contract(st, Kt, wires, False)
                t = contract(t, conj(Kt), bra, False)
                acc = t if acc is None else acc + t
            st = acc
            continue
        bm = _batched_matrix(name, wires, params, extra, B, np)
        if bm is None:
            mat = to(G.unitary_matrix(name, wires, [float(p) for p in params], extra))
            batched_gate = False
        else:
            mat = to(bm)
            batched_gate = True
        st = contract(st, mat, wires, batched_gate)
        if density:
            st = contract(st, conj(mat), [w + n_qubits for w in wires], batched_gate)
    out = st.reshape(B, 2**n_qubits, 2**n_qubits) if density else st.reshape(B, -1)
    return out.numpy() if backend == "torch" else out
