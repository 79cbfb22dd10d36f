"""Known-answer tests ported from the reference's own suite (analytic values only).
Each function takes ``run(circuit_fn, n_qubits, type, obs, args=())`` so the same
assertions pin (a) the oracle and (b) the product on CPU-interpreter and GPU."""

import numpy as np

from qml_essentials_b200 import operations as op


def bell():
    op.H(wires=0)
    op.CX(wires=[0, 1])


def ghz3():
    op.H(wires=0)
    op.CX(wires=[0, 1])
    op.CX(wires=[1, 2])


def ghz4():
    ghz3()
    op.CX(wires=[2, 3])


def ghz_toffoli():
    op.H(wires=0)
    op.CX(wires=[0, 1])
    op.CCX(wires=[0, 1, 2])


def run_all(run, atol=1e-10):
    Z, X = op.PauliZ, op.PauliX
    # reference tests/test_jaqsi.py:352-363 Bell expectation values
    for cls in (X, Z):
        r = run(bell, 2, "expval", [cls(0, record=False), cls(1, record=False)])
        assert np.allclose(r, [0, 0], atol=atol)
    # :365-370
    assert np.allclose(run(bell, 2, "probs", []), [0.5, 0, 0, 0.5], atol=atol)
    # :372-380 RX(theta) -> cos(theta)
    for th in (0.5, 0.7, 2.1):
        r = run(lambda t: op.RX(t, wires=0), 1, "expval", [Z(0, record=False)],
                args=(np.array(th),))
        assert np.allclose(r[0], np.cos(th), atol=max(atol, 1e-6) if atol > 1e-8 else atol)
    # :382-404 GHZ
    e = np.zeros(8); e[[0, 7]] = 0.5
    assert np.allclose(run(ghz3, 3, "probs", []), e, atol=atol)
    assert np.allclose(run(ghz3, 3, "expval", [Z(q, record=False) for q in range(3)]),
                       0, atol=atol)
    e4 = np.zeros(16); e4[[0, 15]] = 0.5
    assert np.allclose(run(ghz4, 4, "probs", []), e4, atol=atol)
    # :406-413 Toffoli
    assert np.allclose(run(ghz_toffoli, 3, "probs", []), e, atol=atol)
    # :415-427 CX on non-adjacent wires populates |000> and |101>

    def nonadj():
        op.H(wires=0)
        op.CX(wires=[0, 2])

    e5 = np.zeros(8); e5[[0, 5]] = 0.5
    assert np.allclose(run(nonadj, 3, "probs", []), e5, atol=atol)
    # :429-491 density invariants
    rho = run(bell, 2, "density", [])
    assert rho.shape == (4, 4)
    assert np.allclose(np.trace(rho), 1, atol=atol)
    assert np.allclose(rho @ rho, rho, atol=atol)
    psi = run(bell, 2, "state", [])
    assert np.allclose(rho, np.outer(psi, psi.conj()), atol=atol)
    rho3 = run(ghz3, 3, "density", [])
    assert np.allclose(np.real(np.diag(rho3)), run(ghz3, 3, "probs", []), atol=atol)
    rho4 = run(ghz4, 4, "density", [])
    assert np.allclose(rho4, rho4.conj().T, atol=atol)
    r1 = run(lambda t: op.RX(t, wires=0), 1, "density", [], args=(np.array(0.7),))
    assert np.allclose(np.real(np.trace(np.diag([1, -1]) @ r1)), np.cos(0.7), atol=1e-8)

    # :664-696 noise auto-routing and validity
    def noisy():
        op.H(wires=0)
        op.BitFlip(0.1, wires=0)

    p = run(noisy, 1, "probs", [])
    assert p.shape == (2,) and np.allclose(p.sum(), 1, atol=atol)

    def noisy_bell():
        bell()
        op.DepolarizingChannel(0.05, wires=0)
        op.DepolarizingChannel(0.05, wires=1)

    rho = run(noisy_bell, 2, "density", [])
    assert np.allclose(np.trace(rho), 1, atol=atol)
    assert np.allclose(rho, rho.conj().T, atol=atol)
    pur = np.real(np.trace(rho @ rho))
    assert pur < 1 - 1e-6

    # reference tests/test_ansaetze.py:82-160 channel limit values
    def rx_pi(ch=None):
        op.RX(np.pi, wires=0)
        if ch is not None:
            ch()

    zo = [Z(0, record=False)]
    assert np.allclose(run(rx_pi, 1, "expval", zo), -1, atol=max(atol, 1e-6))
    assert np.allclose(run(lambda: rx_pi(lambda: op.BitFlip(0.5, wires=0)), 1, "expval", zo),
                       0, atol=max(atol, 1e-6))
    assert np.allclose(
        run(lambda: rx_pi(lambda: op.DepolarizingChannel(0.75, wires=0)), 1, "expval", zo),
        0, atol=max(atol, 1e-6))

    def h_pf(p):
        op.H(wires=0)
        if p:
            op.PhaseFlip(p, wires=0)

    xo = [X(0, record=False)]
    assert np.allclose(run(lambda: h_pf(0), 1, "expval", xo), 1, atol=max(atol, 1e-6))
    assert np.allclose(run(lambda: h_pf(0.5), 1, "expval", xo), 0, atol=max(atol, 1e-6))

    def two(p):
        from qml_essentials_b200.unitary import UnitaryGates

        op.RX(np.pi, wires=0)
        op.CRX(np.pi, wires=[0, 1])
        if p:
            UnitaryGates.NQubitDepolarizingChannel(p, [0, 1])

    z1 = [Z(1, record=False)]
    assert np.allclose(run(lambda: two(0), 2, "expval", z1), -1, atol=max(atol, 1e-6))
    # the reference asserts ~0 within 0.1 (tests/test_ansaetze.py:152-154); the exact
    # value is -(1 - p*15/16 - p/16) + ... = -(1 - p) = -1/16 for p = 15/16
    r = run(lambda: two(15 / 16), 2, "expval", z1)
    assert np.allclose(r, 0, atol=0.1)
    assert np.allclose(r, -1 / 16, atol=max(atol, 1e-6))


def run_pennylane_conventions(run, atol=1e-10):
    """The circuits of the reference's PennyLane-equality tests (tests/test_jaqsi.py:494-661)
    against the CLOSED FORMS that follow from PennyLane's documented gate / channel
    definitions (the reference asserts equality with ``default.qubit`` / ``default.mixed``;
    PennyLane is absent here, its definitions are standard).  Independent of the oracle's
    matrix code: amplitudes are written out by hand."""
    c2 = lambda t: np.cos(t / 2) ** 2
    s2 = lambda t: np.sin(t / 2) ** 2

    def ctrl(gate, theta, both):
        def f():
            op.H(wires=0)
            if both:
                op.H(wires=1)
            if theta is None:
                gate(wires=[0, 1])
            else:
                gate(theta, wires=[0, 1])
        return f

    # :496-536 controlled gates on H|0> (x) |0>  (CZ: H on both)
    want = {
        "CY": (op.CY, None, False, [0.5, 0, 0, 0.5]),
        "CZ": (op.CZ, None, True, [0.25] * 4),
        "CRX": (op.CRX, 1.3, False, [0.5, 0, 0.5 * c2(1.3), 0.5 * s2(1.3)]),
        "CRY": (op.CRY, 0.9, False, [0.5, 0, 0.5 * c2(0.9), 0.5 * s2(0.9)]),
        "CRZ": (op.CRZ, 2.1, False, [0.5, 0, 0.5, 0]),
    }
    for name, (g, th, both, probs) in want.items():
        assert np.allclose(run(ctrl(g, th, both), 2, "probs", []), probs, atol=atol), name
    # CRZ phases (state, not only probabilities): diag(1, 1, e^{-i t/2}, e^{+i t/2})
    st = run(ctrl(op.CRZ, 2.1, False), 2, "state", [])
    assert np.allclose(st, [2 ** -0.5, 0, 2 ** -0.5 * np.exp(-1.05j), 0], atol=atol)
    # :539-555 Rot(phi, theta, omega) = RZ(omega) RY(theta) RZ(phi)
    r = run(lambda: op.Rot(0.4, 1.2, 2.5, wires=0), 1, "probs", [])
    assert np.allclose(r, [c2(1.2), s2(1.2)], atol=atol)
    st = run(lambda: op.Rot(0.4, 1.2, 2.5, wires=0), 1, "state", [])
    assert np.allclose(st, [np.exp(-0.5j * (0.4 + 2.5)) * np.cos(0.6),
                            np.exp(-0.5j * (0.4 - 2.5)) * np.sin(0.6)], atol=atol)

    # :589-639 one channel after RX(theta)|0> = cos(t/2)|0> - i sin(t/2)|1>
    def rho0(t):
        c, s = np.cos(t / 2), np.sin(t / 2)
        return np.array([[c * c, 1j * c * s], [-1j * c * s, s * s]])

    X = np.array([[0, 1], [1, 0]], dtype=complex)
    Y = np.array([[0, -1j], [1j, 0]])
    Z = np.diag([1.0 + 0j, -1.0])

    def after(ch, p, t):
        def f(theta):
            op.RX(theta, wires=0)
            ch(p, wires=0)
        return run(f, 1, "density", [], args=(np.array(t),))

    r0 = rho0(0.8)
    assert np.allclose(after(op.BitFlip, 0.15, 0.8), 0.85 * r0 + 0.15 * X @ r0 @ X, atol=atol)
    r0 = rho0(1.1)
    assert np.allclose(after(op.PhaseFlip, 0.2, 1.1), 0.8 * r0 + 0.2 * Z @ r0 @ Z, atol=atol)
    r0 = rho0(0.6)
    dep = 0.88 * r0 + 0.04 * (X @ r0 @ X + Y @ r0 @ Y + Z @ r0 @ Z)
    assert np.allclose(dep, (1 - 0.16) * r0 + 0.16 * np.eye(2) / 2, atol=1e-14)
    assert np.allclose(after(op.DepolarizingChannel, 0.12, 0.6), dep, atol=atol)
    r0, g = rho0(1.3), 0.25
    ad = np.array([[r0[0, 0] + g * r0[1, 1], np.sqrt(1 - g) * r0[0, 1]],
                   [np.sqrt(1 - g) * r0[1, 0], (1 - g) * r0[1, 1]]])
    assert np.allclose(after(op.AmplitudeDamping, g, 1.3), ad, atol=atol)
    r0, g = rho0(0.9), 0.3
    pdm = np.array([[r0[0, 0], np.sqrt(1 - g) * r0[0, 1]], [np.sqrt(1 - g) * r0[1, 0], r0[1, 1]]])
    assert np.allclose(after(op.PhaseDamping, g, 0.9), pdm, atol=atol)

    # :642-661 thermal relaxation, T2 <= T1 (PennyLane's six-operator form, pe = 0)
    pe, t1, t2, tg = 0.0, 1e-4, 5e-5, 1e-6
    e1, e2 = np.exp(-tg / t1), np.exp(-tg / t2)
    p_reset = 1 - e1
    pz = (1 - p_reset) * (1 - e2 / e1) / 2
    pr0 = (1 - pe) * p_reset
    pid = 1 - pz - pr0
    r0 = rho0(1.0)
    tr = np.array([[(pid + pz) * r0[0, 0] + pr0 * (r0[0, 0] + r0[1, 1]), (pid - pz) * r0[0, 1]],
                   [(pid - pz) * r0[1, 0], (pid + pz) * r0[1, 1]]])

    def thermal(theta):
        op.RX(theta, wires=0)
        op.ThermalRelaxationError(pe, t1, t2, tg, wires=0)

    got = run(thermal, 1, "density", [], args=(np.array(1.0),))
    assert np.allclose(got, tr, atol=atol)
