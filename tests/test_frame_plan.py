"""Host planner of the on-chip frame engine (strategy 3), checked WITHOUT a GPU: the step
program dumped by ``qmlb_plan_describe`` is executed by a NumPy emulation of ``k_frame``
(tests/_frame_emulator.py) and compared with the plain program interpreter - which the
rest of the CPU suite pins to the reference-faithful oracle."""

import warnings

import numpy as np
import pytest

import _frame_emulator as fe
from _interp_executor import InterpExecutor
from qml_essentials_b200 import backend
from qml_essentials_b200.model import Model


@pytest.fixture(scope="module")
def lib():
    return backend.load_library()


def _both(lib, n, L, ct, typ, noise=None, precision="complex128", B_I=2, B_P=2, **kw):
    ex = fe.FrameEmuExecutor(lib)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = Model(n, L, ct, precision=precision, **kw)
        rng = np.random.default_rng(1)
        params = rng.uniform(0, 2 * np.pi, (B_P, *m._params_shape))
        inputs = rng.uniform(-1, 1, (B_I, 1))
        call = lambda: np.asarray(m(params=params, inputs=inputs, execution_type=typ,
                                    noise_params=dict(noise) if noise else None))
        m.script.executor = ex
        got = call()
        m.script.executor = InterpExecutor()
        want = call()
    return ex, float(np.abs(got - want).max())


NOISE = {"Depolarizing": 0.01, "AmplitudeDamping": 0.02}


@pytest.mark.parametrize("n,L,ct,typ,noise", [
    (6, 3, "Circuit_15", "density", None),            # BASELINE config 3 circuit
    (6, 2, "Circuit_19", "expval", None),             # controlled rotations: parity-row controls
    (7, 2, "Hardware_Efficient", "probs", None),
    (9, 1, "Strongly_Entangling", "state", None),
    (3, 2, "Strongly_Entangling", "density", NOISE),  # 4x4 superoperators on (ket, bra)
    (4, 2, "Strongly_Entangling", "expval", NOISE),
    (4, 1, "Circuit_6", "probs", {"BitFlip": 0.1, "MultiQubitDepolarizing": 0.05}),  # 16x16
    (6, 1, "Circuit_9", "expval", None),              # H + CZ: diagonal ops on parity rows
])
def test_single_cta_schedules(lib, monkeypatch, n, L, ct, typ, noise):
    monkeypatch.setenv("QMLB_PTM", "0")  # the complex engine; strategy 5 has its own tests
    ex, err = _both(lib, n, L, ct, typ, noise)
    assert ex.frame_runs == 1, "planner did not choose the frame engine"
    assert err < 1e-12


@pytest.mark.parametrize("n,L,ct,typ,noise,precision", [
    (8, 1, "Strongly_Entangling", "expval", NOISE, "complex128"),  # config 4: cluster of 8
    (7, 1, "Strongly_Entangling", "probs", {"Depolarizing": 0.01}, "complex128"),
    (14, 1, "Hardware_Efficient", "expval", None, "complex128"),
    (15, 1, "Circuit_19", "probs", None, "complex64"),
])
def test_cluster_schedules(lib, monkeypatch, n, L, ct, typ, noise, precision):
    """States beyond one CTA's shared memory: outer (cluster-rank) bits, relayouts through
    distributed shared memory, CX folded with an outer control."""
    monkeypatch.setenv("QMLB_PTM", "0")
    ex, err = _both(lib, n, L, ct, typ, noise, precision, B_I=1, B_P=1)
    assert ex.frame_runs == 1
    geo, steps = fe.parse(ex.steps[0])
    assert geo["outer_bits"] >= 1
    assert any(s[0] == "relayout" for s in steps)
    assert err < (1e-12 if precision == "complex128" else 1e-9)


def test_config4_schedule_shape(lib, monkeypatch):
    """BASELINE config 4 (8 qubits, 4 layers, depolarizing + amplitude damping) in the COMPLEX
    engine (density output; <Z> / probs go through strategy 5): 472 tape
    ops -> at most 100 sub-passes and 12 coalesced cluster exchanges (each preceded by a
    tile-local shuffle); every CX is folded (no permutation op reaches the kernel)."""
    monkeypatch.setenv("QMLB_PTM", "0")
    ex, err = _both(lib, 8, 4, "Strongly_Entangling", "expval", NOISE, B_I=1, B_P=1)
    geo, steps = fe.parse(ex.steps[0])
    assert (geo["tile_bits"], geo["outer_bits"], geo["threads"]) == (13, 3, 256)
    n_sub = sum(1 for s in steps if s[0] == "subpass")
    n_exch = sum(1 for s in steps if s[0] == "relayout" and not s[2])
    n_local = sum(1 for s in steps if s[0] == "relayout" and s[2])
    assert n_sub <= 100 and n_exch <= 12 and n_local <= 14
    # every (ket, bra) superoperator sits on a canonical register pair (kernel fast path)
    for s in steps:
        if s[0] == "subpass":
            assert all((o["j0"], o["j1"]) in ((3, 2), (1, 0)) for o in s[4] if o["code"] == 1)
            assert s[5] >= 16  # straight-line fast path
    assert err < 1e-12


@pytest.mark.parametrize("n,L,ct,precision", [
    (12, 2, "Hardware_Efficient", "complex128"),
    (14, 3, "Hardware_Efficient", "complex128"),
    (13, 2, "Circuit_19", "complex128"),
    (12, 2, "Strongly_Entangling", "complex128"),
    (13, 1, "Circuit_9", "complex128"),
])
def test_streamed_tile_schedules(lib, monkeypatch, n, L, ct, precision):
    """Strategy 4 (state in HBM, tile passes): with the tile shrunk to 2^10 amplitudes the
    emulator can run whole schedules - HBM bit layout per pass, bits that stay / arrive /
    leave between passes, final bit map of the <Z> sweep."""
    monkeypatch.setenv("QMLB_FSTREAM_TILE_BITS", "10")
    monkeypatch.setenv("QMLB_FSTREAM_LOW_BITS", "5")
    monkeypatch.setenv("QMLB_FRAME", "0")
    ex, err = _both(lib, n, L, ct, "expval", None, precision, B_I=1, B_P=2)
    assert ex.frame_runs == 1 and ex.steps[0].startswith("strategy 4")
    geo, steps = fe.parse(ex.steps[0])
    assert len(geo["passes"]) >= 2 and sorted(geo["final_hpos"]) == list(range(n))
    assert sum(p["steps"] for p in geo["passes"]) == len(steps)
    assert err < 1e-12


def test_config5_pass_count(lib):
    """BASELINE config 5 (32 qubits, 8 layers Hardware_Efficient, complex64): 1 408 tape
    gates -> at most 20 HBM passes (the register-group stream of round 1 needed 77), every
    step on a straight-line fast path."""
    import test_cabi

    plan = test_cabi._plan_of(32, 8, "Hardware_Efficient", "complex64")
    text = backend.plan_describe(lib, plan.program, plan.out_type, plan.obs_recs, plan.obs_pool,
                                 "complex64")
    assert text.startswith("strategy 4")
    geo, steps = fe.parse(text)
    assert len(geo["passes"]) <= 20
    assert all(s[5] >= 64 for s in steps if s[0] == "subpass")
    assert all(s[2] for s in steps if s[0] == "relayout")


@pytest.mark.parametrize("n,L,ct,precision", [
    (32, 8, "Hardware_Efficient", "complex64"), (28, 4, "Hardware_Efficient", "complex128"),
    (20, 2, "Circuit_19", "complex64"), (22, 2, "Strongly_Entangling", "complex64"),
])
def test_streamed_tiles_stage_one_pass_of_matrices(lib, n, L, ct, precision):
    """Strategy 4 stages only the matrices of the current HBM pass next to the tile: inside a
    pass the shared-memory ranges of the ops are disjoint and fit ``mat_cap``, every op keeps
    its own offset in the element's row of evaluated matrices, and tile + matrices + two
    step records leave room for three CTAs per SM at the default tile size."""
    import test_cabi

    plan = test_cabi._plan_of(n, L, ct, precision)
    text = backend.plan_describe(lib, plan.program, plan.out_type, plan.obs_recs, plan.obs_pool,
                                 precision)
    assert text.startswith("strategy 4")
    geo, steps = fe.parse(text)
    cs = 16 if precision == "complex128" else 8
    assert geo["mat_cap"] % 16 == 0 and geo["mat_cap"] <= geo["premat_row"] + 15
    assert geo["smem"] == (cs << geo["tile_bits"]) + geo["mat_cap"] * cs + 2 * 1024 + 320 * 4 + 64
    assert 3 * (geo["smem"] + 1024) <= 227 * 1024
    seen_rows = set()
    for ps in geo["passes"]:
        used = []
        for st in steps[ps["first"]:ps["first"] + ps["steps"]]:
            if st[0] != "subpass":
                continue
            for o in st[4]:
                if o["code"] == 5:  # sign op: no matrix
                    continue
                # diagonal: 2^k entries, controlled 2x2: the 2x2 alone, dense: 4^k
                size = (1 << o["k"]) if o["code"] == 4 else (4 if o["code"] == 3 else 1 << (2 * o["k"]))
                used.append((o["smem_off"], o["smem_off"] + size))
                assert o["premat_off"] not in seen_rows
                seen_rows.add(o["premat_off"])
        used.sort()
        if not used:  # a pass that only brings the bits home
            continue
        assert used[0][0] == 0 and used[-1][1] <= geo["mat_cap"]
        assert all(a[1] == b[0] for a, b in zip(used, used[1:]))  # compact, no overlap


@pytest.mark.parametrize("n,L,ct,typ,noise", [
    (3, 2, "Strongly_Entangling", "expval", NOISE),
    (4, 2, "Hardware_Efficient", "probs", NOISE),
    (5, 1, "Circuit_15", "expval", {"BitFlip": 0.05, "PhaseDamping": 0.1}),
    (4, 1, "Circuit_6", "expval", {"PhaseFlip": 0.1, "ThermalRelaxation": None}),
    (8, 1, "Strongly_Entangling", "probs", NOISE),
    (4, 2, "Strongly_Entangling", "density", NOISE),
])
def test_pauli_basis_schedules(lib, n, L, ct, typ, noise):
    """Strategy 5: noisy density programs made of 1-qubit (ket, bra) ops and CX pairs evolve
    as the REAL Pauli-coefficient vector - transfer matrices T S T^-1, the CX as a folded
    linear map on the (x, z) bits times the Aaronson-Gottesman sign.  The emulation converts
    the final coefficients back to rho and compares with the interpreter's complex
    evolution."""
    noise = {k: v for k, v in noise.items() if v is not None}
    ex, err = _both(lib, n, L, ct, typ, noise, B_I=2 if n < 8 else 1, B_P=2 if n < 8 else 1)
    assert ex.frame_runs == 1
    geo, steps = fe.parse(ex.steps[0])
    if ct == "Circuit_6":  # controlled rotations: not a Clifford + 1-qubit program
        assert not geo.get("ptm")
    else:
        assert ex.steps[0].startswith("strategy 5") and geo["ptm"] == 1
        assert any(o["code"] == fe.FOP_SIGN for s in steps if s[0] == "subpass" for o in s[4])
    assert err < 1e-12


def test_relayout_lane_tables(lib, monkeypatch):
    """Tile-local relayouts of the complex engines: the planner's lane table makes gather and
    store conflict-free in the bank model (consecutive destinations: 2-3.6x the ideal)."""
    ex, err = _both(lib, 14, 3, "Hardware_Efficient", "expval", None, B_I=1, B_P=1)
    geo, _ = fe.parse(ex.steps[0])
    assert geo["relayouts"] and fe.relayout_wavefronts(geo, 16, swizzled=False) < 1.1
    for r in geo["relayouts"]:
        r["htab"] = [0] * 32
    assert fe.relayout_wavefronts(geo, 16, swizzled=False) > 1.5
    assert err < 1e-12


def test_config4_pauli_basis_shape(lib):
    """BASELINE config 4 with <Z> output: 512 KiB of real coefficients per evaluation -> a
    cluster of 4 CTAs (the complex form needs 8), 512 threads with two items each."""
    ex, err = _both(lib, 8, 4, "Strongly_Entangling", "expval", NOISE, B_I=1, B_P=1)
    geo, steps = fe.parse(ex.steps[0])
    assert (geo["ptm"], geo["tile_bits"], geo["outer_bits"], geo["threads"]) == (1, 14, 2, 512)
    # item maps of the swizzled tile: bijections, and (nearly) conflict-free wavefronts
    fe.check_item_maps(geo)
    assert fe.smem_wavefronts(geo, 8) < 1.05
    assert sum(1 for s in steps if s[0] == "relayout" and not s[2]) <= 10
    assert err < 1e-12
