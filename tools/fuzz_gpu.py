"""Randomised GPU-vs-interpreter sweep (development aid): every ansatz x qubit count x depth x
noise mix x output type through the CUDA executor and through the oracle's program
interpreter on the SAME compiled plan inputs.  Prints the worst cases.
    python tools/fuzz_gpu.py [n_cases] [seed]"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.join(os.path.dirname(__file__), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from _interp_executor import InterpExecutor  # noqa: E402
from qml_essentials_b200.ansaetze import Ansaetze  # noqa: E402
from qml_essentials_b200.model import Model  # noqa: E402

NOISES = [None, None, {"Depolarizing": 0.02}, {"AmplitudeDamping": 0.05, "PhaseDamping": 0.03},
          {"BitFlip": 0.02, "PhaseFlip": 0.04, "Depolarizing": 0.01},
          {"AmplitudeDamping": 0.03, "Measurement": 0.02, "StatePreparation": 0.01},
          {"MultiQubitDepolarizing": 0.02, "BitFlip": 0.01}]


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
    big = len(sys.argv) > 3 and sys.argv[3] == "big"  # statevectors of 12..19 qubits
    names = [c.__name__ for c in Ansaetze.get_available()]
    names = [c for c in names if c not in ("No_Ansatz", "GHZ")]  # no parameters to batch
    worst = []
    fails = 0
    for case in range(n_cases):
        ct = names[rng.integers(len(names))]
        noise = None if big else NOISES[rng.integers(len(NOISES))]
        nmax = 8 if noise else 13
        n = int(rng.integers(12, 20)) if big else int(rng.integers(2, nmax + 1))
        L = int(rng.integers(1, 3 if big else 4))
        typ = ["expval", "probs", "density" if (noise or n <= 7) else "state", "state"][rng.integers(4)]
        if noise and typ == "state":
            typ = "density"
        prec = ["complex128", "complex64"][rng.integers(2)]
        B_I, B_P = int(rng.integers(1, 6)), int(rng.integers(1, 5))
        if big:
            B_I, B_P = int(rng.integers(1, 3)), 1
            typ = ["expval", "probs", "state"][rng.integers(3)]
        elif n <= 5 and not noise:  # register kernel: batch shapes that cut its CTA tiles
            B_I, B_P = int(rng.integers(1, 41)), int(rng.integers(1, 41))
        tag = f"{ct} n={n} L={L} {typ} {prec} noise={sorted(noise) if noise else None} B=({B_I},{B_P})"
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                m = Model(n, L, ct, precision=prec)
                params = rng.uniform(0, 2 * np.pi, (B_P, *m._params_shape))
                inputs = rng.uniform(-1, 1, (B_I, 1))
                kw = dict(params=params, inputs=inputs, execution_type=typ,
                          noise_params=dict(noise) if noise else None)
                got = np.asarray(m(**kw))
                m2 = Model(n, L, ct, precision=prec)
                m2.script.executor = InterpExecutor()
                kw["noise_params"] = dict(noise) if noise else None
                want = np.asarray(m2(**kw))
            err = float(np.abs(got - want).max())
        except Exception as exc:  # noqa: BLE001
            print("EXC", tag, type(exc).__name__, str(exc)[:200], flush=True)
            fails += 1
            continue
        tol = 1e-10 if prec == "complex128" else 2e-5
        worst.append((err / tol, err, tag))
        if err > tol:
            fails += 1
            print("FAIL", f"{err:.3e}", tag, flush=True)
    worst.sort(reverse=True)
    for r, e, t in worst[:6]:
        print(f"worst {e:.3e} ({r:.2f} x tol) {t}")
    print(f"cases {n_cases} failures {fails}")


main()
