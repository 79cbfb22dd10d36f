"""Times the BASELINE configs 3-5 (or scaled-down versions) on one GPU through the
product API and prints one JSON line per probe.  Development aid, not the bench.

    python tools/probe_scale.py [cfg3] [cfg4:BATCH] [cfg5:NQUBITS] ...
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch  # noqa: E402

from qml_essentials_b200 import config, script  # noqa: E402
from qml_essentials_b200.model import Model  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def info(model):
    plan = [p for p in model.script._jit_cache.values() if hasattr(p, "program") and p.device]
    h = list(plan[-1].device.values())[0]
    return {"strategy": h.strategy, "passes": h.n_passes, "device_ops": h.n_device_ops}


def cfg3(batch=20000):
    config.set_precision("complex128")
    m = Model(n_qubits=6, n_layers=3, circuit_type="Circuit_15")
    rng = np.random.default_rng(1000)
    params = rng.uniform(0, 2 * np.pi, (batch, *m._params_shape))
    for typ in ("density", "state", "probs"):
        t = timed(lambda: m(params=params, execution_type=typ))
        print(json.dumps({"probe": "cfg3", "type": typ, "batch": batch, "s": t,
                          "evals_per_s": batch / t, **info(m)}), flush=True)


def cfg4(batch=256, precision="complex128"):
    config.set_precision(precision)
    m = Model(n_qubits=8, n_layers=4, circuit_type="Strongly_Entangling")
    rng = np.random.default_rng(1000)
    params = rng.uniform(0, 2 * np.pi, (1, *m._params_shape))
    x = np.linspace(-np.pi, np.pi, batch).reshape(-1, 1)
    noise = {"Depolarizing": 0.01, "AmplitudeDamping": 0.02}
    for typ in ("expval", "density"):
        t = timed(lambda: m(params=params, inputs=x, noise_params=noise, execution_type=typ),
                  reps=2)
        print(json.dumps({"probe": "cfg4", "precision": precision, "type": typ, "batch": batch,
                          "s": t, "evals_per_s": batch / t, **info(m)}), flush=True)


def cfg5(n=26):
    config.set_precision("complex64")
    m = Model(n_qubits=n, n_layers=8, circuit_type="Hardware_Efficient")
    rng = np.random.default_rng(1000)
    params = rng.uniform(0, 2 * np.pi, (1, *m._params_shape))
    x = np.array([[0.5]])
    t = timed(lambda: m(params=params, inputs=x, execution_type="expval"), reps=2)
    i = info(m)
    state_bytes = 8 * 2**n
    print(json.dumps({"probe": "cfg5", "n": n, "s": t, **i,
                      "gbps_per_pass": 2 * state_bytes * i["passes"] / t / 1e9}), flush=True)


if __name__ == "__main__":
    for a in sys.argv[1:] or ["cfg3", "cfg4:64", "cfg5:24"]:
        name, _, arg = a.partition(":")
        try:
            if name == "cfg3":
                cfg3(int(arg) if arg else 20000)
            elif name == "cfg4":
                cfg4(int(arg) if arg else 256)
            elif name == "cfg4f":
                cfg4(int(arg) if arg else 256, "complex64")
            elif name == "cfg5":
                cfg5(int(arg) if arg else 26)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"probe": a, "error": f"{type(e).__name__}: {e}"}), flush=True)
