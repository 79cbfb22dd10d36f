"""Device-only timing (CUDA events around qmlb_run on staged, HBM-resident arguments) of the
BASELINE configs 1, 3, 4 through the product plan.  Development aid; bench.py carries the
judged numbers.   python tools/probe_dev.py cfg3 cfg4:4096 cfg4f:4096 cfg1"""
import json
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch  # noqa: E402

from qml_essentials_b200 import config, script  # noqa: E402
from qml_essentials_b200.model import Model  # noqa: E402


class Spy:
    def __init__(self, inner):
        self.inner, self.last = inner, None

    def __getattr__(self, k):
        return getattr(self.inner, k)

    def execute(self, plan, host_args, batch, chunk=None, to_host=True):
        self.last = (plan, host_args, batch)
        return self.inner.execute(plan, host_args, batch, chunk, to_host)


def timed_call(model, reps=5, **kw):
    ex = script.get_executor()
    spy = Spy(ex)
    model.script.executor = spy
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model(**kw)
    plan, host_args, batch = spy.last
    call = ex.stage(plan, host_args, batch)
    call.launch()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        call.launch()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    h = call.h
    return min(ts), batch, {"strategy": h.strategy, "passes": h.n_passes}


def main():
    for a in sys.argv[1:]:
        name, _, arg = a.partition(":")
        rng = np.random.default_rng(1000)
        if name == "cfg1":
            config.set_precision("complex128")
            m = Model(2, 1, "Circuit_19")
            kw = dict(params=rng.uniform(0, 2 * np.pi, (1, *m._params_shape)),
                      inputs=np.linspace(-np.pi, np.pi, 1024).reshape(-1, 1))
            types = ["expval"]
        elif name == "cfg3":
            config.set_precision("complex128")
            m = Model(6, 3, "Circuit_15")
            kw = dict(params=rng.uniform(0, 2 * np.pi, (int(arg or 20000), *m._params_shape)))
            types = ["state", "expval", "probs"]
        elif name in ("cfg4", "cfg4f"):
            config.set_precision("complex128" if name == "cfg4" else "complex64")
            m = Model(8, 4, "Strongly_Entangling")
            kw = dict(params=rng.uniform(0, 2 * np.pi, (1, *m._params_shape)),
                      inputs=np.linspace(-np.pi, np.pi, int(arg or 4096)).reshape(-1, 1),
                      noise_params={"Depolarizing": 0.01, "AmplitudeDamping": 0.02})
            types = ["expval", "density"]
        elif name == "sv":  # sv:N:BATCH  statevector circuit of N qubits (frame engine sizes)
            n, _, bs = arg.partition(":")
            config.set_precision("complex128")
            m = Model(int(n), 3, "Hardware_Efficient")
            kw = dict(params=rng.uniform(0, 2 * np.pi, (int(bs or 256), *m._params_shape)))
            types = ["expval"]
        else:
            continue
        for typ in types:
            ms, batch, info = timed_call(m, execution_type=typ, **kw)
            print(json.dumps({"probe": a, "type": typ, "batch": batch, "device_ms": ms,
                              "evals_per_s": batch / (ms * 1e-3), **info}), flush=True)


if __name__ == "__main__":
    main()
