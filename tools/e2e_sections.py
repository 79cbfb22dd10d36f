"""Wall-clock split of one Coefficients.get_spectrum call (no profiler): cumulative time
inside the main host-side functions, averaged over many calls.  Development aid."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
from qml_essentials_b200 import backend, coefficients, model as model_mod, script  # noqa: E402
from qml_essentials_b200.coefficients import Coefficients  # noqa: E402

acc = {}


def wrap(obj, name, label=None):
    fn = getattr(obj, name)
    label = label or f"{getattr(obj, '__name__', type(obj).__name__)}.{name}"

    def inner(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            acc[label] = acc.get(label, 0.0) + time.perf_counter() - t0

    setattr(obj, name, inner)


model, params, inputs = bench.workload()
model.params = params
for _ in range(5):
    Coefficients.get_spectrum(model, mfs=8, shift=True, trim=True)
ex = script.get_executor()
wrap(coefficients.Coefficients, "_device_spectrum") if hasattr(coefficients.Coefficients, "_device_spectrum") else None
wrap(model_mod.Model, "_forward")
wrap(script.Script, "execute")
wrap(script.Script, "_execute_batched")
wrap(script.Script, "_signature")
wrap(script.Script, "_device_args")
wrap(backend.CudaExecutor, "execute")
wrap(backend.CudaExecutor, "stage")
wrap(backend.CudaExecutor, "to_device")
wrap(backend.CudaExecutor, "grid_dft")
wrap(backend.CudaExecutor, "to_host")
wrap(backend.DeviceCall, "launch")
wrap(backend.DeviceCall, "__init__", "DeviceCall.__init__")
N = 200
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(N):
    Coefficients.get_spectrum(model, mfs=8, shift=True, trim=True)
total = time.perf_counter() - t0
print(f"get_spectrum {total / N * 1e3:.3f} ms per call")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print(f"  {k:40s} {v / N * 1e3:.3f} ms")
