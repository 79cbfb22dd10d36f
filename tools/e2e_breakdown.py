"""Where the end-to-end time of Coefficients.get_spectrum goes (development aid)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
from qml_essentials_b200.coefficients import Coefficients  # noqa: E402

model, params, inputs = bench.workload()
model.params = params
for _ in range(5):
    Coefficients.get_spectrum(model, mfs=8, shift=True, trim=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    Coefficients.get_spectrum(model, mfs=8, shift=True, trim=True)
print("get_spectrum ms", (time.perf_counter() - t0) / 20 * 1e3)
t0 = time.perf_counter()
for _ in range(20):
    model(params=params, inputs=inputs)
print("model.__call__ ms", (time.perf_counter() - t0) / 20 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    Coefficients.get_spectrum(model, mfs=8, shift=True, trim=True)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
