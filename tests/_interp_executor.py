"""Test-only executor: runs compiled programs through the oracle's NumPy
interpreter so the CPU suite can exercise recording, compilation, batching and
the Model API without a GPU.  Installed via ``script._set_executor_for_testing``
by the ``cpu_executor`` fixture; never reachable from the product."""

import numpy as np

from oracle import program_interp as pi
from oracle import sim as osim

_NAMES = {0: "state", 1: "probs", 2: "expval", 3: "density"}


class InterpExecutor:
    name = "oracle-interp"

    def _states(self, plan, host_args, batch):
        args = [(a[0], a[1], a[2]) if a is not None else (None, 1, 1) for a in host_args]
        return pi.Interp(plan.program, args, batch).run()

    def execute(self, plan, host_args, batch, chunk=None):
        st = self._states(plan, host_args, batch)
        out = pi.measure(plan.program, st, _NAMES[plan.out_type], plan.obs_recs,
                         plan.obs_pool)
        if plan.precision == "complex64":
            out = out.astype(np.complex64 if np.iscomplexobj(out) else np.float32)
        return out

    def execute_shots(self, plan, host_args, batch, uniforms, chunk=None):
        st = self._states(plan, host_args, batch)
        probs = pi.measure(plan.program, st, "probs")
        dim = probs.shape[1]
        counts = np.zeros((batch, dim), dtype=np.int32)
        for b in range(batch):
            idx = np.minimum(osim.choice_indices(probs[b], uniforms[b]), dim - 1)
            np.add.at(counts[b], idx, 1)
        return counts
