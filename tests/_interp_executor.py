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

    def execute(self, plan, host_args, batch, chunk=None, to_host=True):
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

    # analysis reductions (device kernels qmlb_purity / qmlb_overlap_fidelity in the product)
    def purities(self, states, n_qubits, is_density):
        st = np.asarray(states)
        B, dim = st.shape[0], 2 ** n_qubits
        rho = st.reshape(B, dim, dim) if is_density else np.einsum("bi,bj->bij", st, st.conj())
        out = np.zeros((B, n_qubits))
        t = rho.reshape((B,) + (2,) * (2 * n_qubits))
        for q in range(n_qubits):
            red = np.trace(t, axis1=1 + q, axis2=1 + n_qubits + q)
            red = red.reshape(B, dim // 2, dim // 2)
            out[:, q] = np.real(np.einsum("bij,bji->b", red, red))
        return _Host(out)

    def overlap_fidelities(self, states, n_qubits):
        st = np.asarray(states)
        half = st.shape[0] // 2
        return _Host(np.abs(np.einsum("bi,bi->b", st[:half].conj(), st[half:])) ** 2)


class _Host:
    """numpy array with the ``.cpu().numpy()`` surface of a torch tensor."""

    def __init__(self, a):
        self.a = a

    def cpu(self):
        return self

    def numpy(self):
        return self.a


class InterpShardEngine:
    """Shard-local work of sharded.ShardedExecutor on NumPy (oracle interpreter) and the
    exchange over whatever torch.distributed backend the test initialised (gloo)."""

    def make(self, prog, precision):
        return prog

    def alloc(self, nl, precision):
        return np.zeros(2 ** nl, dtype=np.complex128)

    def stage_args(self, host_args):
        return [(a[0], a[1], a[2]) if a is not None else (None, 1, 1) for a in host_args]

    def evolve(self, prog, staged, state, init_mode):
        if init_mode:
            state[:] = 0
            if init_mode == 1:
                state[0] = 1.0
        return pi.Interp(prog, staged, 1).run(state[None, :])[0]

    def exchange(self, state, g):
        """all_to_all_single semantics (chunk s -> rank s) emulated with all_gather:
        gloo has no alltoall."""
        import torch
        import torch.distributed as dist

        size, rank = dist.get_world_size(), dist.get_rank()
        t = torch.from_numpy(np.ascontiguousarray(state).view(np.float64).copy())
        every = [torch.empty_like(t) for _ in range(size)]
        dist.all_gather(every, t)
        return np.concatenate([e.numpy().view(np.complex128).reshape(size, -1)[rank]
                               for e in every])

    def zsums(self, state, nl):
        p = np.abs(state) ** 2
        idx = np.arange(2 ** nl)
        out = np.zeros(33)
        for q in range(nl):
            out[q] = p[(idx >> q) & 1 == 1].sum()
        out[32] = p.sum()
        return out

    def to_host(self, state):
        return state
