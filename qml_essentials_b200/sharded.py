"""Qubit-sharded statevector across GPUs (large-n regime).

The reference cannot hold a state beyond host RAM (every ``einsum`` allocates a fresh
copy, simulation.py:103) - there is nothing to mirror; this is the new scaling axis.

Layout.  With ``G = 2^g`` ranks (one process per GPU) a state of ``n`` bits is split by
its top ``g`` *physical* bits: rank ``r`` holds the ``2^(n-g)`` amplitudes whose physical
bits ``n-g .. n-1`` spell ``r``.  A permutation ``pos`` maps the program's logical state
bits to physical positions, so a global<->local qubit swap is data movement plus a
relabelling - the program itself is never rewritten.

Execution.  The compiled program (the same flat program the single-GPU path runs) is cut
into *epochs*: maximal runs of ops whose bits are all local under the current ``pos``.
An epoch is one streaming program on ``n-g`` bits (``qmlb_evolve``: the k_stream fused
gate passes, in place on the shard).  When an op needs a bit that is currently global:

1. the ``g`` local logical bits whose next use is farthest away are moved to the top
   ``g`` local positions by local SWAPs (2-bit permutations fused into the epoch's
   passes - register renaming inside k_stream);
2. ONE exchange swaps physical bits ``n-2g .. n-g-1`` with ``n-g .. n-1``: chunk ``s`` of
   the shard goes to rank ``s`` and the chunk received from rank ``s`` lands at chunk
   ``s`` - exactly ``all_to_all_single`` with equal splits over NCCL / NVSwitch
   ((G-1)/G of the shard leaves each GPU);
3. ``pos`` is updated and the scan continues.

Measurement.  ``qmlb_zsums`` gives every rank, in one sweep, the probability mass with
each local bit set and its total; bits that are global contribute the rank's total where
the rank index has that bit set.  One all-reduce of ``n + 1`` doubles yields every
``<Z_q>`` on every rank.

Scope this round: statevector programs, batch 1, single-qubit Z expectation values
(BASELINE config 5) and, for tests at small n, the gathered state.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import replace
from typing import List, Optional, Tuple

import numpy as np

from . import compiler, parallel
from .compiler import OP_DIAG, OP_PERM, Program

_SWAP_PERM = (0.0, 2.0, 1.0, 3.0)  # new[v] = old[perm[v]] exchanges the two bits
_X_PERM = (1.0, 0.0)
_CX_CONTROL_FIRST = (0, 1, 3, 2)     # CX with control = bits[0]
_CX_CONTROL_SECOND = (0, 3, 2, 1)    # CX with control = bits[1]


def plan_epochs(prog: Program, g: int, rank: int = 0):
    """Cut ``prog.ops`` into epochs of local ops separated by exchanges.

    Returns ``(steps, pos, consts)``: ``steps`` is a list of ``("ops", ndarray of
    OP_DTYPE with PHYSICAL local bits)`` and ``("exchange",)`` entries, ``pos`` the final
    logical -> physical map, ``consts`` the constant pool extended by the SWAP table.

    An op whose GLOBAL bits are all controls needs no exchange: on this rank the control
    value is a constant (a bit of ``rank``), so CX / controlled 2x2 gates with a global
    control become a plain X / 2x2 on the target or nothing at all.  The exchange schedule
    does not depend on ``rank`` (every rank cuts the epochs at the same ops); only the ops
    emitted for such gates do."""
    n = prog.n_bits
    nl = n - g
    if g < 0 or nl < g:
        raise ValueError(f"cannot shard {n} state bits over 2^{g} ranks")
    consts = np.concatenate([np.asarray(prog.consts, dtype=np.float64), _SWAP_PERM, _X_PERM])
    swap_aux = len(prog.consts)
    x_aux = swap_aux + len(_SWAP_PERM)
    pos = list(range(n))
    ops = prog.ops
    bits_of = [list(int(b) for b in o["bits"][: o["k"]]) for o in ops]

    def control_target(i):
        """(control bit, target bit) of a CX / controlled 2x2, else None."""
        o = ops[i]
        if int(o["kind"]) == compiler.OP_CTRL1:
            return bits_of[i][0], bits_of[i][1]
        if int(o["kind"]) == OP_PERM and int(o["k"]) == 2:
            perm = tuple(int(v) for v in prog.consts[int(o["aux"]): int(o["aux"]) + 4])
            if perm == _CX_CONTROL_FIRST:
                return bits_of[i][0], bits_of[i][1]
            if perm == _CX_CONTROL_SECOND:
                return bits_of[i][1], bits_of[i][0]
        return None

    ct_of = [control_target(i) for i in range(len(ops))]

    # next use of every logical bit at or after op index i (for the victim choice)
    INF = len(ops) + 1
    next_use = np.full((len(ops) + 1, n), INF, dtype=np.int64)
    for i in range(len(ops) - 1, -1, -1):
        next_use[i] = next_use[i + 1]
        for b in bits_of[i]:
            # a control does not have to be local (see above): only targets count as uses
            if ct_of[i] is None or b != ct_of[i][0]:
                next_use[i, b] = i

    steps: List[tuple] = []
    cur: List[tuple] = []

    def emit(kind, k, src, aux, bits):
        cur.append((kind, k, src, aux, list(bits) + [0] * (compiler.MAX_OP_BITS - len(bits))))

    def flush():
        nonlocal cur
        arr = np.zeros(len(cur), dtype=compiler.OP_DTYPE)
        for j, rec in enumerate(cur):
            arr[j] = rec
        steps.append(("ops", arr))
        cur = []

    for i, o in enumerate(ops):
        lb = bits_of[i]
        if g and ct_of[i] is not None and pos[ct_of[i][0]] >= nl and pos[ct_of[i][1]] < nl:
            ctl, tgt = ct_of[i]
            if (rank >> (pos[ctl] - nl)) & 1:  # control reads 1 on this rank
                if int(o["kind"]) == OP_PERM:
                    emit(OP_PERM, 1, -1, x_aux, [pos[tgt]])
                else:
                    emit(compiler.OP_MAT, 1, int(o["src"]), 0, [pos[tgt]])
            continue
        if g and any(pos[b] >= nl for b in lb):
            if len(lb) > nl - g:
                raise ValueError("operation too wide for the local register after an exchange")
            # victims: local logical bits not used by this op, farthest next use first
            local_logical = [b for b in range(n) if pos[b] < nl and b not in lb]
            local_logical.sort(key=lambda b: (-next_use[i, b], -pos[b]))
            victims = local_logical[:g]
            top = list(range(nl - g, nl))
            inv = {pos[b]: b for b in range(n)}
            free_top = [p for p in top if inv[p] not in victims]
            for v in victims:
                if pos[v] >= nl - g:
                    continue
                tgt = free_top.pop()
                other = inv[tgt]
                emit(OP_PERM, 2, -1, swap_aux, [pos[v], tgt])
                pos[v], pos[other] = tgt, pos[v]
                inv[pos[v]], inv[pos[other]] = v, other
            flush()
            steps.append(("exchange",))
            inv = {pos[b]: b for b in range(n)}
            for j in range(g):
                a, b2 = inv[nl - g + j], inv[nl + j]
                pos[a], pos[b2] = nl + j, nl - g + j
        emit(int(o["kind"]), int(o["k"]), int(o["src"]), int(o["aux"]), [pos[b] for b in lb])
    flush()
    return steps, pos, consts


def epoch_program(prog: Program, ops: np.ndarray, consts: np.ndarray, nl: int) -> Program:
    """The sub-program of one epoch: same sources / angles / constants, local ops."""
    return replace(prog, n_qubits=nl, n_bits=nl, density=False, ops=ops, consts=consts,
                   meta={})


# ---------------------------------------------------------------------------------------
class CudaShardEngine:
    """Shard-local work through the C ABI.

    Exchange, two forms:

    * fused (default when torch symmetric memory can map the peers' shards): the shard
      lives in two symmetric buffers; after a device-side barrier the first gate pass of
      the next epoch reads its amplitudes straight from the peers over NVLink
      (``qmlb_evolve_peer``) and writes the other buffer - no separate exchange step, no
      extra HBM round trip, transfer overlapped with the arithmetic;
    * NCCL ``all_to_all_single`` into a second buffer, then the epoch in place (fallback,
      and the reference point for the fused form).
    """

    def __init__(self, executor=None, fused: Optional[bool] = None):
        from . import script

        self.ex = executor or script.get_executor()
        self.torch = self.ex.torch
        self.lib = self.ex.lib
        self.want_fused = fused
        self.fused_exchange = False
        self._symm = {}   # (nl, precision) -> (buffers, handles)
        self._cur = None  # (key, index of the buffer that holds the state)

    def _symmetric_buffers(self, nl: int, precision: str):
        key = (nl, precision)
        if key in self._symm:
            return self._symm[key]
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        real = self.torch.float64 if precision == "complex128" else self.torch.float32
        bufs, hdls = [], []
        for _ in range(2):
            t = symm.empty(2 * 2 ** nl, dtype=real, device=self.ex.device)
            hdls.append(symm.rendezvous(t, dist.group.WORLD))
            bufs.append(self.torch.view_as_complex(t.view(-1, 2)))
        self._symm[key] = (bufs, hdls)
        return self._symm[key]

    def make(self, prog: Program, precision: str):
        from .backend import QMLB_DESC_FORCE_STREAM, ProgramHandle

        empty_obs = np.zeros(0, dtype=compiler.OBS_DTYPE)
        with self.torch.cuda.device(self.ex.device):
            return ProgramHandle(self.lib, prog, compiler.OUT_STATE, empty_obs,
                                 np.zeros(0, dtype=np.float64), precision,
                                 flags=QMLB_DESC_FORCE_STREAM)

    def _use_fused(self) -> bool:
        """Measured (profiles/r1_qubit_sharded_fused_vs_nccl_*): with 2 ranks the fused
        form is 14 % faster end to end (n = 31); with 4 ranks, where 3/4 of the first
        pass' reads are remote, it is 8 % slower than all_to_all + local pass (n = 32) -
        the gate-pass kernel reaches only ~360 GB/s of peer reads.  Until that is fixed
        the fused form is the default for 2 ranks and opt-in (QMLB_SHARD_FUSED=1 or
        ``fused=True``) beyond."""
        import os

        size = parallel.world()[1]
        if self.want_fused is not None:
            return bool(self.want_fused) and size in (2, 4, 8)
        env = os.environ.get("QMLB_SHARD_FUSED")
        if env is not None:
            return env not in ("0", "") and size in (2, 4, 8)
        return size == 2

    def alloc(self, nl: int, precision: str):
        if self._use_fused():
            try:
                bufs, _ = self._symmetric_buffers(nl, precision)
                self.fused_exchange = True
                self._cur = ((nl, precision), 0)
                return bufs[0]
            except Exception as exc:  # noqa: BLE001
                if self.want_fused:
                    raise
                import logging

                logging.getLogger(__name__).info(
                    "symmetric memory unavailable (%s); exchanging through NCCL", exc)
        self.fused_exchange = False
        dt = self.torch.complex128 if precision == "complex128" else self.torch.complex64
        return self.torch.empty(2 ** nl, dtype=dt, device=self.ex.device)

    def exchange_evolve(self, handle, staged, state):
        """Fused global<->local swap + first epoch pass (+ the rest of the epoch)."""
        key, cur = self._cur
        bufs, hdls = self._symm[key]
        rank, size = parallel.world()
        hdls[cur].barrier()  # device side, stream ordered: every rank finished writing `cur`
        dst = bufs[1 - cur]
        dev, c_args, n = staged
        ptrs = (C.c_void_p * size)(*[int(p) for p in hdls[cur].buffer_ptrs])
        ws = self.torch.empty(1 << 20, dtype=self.torch.uint8, device=self.ex.device)
        rc = self.lib.qmlb_evolve_peer(
            handle.ptr, c_args, n, dst.data_ptr(), ptrs, size, rank, ws.data_ptr(), ws.numel(),
            self.torch.cuda.current_stream(self.ex.device).cuda_stream)
        if rc != 0:
            from .backend import BackendError

            raise BackendError(f"qmlb_evolve_peer: {self.lib.qmlb_last_error().decode()}")
        self._cur = (key, 1 - cur)
        return dst

    def stage_args(self, host_args):
        from .backend import _Arg

        dev = self.ex.to_device(host_args)
        c_args = (_Arg * max(len(dev), 1))()
        for i, a in enumerate(dev):
            c_args[i] = _Arg(None, 0, 1, 1) if a is None else _Arg(
                a[0].data_ptr(), a[0].shape[1], int(a[1]), int(a[2]))
        return dev, c_args, len(dev)

    def evolve(self, handle, staged, state, init_mode: int):
        dev, c_args, n = staged
        ws_bytes = int(self.lib.qmlb_workspace_bytes(handle.ptr, c_args, n, 1))
        # state-size part of the estimate is not needed: the shard is evolved in place
        ws = self.torch.empty(max(min(ws_bytes, 1 << 20), 1), dtype=self.torch.uint8,
                              device=self.ex.device)
        rc = self.lib.qmlb_evolve(handle.ptr, c_args, n, 1, 0, state.data_ptr(), init_mode,
                                  ws.data_ptr(), ws.numel(),
                                  self.torch.cuda.current_stream(self.ex.device).cuda_stream)
        if rc != 0:
            from .backend import BackendError

            raise BackendError(f"qmlb_evolve: {self.lib.qmlb_last_error().decode()}")
        return state

    def exchange(self, state, g: int):
        import torch.distributed as dist

        out = self.torch.empty_like(state)
        dist.all_to_all_single(out, state)
        return out

    def zsums(self, state, nl: int) -> np.ndarray:
        from .backend import QMLB_C64, QMLB_C128

        dt = QMLB_C128 if state.dtype == self.torch.complex128 else QMLB_C64
        nb = int(self.lib.qmlb_zsums_workspace_bytes(1, nl))
        ws = self.torch.empty(max(nb, 1), dtype=self.torch.uint8, device=self.ex.device)
        out = self.torch.empty(33, dtype=self.torch.float64, device=self.ex.device)
        rc = self.lib.qmlb_zsums(state.data_ptr(), dt, 1, nl, out.data_ptr(), ws.data_ptr(), nb,
                                 self.torch.cuda.current_stream(self.ex.device).cuda_stream)
        if rc != 0:
            from .backend import BackendError

            raise BackendError(f"qmlb_zsums: {self.lib.qmlb_last_error().decode()}")
        return out.cpu().numpy()

    def to_host(self, state) -> np.ndarray:
        return state.cpu().numpy()

    def sync(self):
        self.torch.cuda.synchronize(self.ex.device)


class ShardedExecutor:
    """Drop-in for ``Script.executor``: runs one statevector circuit split over all ranks
    of the initialised ``torch.distributed`` group and returns, on every rank, the
    single-qubit Z expectation values (or the gathered state for small ``n``)."""

    name = "cuda-sm100a-qubit-sharded"

    def __init__(self, engine=None):
        self.engine = engine
        self._default_engine = None
        self.stats = {}

    def _world(self) -> Tuple[int, int, int]:
        rank, size = parallel.world()
        g = int(np.log2(size))
        if 2 ** g != size:
            raise ValueError("qubit sharding needs a power-of-two number of ranks")
        return rank, size, g

    def execute(self, plan, host_args, batch: int, chunk: Optional[int] = None,
                to_host: bool = True):
        if batch != 1 or plan.program.density:
            raise ValueError("qubit sharding runs one statevector circuit at a time")
        if self.engine is None and self._default_engine is None:
            self._default_engine = CudaShardEngine()
        eng = self.engine or self._default_engine
        rank, size, g = self._world()
        prog = plan.program
        n, nl = prog.n_bits, prog.n_bits - g
        # epochs and their device programs are built once per plan and rank count
        key = ("sharded", size, id(eng))
        cached = plan.device.get(key)
        if cached is None:
            steps, pos, consts = plan_epochs(prog, g, rank)
            compiled = []
            for st in steps:
                if st[0] == "exchange":
                    compiled.append(("exchange", None))
                elif len(st[1]) or not compiled:
                    compiled.append(("ops", eng.make(epoch_program(prog, st[1], consts, nl),
                                                     plan.precision)))
            cached = plan.device[key] = (compiled, pos)
        compiled, pos = cached
        staged = eng.stage_args(host_args)

        import time

        state = eng.alloc(nl, plan.precision)
        first, n_exchange, n_epochs = True, 0, 0
        t_ex = 0.0
        timed = getattr(eng, "sync", None)
        fused = bool(getattr(eng, "fused_exchange", False))
        skip = False
        for idx, (kind, handle) in enumerate(compiled):
            if skip:
                skip = False
                continue
            if kind == "exchange" and fused and idx + 1 < len(compiled) \
                    and compiled[idx + 1][0] == "ops":
                state = eng.exchange_evolve(compiled[idx + 1][1], staged, state)
                n_exchange += 1
                n_epochs += 1
                skip = True
                continue
            if kind == "exchange":
                if timed:
                    timed()
                    t0 = time.perf_counter()
                state = eng.exchange(state, g)
                if timed:
                    timed()
                    t_ex += time.perf_counter() - t0
                n_exchange += 1
                continue
            init = (1 if rank == 0 else 2) if first else 0
            state = eng.evolve(handle, staged, state, init)
            first = False
            n_epochs += 1
        self.stats = {"exchanges": n_exchange, "epochs": n_epochs, "local_bits": nl,
                      "ranks": size, "exchange_seconds": t_ex,
                      "exchange": "fused peer loads (qmlb_evolve_peer)" if fused
                      else "all_to_all_single", "bytes_sent_per_exchange":
                      (size - 1) * (2 ** nl // size) * (16 if plan.precision == "complex128"
                                                        else 8)}

        if plan.out_type == compiler.OUT_EXPVAL:
            return self._expvals(eng, plan, state, pos, rank, g, nl)
        if plan.out_type == compiler.OUT_STATE:
            return self._gather(eng, state, pos, n, nl)
        raise ValueError("qubit sharding supports 'expval' (Z observables) and 'state'")

    # -- measurement -------------------------------------------------------------------
    def _expvals(self, eng, plan, state, pos, rank, g, nl):
        n = plan.program.n_bits
        sums = eng.zsums(state, nl)  # [bit q set mass for q < nl ..., total at 32]
        contrib = np.zeros(n + 1, dtype=np.float64)
        contrib[n] = sums[32]
        for logical in range(n):
            p = pos[logical]
            if p < nl:
                contrib[logical] = sums[p]
            elif (rank >> (p - nl)) & 1:
                contrib[logical] = sums[32]
        tot = parallel.allreduce_sum(contrib)
        out = np.zeros((1, len(plan.obs_recs)))
        for j, ob in enumerate(plan.obs_recs):
            zmask = int(ob["zmask"])
            if ob["kind"] != compiler.OBS_ZSTRING or bin(zmask).count("1") != 1:
                raise ValueError("qubit sharding measures single-qubit Z observables")
            out[0, j] = tot[n] - 2.0 * tot[zmask.bit_length() - 1]
        if plan.precision == "complex64":
            out = out.astype(np.float32)
        return out

    def _gather(self, eng, state, pos, n, nl):
        """Full state in logical bit order on every rank (tests / small n only)."""
        local = np.ascontiguousarray(eng.to_host(state))
        full = parallel.allgather_concat(local[None, :], axis=0).reshape(-1)  # physical order
        idx = np.arange(2 ** n, dtype=np.int64)
        phys = np.zeros_like(idx)
        for logical in range(n):
            phys |= ((idx >> logical) & 1) << pos[logical]
        return full[phys][None, :]
