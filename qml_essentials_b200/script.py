"""Circuit container and executor: the backend seam.

Same public contract as the reference's ``qml_essentials/script.py``:
``Script(f, n_qubits=None)`` (script.py:84), ``execute(type, obs, *, args, kwargs,
in_axes, shots, key)`` (script.py:137-147), ``_record`` (script.py:97) and a plan
cache keyed on the call signature (script.py:469-542).  Underneath, instead of
``jax.vmap`` + ``jit`` over einsum kernels:

record once with affine proxies  ->  compile to a flat device program
(:mod:`.compiler`)  ->  one CUDA launch sequence over the whole batch through the
C ABI (:mod:`.backend`, ``include/qmlb200.h``).

There is no CPU execution path: without the CUDA library and a GPU,
``execute`` raises :class:`~.backend.BackendUnavailable`.
"""

from __future__ import annotations

import hashlib
import logging
from dataclasses import dataclass, field
from typing import Any, Callable, List, Optional, Tuple, Union

import numpy as np

from . import compiler, config, memory, rng, simulation
from .operations import KrausChannel, Operation
from .symbolic import NonAffineProduct, Sym, SymArray, SymbolicError
from .tape import pulse_recording, recording
from .unitary import UnitaryGates

log = logging.getLogger(__name__)

_OUT_TYPES = {
    "state": compiler.OUT_STATE,
    "probs": compiler.OUT_PROBS,
    "expval": compiler.OUT_EXPVAL,
    "density": compiler.OUT_DENSITY,
}


class BatchAxis:
    """``in_axes`` entry for a *factor* of the batch: element ``b`` of the flat
    batch reads row ``(b // div) % mod`` of the argument along ``axis``.

    A plain int ``ax`` is ``BatchAxis(ax, 1, B)``.  ``Model`` uses factors to run
    the (inputs x params x pulse) product batch (model.py:1414-1483) without
    materialising the repeated arrays.
    """

    __slots__ = ("axis", "div", "mod", "total")

    def __init__(self, axis: int, div: int, mod: int, total: int):
        self.axis, self.div, self.mod, self.total = int(axis), int(div), int(mod), int(total)

    def _key(self):
        return ("BatchAxis", self.axis, self.div, self.mod, self.total)

    def __repr__(self):
        return f"BatchAxis(axis={self.axis}, div={self.div}, mod={self.mod})"


class LazyKeys:
    """``split(parent, num)`` that is only evaluated if the recorded circuit
    actually draws random numbers (GateError)."""

    def __init__(self, parent: rng.PRNGKey, num: int):
        self.parent, self.num = parent, int(num)
        self._keys = None

    @property
    def shape(self):
        return (self.num,)

    def materialize(self) -> rng.PRNGKey:
        if self._keys is None:
            self._keys = rng.split(self.parent, self.num)
        return self._keys


def _make_hashable(obj):
    """dict/list/set -> nested tuples for cache keys (script.py:14-28)."""
    if isinstance(obj, dict):
        return tuple(sorted((k, _make_hashable(v)) for k, v in obj.items()))
    if isinstance(obj, (list, tuple)):
        return tuple(_make_hashable(x) for x in obj)
    if isinstance(obj, set):
        return frozenset(_make_hashable(x) for x in obj)
    if isinstance(obj, np.ndarray):
        return ("ndarray", obj.shape, str(obj.dtype), hashlib.sha1(
            np.ascontiguousarray(obj).tobytes()).hexdigest())
    return obj


def _is_float_array(a) -> bool:
    return isinstance(a, np.ndarray) and a.dtype.kind in "fiu" and a.dtype != bool


def _as_array(a):
    """Accept numpy arrays, python scalars/lists of numbers and torch tensors."""
    if isinstance(a, np.ndarray):
        return a
    if hasattr(a, "detach") and hasattr(a, "cpu"):  # torch tensor
        return a.detach().cpu().numpy()
    if isinstance(a, (float, np.floating)):
        return np.asarray(a, dtype=np.float64)
    return a


@dataclass
class _Plan:
    """Compiled artefacts for one call signature (the reference's ``_BatchPlan``,
    script.py:31-53)."""

    program: Any
    out_type: int
    obs_recs: np.ndarray
    obs_pool: np.ndarray
    n_qubits: int
    use_density: bool  # reference flag: density requested or noisy tape
    density_program: bool  # a 4^n state is evolved
    n_ops: int
    slots: List[tuple]  # per device-arg slot: ("pos", i) | ("noise",) | ("table",) | None
    noise: Optional[rng.NoiseRecorder]
    baked: Tuple[int, ...]
    table: Optional[np.ndarray] = None
    precision: str = "complex128"
    device: dict = field(default_factory=dict)  # executor-owned cache (program handle)


_EXECUTOR = None


def get_executor():
    """The process-wide executor: the CUDA library behind the C ABI."""
    global _EXECUTOR
    if _EXECUTOR is None:
        from .backend import CudaExecutor

        _EXECUTOR = CudaExecutor()
    return _EXECUTOR


def _set_executor_for_testing(executor) -> None:
    """Test hook (tests/ only): lets the CPU suite drive the host logic with the
    oracle's program interpreter.  The product never calls this."""
    global _EXECUTOR
    _EXECUTOR = executor


class Script:
    """Records a circuit function and executes it on the GPU."""

    def __init__(self, f: Callable[..., None], n_qubits: Optional[int] = None,
                 precision: Optional[str] = None) -> None:
        self.f = f
        self._n_qubits = n_qubits
        self.precision = precision
        # hashable extra cache-key component for state the circuit function reads
        # besides its arguments (Model: zero-input shortcut, re-upload mask, ...)
        self.cache_salt = None
        # optional executor for this Script only (e.g. sharded.ShardedExecutor); None ->
        # the process-wide CUDA executor
        self.executor = None
        self._jit_cache: dict = {}

    # -- recording ------------------------------------------------------------
    def _record(self, *args, **kwargs) -> List[Operation]:
        """Run the circuit function, return the recorded operations (script.py:97-115)."""
        with recording() as tape:
            self.f(*args, **kwargs)
        return tape

    def pulse_events(self, *args, **kwargs) -> list:
        with pulse_recording() as events:
            with recording():
                self.f(*args, **kwargs)
        return events

    # -- public entry -----------------------------------------------------------
    def execute(
        self,
        type: str = "expval",
        obs: Optional[List[Operation]] = None,
        *,
        args: tuple = (),
        kwargs: Optional[dict] = None,
        in_axes: Optional[Tuple] = None,
        shots: Optional[int] = None,
        key=None,
        device_result: bool = False,
    ):
        """Execute the circuit and measure (script.py:137-219).

        ``device_result=True`` (extension, used by the analysis callers) returns the
        batched result as a device-resident ``torch`` tensor instead of copying it to
        the host, so reductions (purities, pair fidelities) can run on the GPU.

        ``type``: ``"expval"`` | ``"probs"`` | ``"state"`` | ``"density"``.  Without
        ``in_axes`` the result has the bare measurement shape; with ``in_axes`` (one
        entry per positional argument, ``jax.vmap`` convention: int batch axis or
        ``None`` to broadcast) it gains a leading batch axis.
        """
        obs = list(obs) if obs is not None else []
        kwargs = kwargs or {}
        if shots is not None and key is None:
            key = rng.key(0)  # script.py:189-190
        args = tuple(_as_array(a) for a in args)
        batched = in_axes is not None
        if batched and len(in_axes) != len(args):
            raise ValueError(
                f"in_axes has {len(in_axes)} entries but args has {len(args)}. "
                "Provide one in_axes entry per positional argument."
            )
        if not batched:
            in_axes = (None,) * len(args)
        batch = self._batch_size(args, in_axes)
        result = self._execute_batched(type, obs, args, kwargs, tuple(in_axes), batch,
                                       shots, key, device_result)
        return result if batched else result[0]

    # -- helpers ----------------------------------------------------------------
    @staticmethod
    def _batch_size(args: tuple, in_axes: Tuple) -> int:
        for a, ax in zip(args, in_axes):
            if isinstance(ax, BatchAxis):
                return ax.total
        for a, ax in zip(args, in_axes):
            if ax is not None:
                return int(a.shape[ax])
        return 1

    def _precision(self) -> str:
        return self.precision or config.get_precision()

    def _signature(self, type, obs, args, kwargs, in_axes, shots, baked):
        sig = []
        for i, (a, ax) in enumerate(zip(args, in_axes)):
            # the program does not depend on batch sizes or factors, only on the axis
            axk = ("B", ax.axis) if isinstance(ax, BatchAxis) else (
                None if ax is None else ("B", ax))
            if _is_float_array(a):
                if i in baked:
                    sig.append((axk, _make_hashable(a)))
                elif isinstance(ax, BatchAxis) or ax is not None:
                    shp = list(a.shape)
                    shp.pop(ax.axis if isinstance(ax, BatchAxis) else ax)
                    sig.append((axk, tuple(shp), "f"))
                else:
                    sig.append((axk, a.shape, "f"))
            elif isinstance(a, (rng.PRNGKey, LazyKeys)):
                sig.append((axk, "key"))
            else:
                sig.append((axk, _make_hashable(a)))
        # observables with a class-level constant matrix (PauliZ, ...) are identified by
        # name and wires; only free-form ones (Hermitian, parity) hash their matrix
        obs_sig = tuple(
            (o.name, tuple(o.wires)) if getattr(o.__class__, "_matrix", None) is not None
            else (o.name, tuple(o.wires), _make_hashable(np.asarray(o.matrix)))
            for o in obs
        )
        return (
            type,
            tuple(sig),
            _make_hashable(kwargs),
            UnitaryGates.batch_gate_error,  # script.py:472-475
            self._precision(),
            obs_sig,
            ("shots", shots) if shots is not None else None,
            self.cache_salt,
        )

    # -- planning ---------------------------------------------------------------
    def _symbolic_args(self, args, in_axes, baked):
        sym = []
        for i, (a, ax) in enumerate(zip(args, in_axes)):
            axis = ax.axis if isinstance(ax, BatchAxis) else ax
            if _is_float_array(a) and a.size > 0 and i not in baked:
                shape = list(a.shape)
                if axis is not None:
                    shape.pop(axis)
                sym.append(SymArray.leaves(i, tuple(shape)) if shape else Sym.leaf(i, 0))
            elif isinstance(a, (rng.PRNGKey, LazyKeys)):
                sym.append(rng.SymKey(i))  # batched or broadcast: drawn at run time
            elif _is_float_array(a) and axis is not None:
                sym.append(np.take(a, 0, axis=axis))  # empty or baked batched arg
            else:
                sym.append(a)
        return sym

    def _build_plan(self, type, obs, args, kwargs, in_axes, batch, for_shots) -> _Plan:
        """Record once, compile once (replaces script.py:272-329)."""
        noise_slot, table_slot = len(args), len(args) + 1
        broadcast = {i for i, ax in enumerate(in_axes) if ax is None}
        baked: set = set()
        tape = rec = None
        fallback = False
        while True:
            try:
                with rng.noise_recording(noise_slot) as rec:
                    tape = self._record(*self._symbolic_args(args, in_axes, baked), **kwargs)
                n_qubits = self._n_qubits or simulation.infer_n_qubits(tape, obs)
                noisy = simulation.has_noise(tape)
                program = compiler.compile_tape(tape, n_qubits, density=noisy)
                break
            except NonAffineProduct as e:
                cands = [a for a in e.args_involved if a in broadcast and a not in baked]
                if cands:
                    baked.add(max(cands))
                    continue
                fallback = True
                break
            except (SymbolicError, TypeError):
                # anything an affine proxy cannot follow (np.cos(theta), comparisons,
                # ...); a genuine TypeError resurfaces in the concrete re-recording
                fallback = True
                break

        table = None
        if fallback:
            log.info("circuit is not affine in its batched arguments; recording per element")
            tapes = [self._record(*self._element_args(args, in_axes, b), **kwargs)
                     for b in range(batch)]
            tape = tapes[0]
            n_qubits = self._n_qubits or simulation.infer_n_qubits(tape, obs)
            noisy = simulation.has_noise(tape)
            program, table = compiler.compile_elementwise(tapes, n_qubits, noisy, table_slot)
            rec = None

        use_density = simulation.uses_density(tape, type)
        if type == "state" and noisy:
            raise ValueError(
                "Measurement type 'state' is not defined for mixed (noisy) circuits. "
                "Use 'density' instead."
            )
        if type not in _OUT_TYPES:
            raise ValueError(f"Unknown measurement type: {type!r}")
        for o in obs:
            if max(o.wires) >= n_qubits:
                raise ValueError(f"observable {o.name} acts outside {n_qubits} qubits")
        obs_recs, obs_pool = compiler.compile_observables(obs, n_qubits)

        slots: List[Optional[tuple]] = [None] * program.n_args
        for s in range(program.n_args):
            if s < len(args):
                slots[s] = ("pos", s)
            elif s == noise_slot and rec is not None and rec.n_cols:
                slots[s] = ("noise",)
            elif s == table_slot:
                slots[s] = ("table",)
        out_type = _OUT_TYPES["probs" if for_shots else type]
        return _Plan(
            program=program, out_type=out_type, obs_recs=obs_recs, obs_pool=obs_pool,
            n_qubits=n_qubits, use_density=use_density, density_program=noisy,
            n_ops=len(tape), slots=slots,
            noise=rec if (rec is not None and rec.n_cols) else None,
            baked=tuple(sorted(baked)), table=table, precision=self._precision(),
        )

    @staticmethod
    def _element_args(args, in_axes, b):
        out = []
        for a, ax in zip(args, in_axes):
            if ax is None:
                out.append(a.materialize() if isinstance(a, LazyKeys) else a)
                continue
            if isinstance(ax, BatchAxis):
                idx, axis = (b // ax.div) % ax.mod, ax.axis
            else:
                idx, axis = b, ax
            if isinstance(a, LazyKeys):
                a = a.materialize()
            if isinstance(a, rng.PRNGKey):
                out.append(a[idx])
            else:
                out.append(np.take(a, idx, axis=axis))
        return out

    def _device_args(self, plan: _Plan, args, in_axes, batch):
        """Per slot ``(array2d float64, div, mod)`` or ``None``."""
        out = []
        for slot in plan.slots:
            if slot is None:
                out.append(None)
            elif slot[0] == "pos":
                i = slot[1]
                a, ax = args[i], in_axes[i]
                if not _is_float_array(a) or i in plan.baked:
                    out.append(None)
                    continue
                if ax is None:
                    out.append((np.ascontiguousarray(a, dtype=np.float64).reshape(1, -1), 1, 1))
                else:
                    axis = ax.axis if isinstance(ax, BatchAxis) else ax
                    rows = np.ascontiguousarray(np.moveaxis(a, axis, 0), dtype=np.float64)
                    rows = rows.reshape(rows.shape[0], -1)
                    if isinstance(ax, BatchAxis):
                        out.append((rows, ax.div, ax.mod))
                    else:
                        out.append((rows, 1, rows.shape[0]))
            elif slot[0] == "noise":
                keys = {}
                for (arg, _path, _n) in plan.noise.recipes:
                    k = args[arg]
                    k = k.materialize() if isinstance(k, LazyKeys) else k
                    ax = in_axes[arg]
                    if ax is None:
                        k = rng.PRNGKey(np.broadcast_to(k.data, (batch, 2)))
                    elif isinstance(ax, BatchAxis):
                        idx = (np.arange(batch) // ax.div) % ax.mod
                        k = rng.PRNGKey(k.data[idx])
                    keys[arg] = k
                out.append((plan.noise.realise(keys, batch), 1, batch))
            elif slot[0] == "table":
                out.append((plan.table, 1, batch))
        return out

    # -- execution ----------------------------------------------------------------
    def _chunk_size(self, cache_key, plan: _Plan, type: str, n_obs: int, batch: int,
                    ex=None, host_args=None) -> int:
        """Largest chunk that fits in HBM, memoised per batch size (script.py:331-356).
        With a device executor the figure is exact (the library reports the workspace of
        the strategy it planned); the arithmetic model of ``memory.py`` is the fallback."""
        mem_key = ("_mem", cache_key, batch)
        chunk = self._jit_cache.get(mem_key)
        if chunk is None:
            if ex is not None and hasattr(ex, "peak_bytes") and host_args is not None:
                avail = int(memory.available_memory_bytes() * 0.8)
                full = ex.peak_bytes(plan, host_args, batch)
                if full <= avail:
                    chunk = batch
                else:
                    probe = min(batch, 1024)
                    per = max(1, -(-ex.peak_bytes(plan, host_args, probe) // probe))
                    chunk = int(max(1, min(batch, avail // per)))
                    memory.log.info(
                        f"Computation requires ~{full / 1024**3:.2f} GB which does not fit in "
                        f"~{avail / 1024**3:.2f} GB. Using chunk size {chunk}.")
            else:
                chunk = memory.compute_chunk_size(
                    plan.n_qubits, batch, type, plan.density_program, n_obs, n_ops=plan.n_ops
                )
            self._jit_cache[mem_key] = chunk
        return chunk

    def _execute_batched(self, type, obs, args, kwargs, in_axes, batch, shots=None, key=None,
                         device_result=False):
        for_shots = shots is not None and type in ("probs", "expval")
        # the set of baked (value-keyed) arguments is discovered on the first build
        probe_key = ("_baked", type, tuple(
            (("B", ax.axis) if isinstance(ax, BatchAxis) else ax) for ax in in_axes),
            _make_hashable({k: v for k, v in kwargs.items()}), len(args), self.cache_salt)
        baked = self._jit_cache.get(probe_key, ())
        cache_key = self._signature(type, obs, args, kwargs, in_axes,
                                    shots if for_shots else None, baked)
        plan = self._jit_cache.get(cache_key)
        if plan is None:
            plan = self._build_plan(type, obs, args, kwargs, in_axes, batch, for_shots)
            if plan.baked != tuple(baked):
                self._jit_cache[probe_key] = plan.baked
                cache_key = self._signature(type, obs, args, kwargs, in_axes,
                                            shots if for_shots else None, plan.baked)
            if plan.table is None:  # per-element fallback plans depend on the values
                self._jit_cache[cache_key] = plan

        ex = self.executor or get_executor()
        host_args = self._device_args(plan, args, in_axes, batch)
        chunk = self._chunk_size(cache_key, plan, "probs" if for_shots else type,
                                 len(obs), batch, ex, host_args)
        if for_shots:
            keys = rng.split(key, batch)  # script.py:481
            uniforms = rng.choice_uniforms(keys, int(shots))
            counts = ex.execute_shots(plan, host_args, batch, uniforms, chunk)
            est = counts / shots  # simulation.py:357
            if type == "probs":
                return est
            diags = np.stack([_lifted_diag(o, plan.n_qubits) for o in obs])
            return np.real(est @ diags.T)  # simulation.py:367-372
        if device_result:
            return ex.execute(plan, host_args, batch, chunk, to_host=False)
        return ex.execute(plan, host_args, batch, chunk)

    # -- drawing (host only; rendering back ends are out of scope) -------------------
    def draw(self, figure: str = "text", args: tuple = (), kwargs: Optional[dict] = None,
             **draw_kwargs: Any):
        if figure not in ("text", "mpl", "tikz", "pulse"):
            raise ValueError(
                f"Invalid figure mode: {figure!r}. Must be 'text', 'mpl', 'tikz', or 'pulse'."
            )
        if figure != "text":
            raise NotImplementedError(
                "only the text listing is provided; matplotlib/TikZ rendering "
                "(drawing.py) is outside the B200 backend scope"
            )
        tape = self._record(*args, **(kwargs or {}))
        ops = [o for o in tape if not isinstance(o, KrausChannel)]
        return "\n".join(repr(o) for o in ops)


def _lifted_diag(ob: Operation, n_qubits: int) -> np.ndarray:
    """diag of the lifted observable without building the 2^n x 2^n matrix."""
    d = np.diag(np.asarray(ob.matrix))
    k = len(ob.wires)
    idx = np.arange(2**n_qubits)
    loc = np.zeros_like(idx)
    for t, w in enumerate(ob.wires):
        loc |= ((idx >> (n_qubits - 1 - w)) & 1) << (k - 1 - t)
    return d[loc]
