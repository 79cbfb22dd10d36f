"""ctypes binding of ``libqmlb200.so`` (``include/qmlb200.h``) and the executor
``Script`` dispatches to.

PyTorch is used for plumbing only: it owns the device buffers (arguments,
output, workspace), the CUDA stream the library launches on, and pinned host
staging.  No computation of the hot path happens in torch or NumPy.

There is NO CPU fallback: a missing library or a missing GPU raises
:class:`BackendUnavailable`.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Tuple

import numpy as np

from . import compiler
from .compiler import Program

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqmlb200.so")

QMLB_C64, QMLB_C128 = 0, 1
_ERRORS = {-1: "invalid program/arguments", -2: "unsupported by the kernels",
           -3: "CUDA error", -4: "workspace too small"}

EXPORTED_SYMBOLS = (
    "qmlb_version", "qmlb_launch_count", "qmlb_last_error", "qmlb_program_create",
    "qmlb_program_destroy", "qmlb_program_info", "qmlb_workspace_bytes", "qmlb_run",
    "qmlb_sample", "qmlb_purity", "qmlb_overlap_fidelity", "qmlb_fma_peak",
    "qmlb_evolve", "qmlb_zsums", "qmlb_zsums_workspace_bytes", "qmlb_plan_describe",
    "qmlb_evolve_peer", "qmlb_grid_dft", "qmlb_coef_moments", "qmlb_allreduce_buffer_bytes",
    "qmlb_allreduce_peer", "qmlb_partial_trace", "qmlb_marginal_probs",
)
QMLB_DESC_FORCE_STREAM = 1


class BackendUnavailable(RuntimeError):
    """The CUDA library or a CUDA device is missing; circuits cannot execute."""


class BackendError(RuntimeError):
    """A library call failed (message from ``qmlb_last_error``)."""


class _Arg(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("stride", C.c_int64), ("div", C.c_int64),
                ("mod", C.c_int64)]


class _Desc(C.Structure):
    _fields_ = [
        ("n_qubits", C.c_int32), ("n_bits", C.c_int32), ("density", C.c_int32),
        ("dtype", C.c_int32), ("out_type", C.c_int32), ("reserved", C.c_int32),
        ("ops", C.c_void_p), ("n_ops", C.c_int32),
        ("sources", C.c_void_p), ("n_sources", C.c_int32),
        ("items", C.c_void_p), ("n_items", C.c_int32),
        ("angles", C.c_void_p), ("n_angles", C.c_int32),
        ("terms", C.c_void_p), ("n_terms", C.c_int32),
        ("consts", C.c_void_p), ("n_consts", C.c_int64),
        ("obs", C.c_void_p), ("n_obs", C.c_int32),
        ("obs_consts", C.c_void_p), ("n_obs_consts", C.c_int64),
        ("pre", C.c_void_p), ("n_pre", C.c_int32),
    ]


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """dlopen the library and declare its prototypes.  Works without a GPU (the
    CPU test-suite checks the exported symbols this way)."""
    if not os.path.exists(path):
        raise BackendUnavailable(
            f"{path} is missing - build it with `python -m qml_essentials_b200.build` "
            "(nvcc, sm_100a).  There is no CPU execution path."
        )
    lib = C.CDLL(path)
    lib.qmlb_version.restype = C.c_int
    lib.qmlb_launch_count.restype = C.c_ulonglong
    lib.qmlb_last_error.restype = C.c_char_p
    lib.qmlb_program_create.argtypes = [C.POINTER(_Desc), C.POINTER(C.c_void_p)]
    lib.qmlb_program_destroy.argtypes = [C.c_void_p]
    lib.qmlb_program_info.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32)]
    lib.qmlb_workspace_bytes.argtypes = [C.c_void_p, C.POINTER(_Arg), C.c_int32, C.c_int64]
    lib.qmlb_workspace_bytes.restype = C.c_size_t
    lib.qmlb_run.argtypes = [C.c_void_p, C.POINTER(_Arg), C.c_int32, C.c_int64, C.c_int64,
                             C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.qmlb_sample.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int32,
                                C.c_int64, C.c_void_p, C.c_void_p]
    lib.qmlb_purity.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int32,
                                C.c_void_p, C.c_void_p]
    lib.qmlb_overlap_fidelity.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int32,
                                          C.c_void_p, C.c_void_p]
    lib.qmlb_fma_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    lib.qmlb_evolve.argtypes = [C.c_void_p, C.POINTER(_Arg), C.c_int32, C.c_int64, C.c_int64,
                                C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.qmlb_zsums_workspace_bytes.argtypes = [C.c_int64, C.c_int32]
    lib.qmlb_zsums_workspace_bytes.restype = C.c_size_t
    lib.qmlb_evolve_peer.argtypes = [C.c_void_p, C.POINTER(_Arg), C.c_int32, C.c_void_p,
                                     C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_void_p,
                                     C.c_size_t, C.c_void_p]
    lib.qmlb_grid_dft.argtypes = [C.c_void_p, C.c_int, C.c_int32, C.c_int64, C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_void_p]
    lib.qmlb_coef_moments.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int32, C.c_int64,
                                      C.c_void_p, C.c_void_p]
    lib.qmlb_allreduce_buffer_bytes.argtypes = [C.c_int64]
    lib.qmlb_allreduce_buffer_bytes.restype = C.c_size_t
    lib.qmlb_allreduce_peer.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int64,
                                        C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    lib.qmlb_partial_trace.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.c_void_p,
                                       C.c_int32, C.c_void_p, C.c_void_p]
    lib.qmlb_marginal_probs.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.c_void_p,
                                        C.c_int32, C.c_void_p, C.c_void_p]
    lib.qmlb_plan_describe.argtypes = [C.POINTER(_Desc), C.c_char_p, C.c_size_t]
    lib.qmlb_zsums.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.c_void_p,
                               C.c_void_p, C.c_size_t, C.c_void_p]
    return lib


def _np_ptr(a: np.ndarray) -> Optional[int]:
    return a.ctypes.data if a.size else None


def make_desc(prog: Program, out_type: int, obs_recs, obs_pool, precision: str,
              flags: int = 0):
    """``qmlb_program_desc`` over the arrays of a compiled program.  Returns the struct
    and the list of arrays that must stay alive while it is in use."""
    dt = QMLB_C128 if precision == "complex128" else QMLB_C64
    keep = [np.ascontiguousarray(x) for x in (
        prog.ops, prog.sources, prog.items, prog.angles, prog.terms, prog.consts,
        obs_recs, obs_pool,
        prog.pre if prog.pre is not None else np.zeros(0, dtype=compiler.PRE_DTYPE))]
    d = _Desc(
        n_qubits=prog.n_qubits, n_bits=prog.n_bits, density=int(prog.density), dtype=dt,
        out_type=int(out_type), reserved=int(flags),
        ops=_np_ptr(keep[0]), n_ops=len(keep[0]),
        sources=_np_ptr(keep[1]), n_sources=len(keep[1]),
        items=_np_ptr(keep[2]), n_items=len(keep[2]),
        angles=_np_ptr(keep[3]), n_angles=len(keep[3]),
        terms=_np_ptr(keep[4]), n_terms=len(keep[4]),
        consts=_np_ptr(keep[5]), n_consts=len(keep[5]),
        obs=_np_ptr(keep[6]), n_obs=len(keep[6]),
        obs_consts=_np_ptr(keep[7]), n_obs_consts=len(keep[7]),
        pre=_np_ptr(keep[8]), n_pre=len(keep[8]),
    )
    return d, keep


def plan_describe(lib, prog: Program, out_type: int, obs_recs, obs_pool, precision: str,
                  flags: int = 0) -> str:
    """Host-only planning (``qmlb_plan_describe``): no GPU needed."""
    d, keep = make_desc(prog, out_type, obs_recs, obs_pool, precision, flags)
    buf = C.create_string_buffer(1 << 22)
    rc = lib.qmlb_plan_describe(C.byref(d), buf, len(buf))
    if rc != 0:
        raise BackendError(f"qmlb_plan_describe: {_ERRORS.get(rc, rc)}: "
                           f"{lib.qmlb_last_error().decode()}")
    return buf.value.decode()


class ProgramHandle:
    """Owns one ``qmlb_program*`` (created on the current CUDA device)."""

    def __init__(self, lib, prog: Program, out_type: int, obs_recs, obs_pool, precision: str,
                 flags: int = 0):
        self.lib = lib
        self.ptr = C.c_void_p()
        dt = QMLB_C128 if precision == "complex128" else QMLB_C64
        d, keep = make_desc(prog, out_type, obs_recs, obs_pool, precision, flags)
        rc = lib.qmlb_program_create(C.byref(d), C.byref(self.ptr))
        if rc != 0:
            raise BackendError(
                f"qmlb_program_create: {_ERRORS.get(rc, rc)}: "
                f"{lib.qmlb_last_error().decode()}")
        s, p, o = C.c_int32(), C.c_int32(), C.c_int32()
        lib.qmlb_program_info(self.ptr, C.byref(s), C.byref(p), C.byref(o))
        self.strategy, self.n_passes, self.n_device_ops = s.value, p.value, o.value
        self.dtype = dt
        self.precision = precision
        self.n_qubits, self.n_bits, self.density = prog.n_qubits, prog.n_bits, prog.density
        self.out_type, self.n_obs = int(out_type), len(obs_recs)

    def __del__(self):
        try:
            if self.ptr:
                self.lib.qmlb_program_destroy(self.ptr)
        except Exception:
            pass


class DeviceCall:
    """One staged batched execution: device-resident arguments, output and
    workspace.  ``launch()`` enqueues the kernels on the current torch stream;
    ``result()`` copies the output back."""

    def __init__(self, ex: "CudaExecutor", handle: ProgramHandle, dev_args, batch: int,
                 batch_offset: int = 0):
        import torch

        self.ex, self.h, self.batch, self.offset = ex, handle, int(batch), int(batch_offset)
        self.dev_args = dev_args  # list of (tensor|None, div, mod)
        cplx = torch.complex128 if handle.precision == "complex128" else torch.complex64
        real = torch.float64 if handle.precision == "complex128" else torch.float32
        dim = 2**handle.n_qubits
        shape, dt = {
            compiler.OUT_STATE: ((batch, dim), cplx),
            compiler.OUT_PROBS: ((batch, dim), real),
            compiler.OUT_EXPVAL: ((batch, handle.n_obs), real),
            compiler.OUT_DENSITY: ((batch, dim, dim), cplx),
        }[handle.out_type]
        self.out = torch.empty(shape, dtype=dt, device=ex.device)
        self.c_args = (_Arg * max(len(dev_args), 1))()
        for i, a in enumerate(dev_args):
            if a is None:
                self.c_args[i] = _Arg(None, 0, 1, 1)
            else:
                t, div, mod = a
                self.c_args[i] = _Arg(t.data_ptr(), t.shape[1], int(div), int(mod))
        ws = ex.lib.qmlb_workspace_bytes(handle.ptr, self.c_args, len(dev_args), self.batch)
        self.ws_bytes = int(ws)
        self.ws = torch.empty(max(self.ws_bytes, 1), dtype=torch.uint8, device=ex.device)

    def launch(self):
        import torch

        stream = torch.cuda.current_stream(self.ex.device).cuda_stream
        rc = self.ex.lib.qmlb_run(
            self.h.ptr, self.c_args, len(self.dev_args), self.batch, self.offset,
            self.out.data_ptr(), self.ws.data_ptr(), self.ws_bytes, stream)
        if rc != 0:
            raise BackendError(
                f"qmlb_run: {_ERRORS.get(rc, rc)}: {self.ex.lib.qmlb_last_error().decode()}")
        return self.out

    def result(self) -> np.ndarray:
        """Device -> host through a pinned staging buffer (torch's caching host
        allocator recycles it); the returned array owns that buffer."""
        import torch

        host = torch.empty(self.out.shape, dtype=self.out.dtype, pin_memory=True)
        host.copy_(self.out, non_blocking=True)
        torch.cuda.current_stream(self.ex.device).synchronize()
        return host.numpy()


class CudaExecutor:
    """Executes compiled plans on the current CUDA device through the C ABI."""

    name = "cuda-sm100a"

    def __init__(self, device: Optional[int] = None):
        try:
            import torch
        except Exception as e:  # pragma: no cover
            raise BackendUnavailable(f"PyTorch is required for device memory: {e}")
        if not torch.cuda.is_available():
            raise BackendUnavailable(
                "no CUDA device is visible; the qml-essentials B200 backend has no CPU "
                "execution path")
        self.lib = load_library()
        self.torch = torch
        self._cache = {}
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None
                                   else device)

    # -- staging ---------------------------------------------------------------
    def handle_for(self, plan) -> ProgramHandle:
        key = ("handle", self.device.index)
        h = plan.device.get(key)
        if h is None:
            with self.torch.cuda.device(self.device):
                h = ProgramHandle(self.lib, plan.program, plan.out_type, plan.obs_recs,
                                  plan.obs_pool, plan.precision)
            plan.device[key] = h
        return h

    def to_device(self, host_args):
        torch = self.torch
        out = []
        for a in host_args:
            if a is None:
                out.append(None)
                continue
            arr, div, mod = a
            t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64))
            out.append((t.to(self.device, non_blocking=True), div, mod))
        return out

    def stage(self, plan, host_args, batch: int, batch_offset: int = 0) -> DeviceCall:
        with self.torch.cuda.device(self.device):
            return DeviceCall(self, self.handle_for(plan), self.to_device(host_args), batch,
                              batch_offset)

    def peak_bytes(self, plan, host_args, batch: int) -> int:
        """Exact device bytes of one launch over ``batch`` elements: the library's own
        workspace figure (hoisted-factor tables, evaluated matrices, state, reduction
        partials - ``qmlb_workspace_bytes``) plus the output and the uploaded arguments.
        ``Script`` chunks on this instead of the arithmetic model of ``memory.py`` so that the
        chunk size follows what the chosen kernel strategy really allocates."""
        h = self.handle_for(plan)
        c_args = (_Arg * max(len(host_args), 1))()
        arg_bytes = 0
        for i, a in enumerate(host_args):
            if a is None:
                c_args[i] = _Arg(None, 0, 1, 1)
            else:
                arr, div, mod = a
                c_args[i] = _Arg(None, int(np.asarray(arr).shape[1]), int(div), int(mod))
                arg_bytes += int(np.asarray(arr).size) * 8
        ws = int(self.lib.qmlb_workspace_bytes(h.ptr, c_args, len(host_args), int(batch)))
        elem = 16 if plan.precision == "complex128" else 8
        dim = 2 ** plan.n_qubits
        out = {compiler.OUT_STATE: dim * elem, compiler.OUT_PROBS: dim * elem // 2,
               compiler.OUT_EXPVAL: max(h.n_obs, 1) * elem // 2,
               compiler.OUT_DENSITY: dim * dim * elem}[h.out_type] * int(batch)
        return ws + out + arg_bytes

    # -- Script-facing entry points -----------------------------------------------
    def execute(self, plan, host_args, batch: int, chunk: Optional[int] = None,
                to_host: bool = True):
        """Run the plan over the batch (in HBM-sized chunks if needed).  Returns a
        NumPy array, or the device tensor when ``to_host`` is False."""
        chunk = batch if not chunk else min(chunk, batch)
        with self.torch.cuda.device(self.device):
            if chunk >= batch:
                call = self.stage(plan, host_args, batch)
                out = call.launch()
                return call.result() if to_host else out
            dev_args = self.to_device(host_args)
            handle = self.handle_for(plan)
            parts = []
            for lo in range(0, batch, chunk):
                call = DeviceCall(self, handle, dev_args, min(chunk, batch - lo), lo)
                out = call.launch()
                parts.append(call.result() if to_host else out)
            if to_host:
                return np.concatenate(parts, axis=0)
            return self.torch.cat(parts, dim=0)

    SHOT_MAX_QUBITS = 14  # qmlb_sample keeps one probability vector in shared memory

    def execute_shots(self, plan, host_args, batch: int, uniforms: np.ndarray,
                      chunk: Optional[int] = None) -> np.ndarray:
        """Exact probabilities on the device, then the shot bookkeeping kernel - chunk by
        chunk: probabilities, the (chunk, shots) uniforms and the int32 counts of one chunk
        are resident at a time (memory-aware like ``execute``)."""
        if plan.n_qubits > self.SHOT_MAX_QUBITS:
            raise BackendError(
                f"shot sampling is limited to {self.SHOT_MAX_QUBITS} qubits (qmlb_sample holds "
                f"the cumulative probabilities of one element in shared memory); got "
                f"{plan.n_qubits}")
        torch = self.torch
        shots = uniforms.shape[1]
        # the uniforms and counts of a chunk take part in the device budget
        per_elem = shots * 8 + (2 ** plan.n_qubits) * 12
        free = int(torch.cuda.mem_get_info(self.device)[0] * 0.5)
        chunk = batch if not chunk else min(chunk, batch)
        chunk = int(max(1, min(chunk, free // max(per_elem, 1))))
        with torch.cuda.device(self.device):
            dev_args = self.to_device(host_args)
            handle = self.handle_for(plan)
            parts = []
            for lo in range(0, batch, chunk):
                n = min(chunk, batch - lo)
                call = DeviceCall(self, handle, dev_args, n, lo)
                probs = call.launch()
                u = torch.from_numpy(np.ascontiguousarray(uniforms[lo:lo + n],
                                                          dtype=np.float64)).to(self.device)
                counts = torch.empty((n, probs.shape[1]), dtype=torch.int32, device=self.device)
                rc = self.lib.qmlb_sample(
                    probs.data_ptr(), handle.dtype, u.data_ptr(), n, plan.n_qubits, shots,
                    counts.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
                if rc != 0:
                    raise BackendError(f"qmlb_sample: {self.lib.qmlb_last_error().decode()}")
                parts.append(counts.cpu().numpy())
            return parts[0] if len(parts) == 1 else np.concatenate(parts, axis=0)

    # -- fused reductions for the analysis callers -----------------------------------
    def purities(self, states, n_qubits: int, is_density: bool):
        """(B, n) single-qubit reduced purities of device-resident states."""
        torch = self.torch
        dt = QMLB_C128 if states.dtype == torch.complex128 else QMLB_C64
        B = states.shape[0]
        out = torch.empty((B, n_qubits), dtype=torch.float64 if dt else torch.float32,
                          device=states.device)
        rc = self.lib.qmlb_purity(states.data_ptr(), dt, int(is_density), B, n_qubits,
                                  out.data_ptr(),
                                  torch.cuda.current_stream(states.device).cuda_stream)
        if rc != 0:
            raise BackendError(f"qmlb_purity: {self.lib.qmlb_last_error().decode()}")
        return out

    def overlap_fidelities(self, states, n_qubits: int):
        """|<psi_b|psi_{b+B/2}>|^2 for b < B/2 of device-resident pure states."""
        torch = self.torch
        dt = QMLB_C128 if states.dtype == torch.complex128 else QMLB_C64
        half = states.shape[0] // 2
        out = torch.empty((half,), dtype=torch.float64 if dt else torch.float32,
                          device=states.device)
        rc = self.lib.qmlb_overlap_fidelity(
            states.data_ptr(), dt, half, n_qubits, out.data_ptr(),
            torch.cuda.current_stream(states.device).cuda_stream)
        if rc != 0:
            raise BackendError(f"qmlb_overlap_fidelity: {self.lib.qmlb_last_error().decode()}")
        return out

    def grid_dft(self, ev, order=None):
        """(n_x, n_p, n_obs) real device expvals -> complex coefficients (mean over
        observables, DFT along the grid axis, 1/n_x normalisation).  ``order``: the
        frequencies (numpy.fft indices) wanted, in output-row order - shift / trim folded
        into the store; default all of them in numpy.fft order."""
        torch = self.torch
        n_x, n_p, n_obs = ev.shape
        dt = QMLB_C128 if ev.dtype == torch.float64 else QMLB_C64
        n_rows, row_of = n_x, None
        if order is not None:
            order = np.asarray(order, dtype=np.int64)
            n_rows = len(order)
            table = np.full(n_x, -1, dtype=np.int32)
            table[order] = np.arange(n_rows, dtype=np.int32)
            key = ("dft_rows", n_x, order.tobytes())
            row_of = self._cache.get(key)
            if row_of is None:
                row_of = self._cache[key] = torch.from_numpy(table).to(ev.device)
        out = torch.empty((n_rows, n_p), dtype=torch.complex128 if dt else torch.complex64,
                          device=ev.device)
        rc = self.lib.qmlb_grid_dft(ev.data_ptr(), dt, n_x, n_p, n_obs,
                                    row_of.data_ptr() if row_of is not None else None,
                                    out.data_ptr(),
                                    torch.cuda.current_stream(ev.device).cuda_stream)
        if rc != 0:
            raise BackendError(f"qmlb_grid_dft: {self.lib.qmlb_last_error().decode()}")
        return out

    def partial_trace(self, rho, n_qubits: int, keep):
        """(B, 2^n, 2^n) device density matrices -> (B, 2^k, 2^k) over the wires ``keep``."""
        torch = self.torch
        keep = sorted(int(q) for q in keep)
        k, B = len(keep), rho.shape[0]
        dt = QMLB_C128 if rho.dtype == torch.complex128 else QMLB_C64
        out = torch.empty((B, 2**k, 2**k), dtype=rho.dtype, device=rho.device)
        arr = (C.c_int32 * k)(*keep)
        rc = self.lib.qmlb_partial_trace(rho.data_ptr(), dt, B, n_qubits, arr, k, out.data_ptr(),
                                         torch.cuda.current_stream(rho.device).cuda_stream)
        if rc != 0:
            raise BackendError(f"qmlb_partial_trace: {self.lib.qmlb_last_error().decode()}")
        return out

    def marginal_probs(self, probs, n_qubits: int, keep):
        """(B, 2^n) device probabilities -> (B, 2^k) marginal over the wires ``keep``."""
        torch = self.torch
        keep = sorted(int(q) for q in keep)
        k, B = len(keep), probs.shape[0]
        dt = QMLB_C128 if probs.dtype == torch.float64 else QMLB_C64
        out = torch.empty((B, 2**k), dtype=probs.dtype, device=probs.device)
        arr = (C.c_int32 * k)(*keep)
        rc = self.lib.qmlb_marginal_probs(probs.data_ptr(), dt, B, n_qubits, arr, k,
                                          out.data_ptr(),
                                          torch.cuda.current_stream(probs.device).cuda_stream)
        if rc != 0:
            raise BackendError(f"qmlb_marginal_probs: {self.lib.qmlb_last_error().decode()}")
        return out

    def to_host(self, t) -> np.ndarray:
        """Device tensor -> NumPy through pinned staging."""
        host = self.torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        host.copy_(t, non_blocking=True)
        self.torch.cuda.current_stream(t.device).synchronize()
        return host.numpy()

    def coef_moments(self, coef, rows):
        """Sufficient statistics over the samples of the selected coefficient rows:
        ``(sum c_i, sum |c_i|^2, sum conj(c_i) c_j)`` as complex128 device tensors."""
        torch = self.torch
        K, n_p = len(rows), coef.shape[1]
        dt = QMLB_C128 if coef.dtype == torch.complex128 else QMLB_C64
        idx = torch.as_tensor(np.asarray(rows, dtype=np.int32), device=coef.device)
        out = torch.empty(2 * K + K * K, dtype=torch.complex128, device=coef.device)
        rc = self.lib.qmlb_coef_moments(coef.data_ptr(), dt, idx.data_ptr(), K, n_p,
                                        out.data_ptr(),
                                        torch.cuda.current_stream(coef.device).cuda_stream)
        if rc != 0:
            raise BackendError(f"qmlb_coef_moments: {self.lib.qmlb_last_error().decode()}")
        return out[:K], out[K:2 * K].real, out[2 * K:].reshape(K, K)

    def fma_peak_tflops(self, precision: str) -> float:
        v = C.c_double()
        rc = self.lib.qmlb_fma_peak(QMLB_C128 if precision == "complex128" else QMLB_C64,
                                    C.byref(v))
        if rc != 0:
            raise BackendError(f"qmlb_fma_peak: {self.lib.qmlb_last_error().decode()}")
        return v.value

    def launch_count(self) -> int:
        return int(self.lib.qmlb_launch_count())
