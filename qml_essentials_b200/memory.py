"""Memory estimation and batch chunking, re-derived for HBM.

Interface mirror of the reference's ``qml_essentials/memory.py`` (memory.py:54-345)
with a different model underneath.  The reference budgets host RAM for an XLA
program that materialises a fresh ``(B, 2^n)`` / ``(B, 4^n)`` array per tape
operation (hence its ``n_ops`` factor, memory.py:133-139).  The CUDA kernels
evolve the state in place - in registers / shared memory when one state fits on
chip, otherwise in one HBM buffer that is the output itself for ``state`` /
``density`` results - so the peak is::

    output bytes  +  (B * state bytes   if the state lives in HBM and is not the output)

and the budget is the free HBM of the current device (180 GB on a B200) instead
of ``psutil`` host memory.  Everything is plain integer arithmetic.

The figure is an UPPER bound that does not know which kernel strategy the library will
plan: besides the state it budgets the per-element table of evaluated gate matrices the
batched on-chip / streaming kernels keep (``n_ops`` matrices of up to 4x4 entries).
``Script`` itself chunks on the exact number the library reports
(``qmlb_workspace_bytes`` through ``CudaExecutor.peak_bytes``).
"""

from __future__ import annotations

import logging
from typing import Callable, Tuple

import numpy as np

from . import config

log = logging.getLogger(__name__)

# kept for interface compatibility (memory.py:23); there are no JIT caches to clear
CLEAR_CACHES_BETWEEN_CHUNKS: bool = False

# a statevector of at most this many qubits is evolved in registers by one thread
# (qmlb_reg.cuh) and never touches a workspace when the result is reduced in registers
REGISTER_QUBITS = 5
_SAFETY = 1.1


def _element_sizes() -> Tuple[int, int]:
    """(complex, real) element sizes of the active precision (memory.py:26-33)."""
    elem = config.complex_itemsize()
    return elem, elem // 2


def _output_bytes(type: str, batch_size: int, dim: int, elem: int, real_elem: int,
                  n_obs: int) -> int:
    """Bytes of the returned ``(batch_size, ...)`` array (memory.py:36-51)."""
    if type == "density":
        return batch_size * dim * dim * elem
    if type == "expval":
        return batch_size * max(n_obs, 1) * real_elem
    if type == "probs":
        return batch_size * dim * real_elem
    return batch_size * dim * elem


def state_bytes(n_qubits: int, evolves_density: bool, elem: int) -> int:
    return (4**n_qubits if evolves_density else 2**n_qubits) * elem


def estimate_peak_bytes(
    n_qubits: int,
    batch_size: int,
    type: str,
    use_density: bool,
    n_obs: int = 0,
    n_ops: int = 1,
) -> int:
    """Peak device bytes of one batched launch.

    ``use_density`` says whether a 4^n density matrix is evolved (noisy tape).  A
    noise-free circuit asked for ``"density"`` evolves a statevector and forms
    the outer product once (simulation.py:182-189); ``Script`` passes
    ``use_density=False`` for it.  ``n_ops`` is accepted for signature
    compatibility; in-place evolution makes the peak independent of depth.
    """
    dim = 2**n_qubits
    elem, real_elem = _element_sizes()
    out = _output_bytes(type, batch_size, dim, elem, real_elem, n_obs)
    evolves_density = bool(use_density)
    st = state_bytes(n_qubits, evolves_density, elem)
    in_place = (type == "density" and evolves_density) or (type == "state" and not evolves_density)
    if not evolves_density and n_qubits <= REGISTER_QUBITS and type != "density":
        work = 0  # registers; <Z>, probabilities and the state are written directly
    else:
        # evaluated matrices per element (<= 4x4 complex per op) + the state unless the
        # output buffer is the state
        work = batch_size * (max(n_ops, 1) * 16 * elem + (0 if in_place else st))
    return int((out + work) * _SAFETY)


def available_memory_bytes() -> int:
    """Free bytes on the current CUDA device (host RAM only when no GPU exists, so
    that the arithmetic stays testable on CPU)."""
    try:
        import torch

        if torch.cuda.is_available():
            return int(torch.cuda.mem_get_info()[0])
    except Exception:  # pragma: no cover
        pass
    try:
        import psutil

        return int(psutil.virtual_memory().available)
    except Exception:  # pragma: no cover
        return 4 * 1024**3


def compute_chunk_size(
    n_qubits: int,
    batch_size: int,
    type: str,
    use_density: bool,
    n_obs: int = 0,
    memory_fraction: float = 0.8,
    n_ops: int = 1,
) -> int:
    """Largest chunk of the batch that fits in ``memory_fraction`` of free memory
    next to the full output accumulator (memory.py:186-261).  Returns
    ``batch_size`` when everything fits; never less than 1."""
    avail = int(available_memory_bytes() * memory_fraction)
    full = estimate_peak_bytes(n_qubits, batch_size, type, use_density, n_obs, n_ops=n_ops)
    if full <= avail:
        return batch_size
    dim = 2**n_qubits
    elem, real_elem = _element_sizes()
    accum = _output_bytes(type, batch_size, dim, elem, real_elem, n_obs)
    room = max(avail - accum, elem)
    per_elem = estimate_peak_bytes(n_qubits, 1, type, use_density, n_obs, n_ops=n_ops)
    if per_elem <= 0:
        return batch_size
    chunk = max(1, min(room // per_elem, batch_size))
    if chunk == 1 and per_elem > avail:
        log.warning(
            f"A single batch element requires ~{per_elem / 1024**3:.2f} GB "
            f"but only ~{avail / 1024**3:.2f} GB is available. "
            "Proceeding with chunk_size=1 but OOM is possible."
        )
    log.info(
        f"Computation requires ~{full / 1024**3:.2f} GB which does not fit in "
        f"~{avail / 1024**3:.2f} GB. Using chunk size {chunk}."
    )
    return int(chunk)


def execute_chunked(
    batched_fn: Callable,
    args: tuple,
    in_axes: Tuple,
    batch_size: int,
    chunk_size: int,
    clear_caches: bool = False,
):
    """Run ``batched_fn`` over sub-batches and write into one preallocated result
    (memory.py:264-345).  ``Script`` itself chunks by batch offset on the device;
    this generic helper keeps the reference's callable-based interface."""
    n_chunks = -(-batch_size // chunk_size)
    log.debug(
        f"Memory-aware chunking: splitting batch of {batch_size} into "
        f"{n_chunks} chunks of <={chunk_size} elements."
    )
    output = None
    for c in range(n_chunks):
        lo, hi = c * chunk_size, min((c + 1) * chunk_size, batch_size)
        part = tuple(
            a if ax is None else np.take(a, np.arange(lo, hi), axis=ax)
            for a, ax in zip(args, in_axes)
        )
        res = np.asarray(batched_fn(*part))
        if output is None:
            output = np.zeros((batch_size,) + res.shape[1:], dtype=res.dtype)
        output[lo:hi] = res
    return output
