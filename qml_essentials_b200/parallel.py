"""Batch sharding over GPUs: one process per GPU, ``torch.distributed`` plumbing.

The batched regime shards naturally: every (input, param)
element is an independent circuit, so ranks take contiguous slices of the sample
axis and there is NO data-path collective.  The only exchange is one small
all-reduce of the sufficient statistics behind the averages the analysis callers
report (Fourier-coefficient correlation sums, the expressibility histogram, the
Meyer-Wallach sum) - a few KB.  With NCCL the buffers live on the GPU; the CPU
test-suite drives the same code with gloo.

Nothing here is used unless ``torch.distributed`` has been initialised by the
launcher (``torchrun``); single-process calls see world size 1.
"""

from __future__ import annotations

from typing import Tuple

import numpy as np


def _dist():
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return None
    return dist if dist.is_available() and dist.is_initialized() else None


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when not launched distributed."""
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def shard_bounds(n: int, rank: int = None, size: int = None) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of ``range(n)`` owned by ``rank``: the first
    ``n % size`` ranks hold one extra element."""
    if rank is None or size is None:
        rank, size = world()
    q, r = divmod(n, size)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def allreduce_sum(x: np.ndarray) -> np.ndarray:
    """Element-wise sum over ranks of a small float64 / int64 / complex128 host array.
    Identity for world size 1.  The reduction order is NCCL's (or gloo's) - sums of
    a few doubles, not bitwise tied to the rank count."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return x
    import torch

    x = np.ascontiguousarray(x)
    is_c = np.iscomplexobj(x)
    flat = x.astype(np.complex128).view(np.float64) if is_c else x
    t = torch.from_numpy(np.array(flat, copy=True))
    if d.get_backend() == "nccl":
        t = t.cuda()
    d.all_reduce(t)
    out = t.cpu().numpy()
    if is_c:
        out = out.view(np.complex128)
    return out.reshape(x.shape).astype(x.dtype, copy=False)


def allgather_concat(x: np.ndarray, axis: int = 0) -> np.ndarray:
    """Concatenate per-rank arrays (ragged along ``axis``) in rank order."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return x
    parts = [None] * d.get_world_size()
    d.all_gather_object(parts, np.ascontiguousarray(x))
    return np.concatenate(parts, axis=axis)
