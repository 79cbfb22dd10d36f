"""Tape -> device program compiler.

Replaces the reference's trace-and-``vmap`` step (script.py:272-329) and the
per-gate ``einsum`` loops (simulation.py:86-104,124-128).  A recorded tape (with
affine-proxy angles, see :mod:`.symbolic`) is lowered to a flat program over the
bits of ONE state index:

* statevector: ``N = n`` bits, wire ``q`` is bit ``n-1-q`` (wire 0 = MSB,
  simulation.py:100-101);
* density matrix: ``N = 2n`` bits - rho is evolved as a vector, ket wire ``q`` is
  bit ``2n-1-q`` and bra wire ``q`` is bit ``n-1-q`` (operations.py:505-510).  A
  unitary is ``U`` on the ket bits and ``conj(U)`` on the bra bits; a channel is
  its superoperator ``sum_k K (x) conj(K)`` on (ket bits, bra bits).

Fusion done here:

* runs of single-qubit gates on one wire collapse into one *chain* source
  (RY.RZ.RY.RX -> one 2x2 per element);
* in density mode every run of single-qubit unitaries AND single-qubit channels
  on a wire collapses into one 4x4 *superchain* acting on (ket bit, bra bit) -
  the 272 depolarizing + 8 amplitude-damping channels of BASELINE config 4 cost
  no state pass of their own;
* constant matrices are classified into identity (dropped), permutation (CX,
  SWAP, CCX, ... -> index shuffle), diagonal, controlled-2x2 and dense;
* concrete angles are folded to constants.

The emitted arrays are exactly what ``include/qmlb200.h`` describes.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .operations import (
    Barrier,
    ConstMat,
    DiagPhaseMat,
    KrausChannel,
    KrausSet,
    Operation,
    ProductMat,
    TrigMat,
)
from .symbolic import Sym, is_symbolic

# ---- enums shared with csrc/qmlb_program.h ---------------------------------
OP_MAT, OP_CTRL1, OP_PERM, OP_DIAG = 0, 1, 2, 3
SRC_CONST, SRC_TRIG, SRC_CHAIN, SRC_DIAGPH, SRC_TABLE, SRC_SUPER, SRC_PRE = 0, 1, 2, 3, 4, 5, 6
FLAG_CONJ, FLAG_DIAGVEC, FLAG_ROT_SHIFT = 1, 2, 2  # flags bits 2-3: 1=RX 2=RY 3=RZ
OUT_STATE, OUT_PROBS, OUT_EXPVAL, OUT_DENSITY = 0, 1, 2, 3
OBS_ZSTRING, OBS_DIAG, OBS_DENSE = 0, 1, 2
MAX_OP_BITS = 8

OP_DTYPE = np.dtype(
    [("kind", "<i4"), ("k", "<i4"), ("src", "<i4"), ("aux", "<i4"),
     ("bits", "<i4", (MAX_OP_BITS,))]
)
SRC_DTYPE = np.dtype(
    [("kind", "<i4"), ("k", "<i4"), ("a0", "<i4"), ("a1", "<i4"), ("a2", "<i4"),
     ("angle", "<i4"), ("flags", "<i4"), ("pad", "<i4"), ("kappa", "<f8")]
)
ANGLE_DTYPE = np.dtype([("first", "<i4"), ("n", "<i4"), ("c0", "<f8")])
TERM_DTYPE = np.dtype([("arg", "<i4"), ("offset", "<i4"), ("coeff", "<f8")])
PRE_DTYPE = np.dtype([("src", "<i4"), ("arg", "<i4"), ("local", "<i4"), ("pad", "<i4")])
OBS_DTYPE = np.dtype(
    [("kind", "<i4"), ("k", "<i4"), ("a0", "<i4"), ("pad", "<i4"), ("zmask", "<i8"),
     ("bits", "<i4", (MAX_OP_BITS,))]
)

_TOL = 0.0  # structural classification uses exact zeros (gate constants are exact)


class CompileError(ValueError):
    pass


@dataclass
class Program:
    """Flat device program (host copy)."""

    n_qubits: int
    n_bits: int
    density: bool
    ops: np.ndarray
    sources: np.ndarray
    items: np.ndarray  # int32 source ids referenced by chain / superchain sources
    angles: np.ndarray
    terms: np.ndarray
    consts: np.ndarray  # float64 pool (complex entries interleaved re, im)
    n_args: int  # number of batched argument slots referenced by terms / tables
    pre: np.ndarray = None  # hoisted batch-invariant 2x2 factors (see _Builder.hoist)
    n_tape_ops: int = 0
    meta: dict = field(default_factory=dict)

    @property
    def max_k(self) -> int:
        return int(self.ops["k"].max()) if len(self.ops) else 0


class _Builder:
    def __init__(self):
        self.consts: List[float] = []
        self.sources: List[tuple] = []
        self.items: List[int] = []
        self.angles: List[tuple] = []
        self.terms: List[tuple] = []
        self.ops: List[tuple] = []
        self._angle_cache: Dict[tuple, int] = {}
        self._const_cache: Dict[bytes, int] = {}
        self.max_arg = -1

    # -- pools -------------------------------------------------------------
    def add_complex(self, arr) -> int:
        a = np.ascontiguousarray(np.asarray(arr, dtype=np.complex128)).ravel()
        key = b"c" + a.tobytes()
        hit = self._const_cache.get(key)
        if hit is not None:
            return hit
        if len(self.consts) % 2:
            self.consts.append(0.0)
        off = len(self.consts) // 2
        self.consts.extend(a.view(np.float64).tolist())
        self._const_cache[key] = off
        return off

    def add_real(self, arr) -> int:
        a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64)).ravel()
        key = b"r" + a.tobytes()
        hit = self._const_cache.get(key)
        if hit is not None:
            return hit
        off = len(self.consts)
        self.consts.extend(a.tolist())
        self._const_cache[key] = off
        return off

    def add_angle(self, theta: Sym) -> int:
        terms = tuple(sorted((a, o, c) for (a, o), c in theta.terms.items()))
        key = (theta.const, terms)
        hit = self._angle_cache.get(key)
        if hit is not None:
            return hit
        first = len(self.terms)
        for a, o, c in terms:
            self.terms.append((a, o, c))
            self.max_arg = max(self.max_arg, a)
        self.angles.append((first, len(terms), theta.const))
        self._angle_cache[key] = len(self.angles) - 1
        return len(self.angles) - 1

    def add_source(self, kind, k, a0=0, a1=0, a2=0, angle=-1, flags=0, kappa=0.0) -> int:
        self.sources.append((kind, k, a0, a1, a2, angle, flags, 0, kappa))
        return len(self.sources) - 1

    def add_op(self, kind, bits: Sequence[int], src=-1, aux=0) -> None:
        if len(bits) > MAX_OP_BITS:
            raise CompileError(
                f"operation on {len(bits)} state bits exceeds the supported {MAX_OP_BITS}"
            )
        b = list(bits) + [0] * (MAX_OP_BITS - len(bits))
        self.ops.append((kind, len(bits), src, aux, b))

    # -- matrix sources -------------------------------------------------------
    def source_of(self, spec) -> int:
        """Source id for a ConstMat / TrigMat / ProductMat(k=1) / _TableMat spec."""
        if isinstance(spec, ConstMat):
            k = int(np.log2(spec.matrix.shape[0]))
            return self.add_source(SRC_CONST, k, a0=self.add_complex(spec.matrix))
        if isinstance(spec, TrigMat):
            k = int(np.log2(spec.C0.shape[0]))
            return self.add_source(
                SRC_TRIG, k, a0=self.add_complex(spec.C0), a1=self.add_complex(spec.A),
                a2=self.add_complex(spec.B), angle=self.add_angle(spec.theta),
                kappa=spec.kappa, flags=_rotation_axis(spec) << FLAG_ROT_SHIFT,
            )
        if isinstance(spec, _TableMat):
            self.max_arg = max(self.max_arg, spec.arg)
            return self.add_source(SRC_TABLE, spec.k, a0=spec.arg, a1=spec.offset,
                                   flags=1 if spec.conj else 0)
        raise CompileError(f"cannot build a source from {type(spec).__name__}")

    def chain_source(self, factors: List) -> int:
        """2x2 product source; ``factors[0]`` acts first."""
        factors = _fold_constants(factors)
        if len(factors) == 1:
            return self.source_of(factors[0])
        ids = [self.source_of(f) for f in factors]
        first = len(self.items)
        self.items.extend(ids)
        return self.add_source(SRC_CHAIN, 1, a0=first, a1=len(ids))

    def super_source(self, seq: List) -> int:
        """4x4 superoperator source for a run of 1-qubit unitaries (2x2 specs) and
        1-qubit channels (4x4 ConstMat) on one wire; ``seq[0]`` acts first."""
        seq = _fold_super(seq)
        ids = []
        for kind, payload in seq:
            if kind == "u":
                ids.append(self.chain_source(payload))
            else:
                ids.append(self.add_source(SRC_CONST, 2, a0=self.add_complex(payload)))
        first = len(self.items)
        self.items.extend(ids)
        return self.add_source(SRC_SUPER, 2, a0=first, a1=len(ids))


def _hoist(b: "_Builder"):
    """Hoist batch-invariant factors out of the per-element work.

    Every 2x2 factor of a fused chain depends (through its angle) on the rows of
    some argument slots.  A maximal run of factors that depends on exactly ONE
    slot (constants ride along) is the same for every batch element that reads
    the same row of that slot - e.g. in BASELINE config 2 the RY.RZ.RY product of
    a layer depends only on the parameter sample (1024 distinct values) and the
    encoding RX only on the grid point (264 values), while the batch has 270 336
    elements.  Such runs become `pre` entries: a precompute kernel evaluates them
    once per distinct row into a table and the main program reads them through
    SRC_PRE sources (or evaluates them inline when the slot is per-element).
    """
    src, items = b.sources, b.items

    def angle_slots(aid):
        first, n, _ = b.angles[aid]
        return frozenset(a for (a, _o, _c) in b.terms[first:first + n])

    def elem_slots(sid):
        kind = src[sid][0]
        if kind == SRC_CONST:
            return frozenset()
        if kind == SRC_TRIG:
            return angle_slots(src[sid][5])
        return None  # TABLE etc.: never hoisted

    pre = []
    counts = {}

    def make_pre(seg_src, slot):
        local = counts.get(slot, 0)
        counts[slot] = local + 1
        pre.append((seg_src, slot, local, 0))
        return (SRC_PRE, 1, local, slot, len(pre) - 1, -1, 0, 0, 0.0)

    chain_items = set()
    n_src0 = len(src)
    for sid in range(n_src0):
        if src[sid][0] != SRC_CHAIN:
            continue
        ids = items[src[sid][2]: src[sid][2] + src[sid][3]]
        chain_items.update(ids)
        parts, cur, cur_slot, pend = [], [], None, []
        for i in ids:
            d = elem_slots(i)
            if d is not None and len(d) == 0:
                (cur if cur else pend).append(i)
                continue
            if d is None or len(d) > 1:
                if cur:
                    parts.append(("seg", cur_slot, cur))
                for c in pend:
                    parts.append(("inline", None, [c]))
                parts.append(("inline", None, [i]))
                cur, cur_slot, pend = [], None, []
                continue
            (slot,) = d
            if cur and slot == cur_slot:
                cur.append(i)
            else:
                if cur:
                    parts.append(("seg", cur_slot, cur))
                cur, cur_slot, pend = pend + [i], slot, []
        if cur:
            parts.append(("seg", cur_slot, cur))
        for c in pend:
            parts.append(("inline", None, [c]))
        if not any(p[0] == "seg" for p in parts):
            continue
        new_ids = []
        for kind, slot, seg in parts:
            if kind == "inline":
                new_ids.extend(seg)
                continue
            if len(seg) == 1:
                seg_src = seg[0]
            else:
                first = len(items)
                items.extend(seg)
                src.append((SRC_CHAIN, 1, first, len(seg), 0, -1, 0, 0, 0.0))
                seg_src = len(src) - 1
            src.append(make_pre(seg_src, slot))
            new_ids.append(len(src) - 1)
        if len(new_ids) == 1:
            src[sid] = src[new_ids[0]]
        else:
            first = len(items)
            items.extend(new_ids)
            src[sid] = (SRC_CHAIN, 1, first, len(new_ids), 0, -1, 0, 0, 0.0)
    # elementary 2x2 TRIG sources referenced directly (ops, CTRL1, superchain items)
    for sid in range(n_src0):
        rec = src[sid]
        if rec[0] == SRC_TRIG and rec[1] == 1 and sid not in chain_items:
            d = elem_slots(sid)
            if d is not None and len(d) == 1:
                src.append(rec)
                src[sid] = make_pre(len(src) - 1, next(iter(d)))
    return pre


_PAULIS = (np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.array([[1, 0], [0, -1]]))


def _rotation_axis(spec) -> int:
    """1/2/3 if spec is exactly cos(k t) I - i sin(k t) {X,Y,Z} (kernel fast path)."""
    if spec.C0.shape != (2, 2) or np.any(spec.C0 != 0) or not np.array_equal(spec.A, np.eye(2)):
        return 0
    for axis, P in enumerate(_PAULIS):
        if np.array_equal(spec.B, -1j * P):
            return axis + 1
    return 0


class _TableMat:
    """Per-element matrix read from a batched table argument (fallback path)."""

    symbolic = True

    def __init__(self, k, arg, offset, conj=False):
        self.k, self.arg, self.offset, self.conj = k, arg, offset, conj


# ---------------------------------------------------------------------------
# spec helpers
# ---------------------------------------------------------------------------
def _as_sym(theta) -> Sym:
    return theta if isinstance(theta, Sym) else Sym({}, float(theta))


def _simplify(spec):
    """Fold concrete angles into constants; flatten products."""
    if isinstance(spec, TrigMat):
        if not is_symbolic(spec.theta):
            return ConstMat(spec.evaluate())
        return TrigMat(spec.C0, spec.A, spec.B, spec.kappa, _as_sym(spec.theta))
    if isinstance(spec, DiagPhaseMat):
        if not is_symbolic(spec.theta):
            return ConstMat(spec.evaluate())
        return DiagPhaseMat(spec.marks, _as_sym(spec.theta))
    if isinstance(spec, ProductMat):
        flat = []
        for f in spec.factors:
            f = _simplify(f)
            flat.extend(f.factors if isinstance(f, ProductMat) else [f])
        flat = _fold_constants(flat)
        return flat[0] if len(flat) == 1 else ProductMat(flat)
    return spec


def _fold_constants(factors: List) -> List:
    out: List = []
    for f in factors:
        if isinstance(f, ConstMat) and out and isinstance(out[-1], ConstMat):
            out[-1] = ConstMat(f.matrix @ out[-1].matrix)
        else:
            out.append(f)
    return out


def _conj_spec(spec):
    if isinstance(spec, ConstMat):
        return ConstMat(spec.matrix.conj())
    if isinstance(spec, TrigMat):
        return TrigMat(spec.C0.conj(), spec.A.conj(), spec.B.conj(), spec.kappa, spec.theta)
    if isinstance(spec, ProductMat):
        return ProductMat([_conj_spec(f) for f in spec.factors])
    if isinstance(spec, DiagPhaseMat):
        return DiagPhaseMat(-spec.marks, spec.theta)
    if isinstance(spec, _TableMat):
        return _TableMat(spec.k, spec.arg, spec.offset, not spec.conj)
    raise CompileError(f"cannot conjugate {type(spec).__name__}")


def _fold_super(seq: List) -> List:
    """Merge adjacent constant items of a superchain: ('u', [2x2 specs]) and
    ('c', 4x4 ndarray)."""
    out: List = []
    for kind, payload in seq:
        if kind == "u":
            payload = _fold_constants(list(payload))
            if all(isinstance(f, ConstMat) for f in payload):
                U = payload[0].matrix
                kind, payload = "c", np.kron(U, U.conj())
        if kind == "c" and out and out[-1][0] == "c":
            out[-1] = ("c", payload @ out[-1][1])
        elif kind == "u" and out and out[-1][0] == "u":
            out[-1] = ("u", _fold_constants(out[-1][1] + payload))
        else:
            out.append((kind, payload))
    return out


def kraus_superop(kraus: List[np.ndarray]) -> np.ndarray:
    """sum_k K (x) conj(K), indexed ((ket out, bra out), (ket in, bra in))."""
    return sum(np.kron(K, K.conj()) for K in kraus)


def _is_identity(m) -> bool:
    return m.shape[0] == m.shape[1] and np.array_equal(m, np.eye(m.shape[0]))


def _is_diagonal(m) -> bool:
    return np.count_nonzero(m - np.diag(np.diag(m))) == 0


def _permutation_of(m) -> Optional[np.ndarray]:
    """perm with new[j] = old[perm[j]] if m is a 0/1 permutation matrix."""
    if np.any((m != 0) & (m != 1)):
        return None
    if not (np.all(m.sum(axis=0) == 1) and np.all(m.sum(axis=1) == 1)):
        return None
    return np.argmax(np.abs(m), axis=1).astype(np.int64)


def _support(spec) -> np.ndarray:
    if isinstance(spec, ConstMat):
        return spec.matrix != 0
    if isinstance(spec, TrigMat):
        return (spec.C0 != 0) | (spec.A != 0) | (spec.B != 0)
    raise CompileError("support of non-elementary spec")


def _controlled_split(spec):
    """If a 4x4 elementary spec is identity unless wires[0] (or wires[1]) is 1,
    return (control_position, 2x2 spec); else None."""
    if isinstance(spec, ConstMat):
        mats = {"M": spec.matrix}
    elif isinstance(spec, TrigMat):
        mats = {"C0": spec.C0, "A": spec.A, "B": spec.B}
    else:
        return None
    if next(iter(mats.values())).shape != (4, 4):
        return None
    for ctrl_pos, idx1, idx0 in ((0, [2, 3], [0, 1]), (1, [1, 3], [0, 2])):
        ok = True
        for name, m in mats.items():
            # rows/cols with control = 0 must be identity (from the constant part only)
            blk00 = m[np.ix_(idx0, idx0)]
            cross = np.count_nonzero(m[np.ix_(idx0, idx1)]) + np.count_nonzero(
                m[np.ix_(idx1, idx0)])
            want = np.eye(2) if name in ("M", "C0") else np.zeros((2, 2))
            if cross or not np.array_equal(blk00, want):
                ok = False
                break
        if ok:
            sub = {n: m[np.ix_(idx1, idx1)] for n, m in mats.items()}
            if isinstance(spec, ConstMat):
                return ctrl_pos, ConstMat(sub["M"])
            return ctrl_pos, TrigMat(sub["C0"], sub["A"], sub["B"], spec.kappa, spec.theta)
    return None


# ---------------------------------------------------------------------------
# lowering
# ---------------------------------------------------------------------------
class _Lowerer:
    def __init__(self, n_qubits: int, density: bool):
        self.n = n_qubits
        self.density = density
        self.N = 2 * n_qubits if density else n_qubits
        self.b = _Builder()
        # pending single-qubit work per wire
        self.pending: Dict[int, List] = {}

    # bit positions -----------------------------------------------------------
    def ket_bit(self, w: int) -> int:
        return self.N - 1 - w

    def bra_bit(self, w: int) -> int:
        return self.n - 1 - w

    # pending 1q ----------------------------------------------------------------
    def push_unitary_1q(self, w: int, spec) -> None:
        factors = spec.factors if isinstance(spec, ProductMat) else [spec]
        seq = self.pending.setdefault(w, [])
        if self.density:
            if seq and seq[-1][0] == "u":
                seq[-1] = ("u", seq[-1][1] + list(factors))
            else:
                seq.append(("u", list(factors)))
        else:
            seq.extend(factors)

    def push_channel_1q(self, w: int, superop: np.ndarray) -> None:
        self.pending.setdefault(w, []).append(("c", superop))

    def flush(self, wires) -> None:
        for w in wires:
            seq = self.pending.pop(w, None)
            if not seq:
                continue
            if self.density:
                seq = _fold_super(seq)
                if len(seq) == 1 and seq[0][0] == "c":
                    if _is_identity(seq[0][1]):
                        continue
                    self.emit_const([self.ket_bit(w), self.bra_bit(w)], seq[0][1])
                    continue
                self.b.add_op(OP_MAT, [self.ket_bit(w), self.bra_bit(w)],
                              self.b.super_source(seq))
            else:
                seq = _fold_constants(seq)
                if len(seq) == 1 and isinstance(seq[0], ConstMat):
                    self.emit_const([self.ket_bit(w)], seq[0].matrix)
                    continue
                self.b.add_op(OP_MAT, [self.ket_bit(w)], self.b.chain_source(seq))

    # emission ----------------------------------------------------------------
    def emit_const(self, bits: List[int], m: np.ndarray) -> None:
        """Constant matrix on ``bits`` (bits[0] most significant)."""
        if _is_identity(m):
            return
        perm = _permutation_of(m)
        if perm is not None:
            self.b.add_op(OP_PERM, bits, aux=self.b.add_real(perm.astype(np.float64)))
            return
        if _is_diagonal(m):
            src = self.b.add_source(SRC_CONST, len(bits), a0=self.b.add_complex(np.diag(m)),
                                    flags=2)
            self.b.add_op(OP_DIAG, bits, src)
            return
        if m.shape == (4, 4):
            split = _controlled_split(ConstMat(m))
            if split is not None:
                pos, sub = split
                self.b.add_op(OP_CTRL1, [bits[pos], bits[1 - pos]], self.b.source_of(sub))
                return
        self.b.add_op(OP_MAT, bits, self.b.source_of(ConstMat(m)))

    def emit_elementary(self, bits: List[int], spec) -> None:
        if isinstance(spec, ConstMat):
            self.emit_const(bits, spec.matrix)
            return
        if isinstance(spec, DiagPhaseMat):
            src = self.b.add_source(SRC_DIAGPH, len(bits), a0=self.b.add_real(spec.marks),
                                    angle=self.b.add_angle(spec.theta))
            self.b.add_op(OP_DIAG, bits, src)
            return
        if isinstance(spec, TrigMat) and spec.C0.shape == (4, 4):
            split = _controlled_split(spec)
            if split is not None:
                pos, sub = split
                self.b.add_op(OP_CTRL1, [bits[pos], bits[1 - pos]], self.b.source_of(sub))
                return
        self.b.add_op(OP_MAT, bits, self.b.source_of(spec))

    def emit_unitary_multi(self, wires: List[int], spec) -> None:
        factors = spec.factors if isinstance(spec, ProductMat) else [spec]
        ket = [self.ket_bit(w) for w in wires]
        for f in factors:
            self.emit_elementary(ket, f)
            if self.density:
                self.emit_elementary([self.bra_bit(w) for w in wires], _conj_spec(f))

    def add(self, wires: List[int], spec) -> None:
        if spec is None:
            return
        k = len(wires)
        if isinstance(spec, KrausSet):
            if not self.density:
                raise TypeError(
                    "noise channel cannot be applied to a pure statevector. "
                    "Use execute(type='density') instead."
                )
            S = kraus_superop(spec.ops)
            if k == 1:
                self.push_channel_1q(wires[0], S)
            else:
                self.flush(wires)
                bits = [self.ket_bit(w) for w in wires] + [self.bra_bit(w) for w in wires]
                self.emit_const(bits, S)
            return
        spec = _simplify(spec)
        if k == 1 and not isinstance(spec, DiagPhaseMat):
            if isinstance(spec, ConstMat) and _is_identity(spec.matrix):
                return
            self.push_unitary_1q(wires[0], spec)
            return
        if isinstance(spec, ConstMat) and _is_identity(spec.matrix):
            return
        self.flush(wires)
        self.emit_unitary_multi(wires, spec)

    def finish(self, n_tape_ops: int) -> Program:
        self.flush(sorted(self.pending))
        b = self.b
        pre = _hoist(b)
        ops = np.zeros(len(b.ops), dtype=OP_DTYPE)
        for i, (kind, k, src, aux, bits) in enumerate(b.ops):
            ops[i] = (kind, k, src, aux, bits)
        return Program(
            n_qubits=self.n,
            n_bits=self.N,
            density=self.density,
            ops=ops,
            sources=np.array(b.sources, dtype=SRC_DTYPE),
            items=np.array(b.items, dtype=np.int32),
            angles=np.array(b.angles, dtype=ANGLE_DTYPE),
            terms=np.array(b.terms, dtype=TERM_DTYPE),
            consts=np.array(b.consts, dtype=np.float64),
            n_args=b.max_arg + 1,
            pre=np.array(pre, dtype=PRE_DTYPE),
            n_tape_ops=n_tape_ops,
        )


def compile_tape(tape: List[Operation], n_qubits: int, density: bool) -> Program:
    """Lower a recorded tape.  ``density`` selects rho-as-vector evolution (needed
    iff the tape holds a KrausChannel, simulation.py:42-57,176-181)."""
    low = _Lowerer(n_qubits, density)
    for op in tape:
        if isinstance(op, Barrier):
            continue
        for w in op.wires:
            if w >= n_qubits or w < 0:
                raise CompileError(f"{op.name} acts on wire {w} outside 0..{n_qubits - 1}")
        low.add(list(op.wires), op.spec())
    return low.finish(len(tape))


def compile_elementwise(per_element_tapes: List[List[Operation]], n_qubits: int,
                        density: bool, table_arg: int):
    """Fallback for circuits the affine recorder cannot follow: every element was
    recorded concretely; ops whose matrices differ across the batch read their
    matrix from a per-element table (argument slot ``table_arg``).

    Returns ``(program, table)`` with ``table`` a (B, width) float64 array of
    interleaved complex matrices."""
    first = per_element_tapes[0]
    B = len(per_element_tapes)
    for t in per_element_tapes:
        if len(t) != len(first) or any(
            type(a) is not type(b) or a.wires != b.wires for a, b in zip(t, first)
        ):
            raise CompileError("batch elements recorded structurally different circuits")
    low = _Lowerer(n_qubits, density)
    cols: List[np.ndarray] = []
    width = 0
    for j, op in enumerate(first):
        if isinstance(op, Barrier):
            continue
        if isinstance(op, KrausChannel):
            low.add(list(op.wires), op.spec())
            continue
        mats = np.stack([t[j].matrix for t in per_element_tapes])
        if np.all(mats == mats[0]):
            low.add(list(op.wires), ConstMat(mats[0]))
            continue
        k = len(op.wires)
        spec = _TableMat(k, table_arg, width)
        cols.append(mats.reshape(B, -1))
        width += mats[0].size
        if k == 1:
            low.push_unitary_1q(op.wires[0], spec)
        else:
            low.flush(op.wires)
            ket = [low.ket_bit(w) for w in op.wires]
            low.b.add_op(OP_MAT, ket, low.b.source_of(spec))
            if density:
                low.b.add_op(OP_MAT, [low.bra_bit(w) for w in op.wires],
                             low.b.source_of(_conj_spec(spec)))
    prog = low.finish(len(first))
    if cols:
        table = np.concatenate(cols, axis=1).astype(np.complex128).view(np.float64)
    else:
        table = np.zeros((B, 2), dtype=np.float64)
    prog.n_args = max(prog.n_args, table_arg + 1)
    return prog, np.ascontiguousarray(table)


# ---------------------------------------------------------------------------
# observables
# ---------------------------------------------------------------------------
def compile_observables(obs: List[Operation], n_qubits: int):
    """Observable table + extra constants.  Z-strings and other diagonal
    observables reduce over probabilities (the reference's fast path,
    simulation.py:251-261, extended to parity observables); anything else takes
    the dense path (simulation.py:266-269, :312)."""
    recs = np.zeros(len(obs), dtype=OBS_DTYPE)
    pool: List[float] = []

    def add_complex(a):
        a = np.asarray(a, dtype=np.complex128).ravel()
        if len(pool) % 2:
            pool.append(0.0)
        off = len(pool) // 2
        pool.extend(a.view(np.float64).tolist())
        return off

    for i, ob in enumerate(obs):
        wires = list(ob.wires)
        k = len(wires)
        if k > MAX_OP_BITS:
            raise CompileError(f"observable on {k} wires exceeds {MAX_OP_BITS}")
        m = np.asarray(ob.matrix, dtype=np.complex128)
        bits = [n_qubits - 1 - w for w in wires] + [0] * (MAX_OP_BITS - k)
        if _is_diagonal(m):
            d = np.diag(m)
            zdiag = np.array([(-1.0) ** bin(j).count("1") for j in range(2**k)])
            if np.array_equal(d, zdiag):
                zmask = 0
                for w in wires:
                    zmask |= 1 << (n_qubits - 1 - w)
                recs[i] = (OBS_ZSTRING, k, 0, 0, zmask, bits)
            else:
                recs[i] = (OBS_DIAG, k, add_complex(d), 0, 0, bits)
        else:
            recs[i] = (OBS_DENSE, k, add_complex(m), 0, 0, bits)
    return recs, np.array(pool, dtype=np.float64)
