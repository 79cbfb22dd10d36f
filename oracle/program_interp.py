"""NumPy interpreter of the compiled device program (oracle; test-only).

Executes the flat program emitted by ``qml_essentials_b200.compiler`` with plain
bit arithmetic, following the semantics documented in ``include/qmlb200.h``.  It
exists so that the CPU test-suite can check the tape->program compiler (fusion,
classification, density lowering, batch-factor indexing) against the
reference-faithful einsum oracle without a GPU.  It is NOT reachable from the
product: ``Script.execute`` only ever calls the CUDA library.
"""

import numpy as np

OP_MAT, OP_CTRL1, OP_PERM, OP_DIAG = 0, 1, 2, 3
SRC_CONST, SRC_TRIG, SRC_CHAIN, SRC_DIAGPH, SRC_TABLE, SRC_SUPER, SRC_PRE = 0, 1, 2, 3, 4, 5, 6
OBS_ZSTRING, OBS_DIAG, OBS_DENSE = 0, 1, 2


class Interp:
    def __init__(self, prog, args, batch):
        """``args``: list of ``(array2d, div, mod)`` per argument slot; element ``b``
        reads row ``(b // div) % mod``."""
        self.p = prog
        self.B = batch
        self.args = args
        self.cpool = prog.consts

    def cconst(self, off, n):
        return self.cpool[2 * off : 2 * off + 2 * n].view(np.complex128)

    def arg_rows(self, a):
        arr, div, mod = self.args[a]
        idx = (np.arange(self.B) // div) % mod
        return np.asarray(arr)[idx]

    def angle(self, aid):
        a = self.p.angles[aid]
        th = np.full(self.B, a["c0"], dtype=np.float64)
        for t in self.p.terms[a["first"] : a["first"] + a["n"]]:
            th = th + t["coeff"] * self.arg_rows(t["arg"])[:, t["offset"]]
        return th

    def source(self, sid):
        """(B or 1, d, d) complex matrices (or (B or 1, d) for diagonal sources)."""
        s = self.p.sources[sid]
        kind, k = s["kind"], s["k"]
        d = 2**k
        if kind == SRC_CONST:
            if s["flags"] & 2:
                return self.cconst(s["a0"], d)[None, :]
            return self.cconst(s["a0"], d * d).reshape(1, d, d)
        if kind == SRC_TRIG:
            th = self.angle(s["angle"]) * s["kappa"]
            C0 = self.cconst(s["a0"], d * d).reshape(d, d)
            A = self.cconst(s["a1"], d * d).reshape(d, d)
            Bm = self.cconst(s["a2"], d * d).reshape(d, d)
            return C0[None] + np.cos(th)[:, None, None] * A + np.sin(th)[:, None, None] * Bm
        if kind == SRC_CHAIN:
            ids = self.p.items[s["a0"] : s["a0"] + s["a1"]]
            m = self.source(ids[0])
            for i in ids[1:]:
                m = self.source(i) @ m
            return m
        if kind == SRC_DIAGPH:
            marks = self.cpool[s["a0"] : s["a0"] + d]
            return np.exp(-1j * marks[None, :] * self.angle(s["angle"])[:, None])
        if kind == SRC_TABLE:
            rows = self.arg_rows(s["a0"])
            m = rows[:, 2 * s["a1"] : 2 * s["a1"] + 2 * d * d].copy().view(np.complex128)
            m = m.reshape(self.B, d, d)
            return m.conj() if s["flags"] & 1 else m
        if kind == SRC_PRE:
            # hoisted factor: same value as the source it was hoisted from, evaluated for
            # the row of argument a1 this element reads (table lookup on the device)
            pre = self.p.pre[s["a2"]]
            assert pre["arg"] == s["a1"] and pre["local"] == s["a0"]
            return self.source(pre["src"])
        if kind == SRC_SUPER:
            ids = self.p.items[s["a0"] : s["a0"] + s["a1"]]
            S = np.eye(4, dtype=np.complex128)[None]
            for i in ids:
                if self.p.sources[i]["k"] == 2:
                    S = self.source(i) @ S
                else:
                    U = self.source(i)
                    UU = np.einsum("bij,bkl->bikjl", U, U.conj()).reshape(-1, 4, 4)
                    S = UU @ S
            return S
        raise ValueError(kind)

    def gather_index(self, bits):
        """(2^(N-k), 2^k) table of state indices: row = group, col = local value
        with bits[0] most significant."""
        N = self.p.n_bits
        k = len(bits)
        rest = [b for b in range(N) if b not in bits]
        base = np.zeros(2 ** (N - k), dtype=np.int64)
        g = np.arange(2 ** (N - k))
        for j, b in enumerate(rest):
            base |= ((g >> j) & 1) << b
        loc = np.zeros(2**k, dtype=np.int64)
        v = np.arange(2**k)
        for j, b in enumerate(bits):
            loc |= ((v >> (k - 1 - j)) & 1) << b
        return base[:, None] | loc[None, :]

    def run(self, state=None):
        """Evolve |0..0> (or ``state``, shape (B, 2^n_bits), continued in place)."""
        p = self.p
        if state is None:
            st = np.zeros((self.B, 2**p.n_bits), dtype=np.complex128)
            st[:, 0] = 1.0
        else:
            st = state
        for op in p.ops:
            k = op["k"]
            bits = list(op["bits"][:k])
            idx = self.gather_index(bits)  # (G, 2^k)
            if op["kind"] == OP_MAT:
                m = self.source(op["src"])
                x = st[:, idx]  # (B, G, 2^k)
                st[:, idx] = np.einsum("bij,bgj->bgi", np.broadcast_to(
                    m, (self.B,) + m.shape[1:]), x)
            elif op["kind"] == OP_CTRL1:
                m = self.source(op["src"])
                x = st[:, idx]
                sub = x[:, :, 2:]
                x[:, :, 2:] = np.einsum("bij,bgj->bgi", np.broadcast_to(
                    m, (self.B,) + m.shape[1:]), sub)
                st[:, idx] = x
            elif op["kind"] == OP_PERM:
                perm = p.consts[op["aux"] : op["aux"] + 2**k].astype(np.int64)
                st[:, idx] = st[:, idx][:, :, perm]
            elif op["kind"] == OP_DIAG:
                dvec = self.source(op["src"])
                st[:, idx] = st[:, idx] * np.broadcast_to(
                    dvec, (self.B, dvec.shape[-1]))[:, None, :]
            else:
                raise ValueError(op["kind"])
        return st


def measure(prog, st, out_type, obs=None, obs_pool=None):
    """Apply the measurement of ``include/qmlb200.h`` to interpreter states."""
    n, B = prog.n_qubits, st.shape[0]
    dim = 2**n
    if prog.density:
        rho = st.reshape(B, dim, dim)
        probs = np.real(np.einsum("bii->bi", rho))
    else:
        rho = None
        probs = np.abs(st) ** 2
    if out_type == "state":
        if prog.density:
            raise ValueError("state output of a density program")
        return st
    if out_type == "probs":
        return probs
    if out_type == "density":
        return rho if prog.density else np.einsum("bi,bj->bij", st, st.conj())
    if out_type == "expval":
        out = np.zeros((B, len(obs)))
        for j, ob in enumerate(obs):
            k = ob["k"]
            bits = list(ob["bits"][:k])
            if ob["kind"] == OBS_ZSTRING:
                sign = np.array(
                    [(-1.0) ** bin(i & int(ob["zmask"])).count("1") for i in range(dim)])
                out[:, j] = probs @ sign
                continue
            loc = np.zeros(dim, dtype=np.int64)
            for t, b in enumerate(bits):
                loc |= ((np.arange(dim) >> b) & 1) << (k - 1 - t)
            if ob["kind"] == OBS_DIAG:
                d = obs_pool[2 * ob["a0"] : 2 * ob["a0"] + 2 * 2**k].view(np.complex128)
                out[:, j] = probs @ np.real(d[loc])
                continue
            O = obs_pool[2 * ob["a0"] : 2 * ob["a0"] + 2 * 4**k].view(
                np.complex128).reshape(2**k, 2**k)
            mask = sum(1 << b for b in bits)
            i = np.arange(dim)
            full = np.zeros((dim, dim), dtype=np.complex128)
            for r in range(dim):
                same_rest = (i & ~mask) == (r & ~mask)
                full[r, same_rest] = O[loc[r], loc[i[same_rest]]]
            if prog.density:
                out[:, j] = np.real(np.einsum("ij,bji->b", full, rho))
            else:
                out[:, j] = np.real(np.einsum("bi,ij,bj->b", st.conj(), full, st))
        return out
    raise ValueError(out_type)
