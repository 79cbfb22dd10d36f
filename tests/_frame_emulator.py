"""Test-only NumPy emulation of the on-chip frame engine's STEP PROGRAM
(``qml_essentials_b200/csrc/qmlb_frame_types.h``).

The host planner (``qmlb_frame_plan.cu``) is pure host code and its output is dumped by
``qmlb_plan_describe``; this module parses that dump and executes the steps exactly as the
kernel ``k_frame`` does - items = indices with zeros at the pivots, slots at
``base ^ eoff[v]``, XOR-variant matrices selected by the parity rows, relayouts as GF(2)
linear gathers over (rank, tile index) - with the matrices of the oracle's program
interpreter.  It lets the CPU suite prove the frame bookkeeping (folded CX, masks, parity
rows, cluster exchanges) without a GPU; the CUDA kernel itself is covered by the ``-m gpu``
parity tests.  Never reachable from the product."""

import numpy as np

from oracle import program_interp as pi

FOP_MAT1, FOP_MAT2, FOP_MATK, FOP_CTRL1, FOP_DIAG, FOP_SIGN = 0, 1, 2, 3, 4, 5

# Pauli transfer matrices: rows (I, Z, X, Y), columns (rho00, rho01, rho10, rho11)
_T = np.array([[1, 0, 0, 1], [1, 0, 0, -1], [0, 1, 1, 0], [0, 1j, -1j, 0]], dtype=np.complex128)
_TINV = _T.conj().T / 2
R, D = 4, 16


def parse(text):
    lines = text.strip().split("\n")
    assert lines[0] in ("strategy 3", "strategy 4", "strategy 5"), lines[0]
    geo = {"passes": []}
    steps = []
    lanes = []
    relayouts = []
    for line in lines[1:]:
        tok = line.split()
        if tok[0] == "frame":
            geo.update({tok[i]: int(tok[i + 1]) for i in range(1, len(tok), 2)})
            continue
        if tok[0] == "fstream":
            geo["low_bits"] = int(tok[2])
            geo["final_hpos"] = [int(x) for x in tok[4:]]
            continue
        if tok[0] == "pass":
            it, io = tok.index("tp"), tok.index("opos")
            geo["passes"].append(dict(init=int(tok[2]), first=int(tok[4]), steps=int(tok[6]),
                                      tp=[int(x) for x in tok[it + 1:io]],
                                      opos=[int(x) for x in tok[io + 1:]]))
            continue
        if tok[0] == "relayout":
            local = tok[1] == "local"
            ih = tok.index("htab") if "htab" in tok else len(tok)
            steps.append(("relayout", [int(x) for x in tok[(2 if local else 1):ih]], local))
            relayouts.append(dict(qcol=steps[-1][1], local=local,
                                  htab=[int(x) for x in tok[ih + 1:]]))
            continue
        assert tok[0] == "subpass"
        ip, ie, ipar, io = (tok.index(k) for k in ("pivots", "eoff", "par", "ops"))
        piv = [int(x) for x in tok[ip + 1:ie]]
        eoff = [int(x) for x in tok[ie + 1:ipar]]
        par = [tuple(int(y) for y in x.split(":")) for x in tok[ipar + 1:io]]
        ops = []
        for ent in tok[io + 1:]:
            f = ent.split(":")
            rec = dict(zip(("index", "code", "k", "j0", "j1", "has_c", "flags", "premat_off",
                            "smem_off", "shape"), (int(x) for x in f[:10])))
            rec["idx"] = [int(x) for x in f[10].split(",")] if len(f) > 10 else []
            rec["table"] = [int(x) for x in f[11].split(",")] if len(f) > 11 else []
            ops.append(rec)
        steps.append(("subpass", piv, eoff, par, ops, int(tok[tok.index("fast") + 1])))
        if "ipos" in tok:
            ii, ik, isg = tok.index("ipos"), tok.index("kd"), tok.index("sg")
            lanes.append(dict(ok=int(tok[tok.index("lanes_ok") + 1]),
                              ipos=[int(x) for x in tok[ii + 1:ik]],
                              kd=[int(x) for x in tok[ik + 1:isg]],
                              sg=[int(x) for x in tok[isg + 1:ip]], piv=piv, eoff=eoff, par=par))
            if steps[-1][5] >= 128:
                _check_ptm_fast(steps[-1][5], ops, lanes[-1]["sg"])
    geo["lanes"] = lanes
    geo["relayouts"] = relayouts
    return geo, steps


def relayout_wavefronts(geo, elem_bytes, swizzled):
    """Shared-memory wavefronts (gather + store) of the tile-local relayouts relative to the
    ideal, with the planner's lane table (thread copies d = d_std ^ htab[d_std & mask])."""
    nb = {16: 3, 8: 4, 4: 5}[elem_bytes]
    per = 1 << nb
    T = geo["tile_bits"]
    tmask = (1 << T) - 1

    def bank(x):
        x &= tmask
        if not swizzled:
            return x & (per - 1)
        b = 0
        while x:
            b ^= x & (per - 1)
            x >>= nb
        return b

    got = ideal = 0
    for r in geo["relayouts"]:
        assert all(h & (per - 1) == 0 for h in r["htab"][:per]), "htab keeps the lane bits"
        for start in range(0, min(1 << T, 1024), per):
            ds = [d ^ r["htab"][d & (per - 1)] for d in range(start, start + per)]
            assert sorted(d & (per - 1) for d in ds) == list(range(per))
            srcs = []
            for d in ds:
                s_ = 0
                for q in range(T):
                    if d >> q & 1:
                        s_ ^= r["qcol"][q]
                srcs.append(s_ & tmask)
            for addrs in (ds, srcs):
                banks = {}
                for a in addrs:
                    banks.setdefault(bank(a), set()).add(a)
                got += max(len(v) for v in banks.values())
                ideal += 1
    return got / max(ideal, 1)


def check_item_maps(geo):
    """FrameSubX of every sub-pass: ipos is a bijection of the item bits onto the non-pivot
    tile positions and kd[b] is the address delta of item bit team_bits + b."""
    T, tb = geo["tile_bits"], geo["team_bits"]
    for ln in geo["lanes"]:
        free = [q for q in range(T) if q not in ln["piv"]]
        assert sorted(ln["ipos"]) == free
        for b in range(4):
            want = 0
            if tb + b < len(free):
                want = item_address(ln, 1 << (tb + b))
            assert ln["kd"][b] == want


def item_address(ln, it):
    """Tile index of slot 0 of item `it` (cluster rank 0): scatter of the item bits to ipos,
    then the shift eoff[c] that makes slot v hold logical value v."""
    base = 0
    for b, q in enumerate(ln["ipos"]):
        base |= ((it >> b) & 1) << q
    c = 0
    for j in range(R):
        c |= (bin(base & ln["par"][j][0]).count("1") & 1) << j
    return base ^ ln["eoff"][c]


def smem_wavefronts(geo, elem_bytes):
    """Shared-memory wavefronts of the tile loads of every sub-pass relative to the ideal,
    in the swizzled tile of the Pauli-basis kernel (bank = XOR of the address digits)."""
    nb = {8: 4, 4: 5}[elem_bytes]
    lanes_per = 1 << nb
    got = ideal = 0
    for ln in geo["lanes"]:
        for warp in range(2):
            ads = [item_address(ln, warp * 32 + l) for l in range(32)]
            for v in range(D):
                for h in range(0, 32, lanes_per):
                    banks = {}
                    for a in ads[h:h + lanes_per]:
                        a ^= ln["eoff"][v]
                        bank, x = 0, a
                        while x:
                            bank ^= x & (lanes_per - 1)
                            x >>= nb
                        banks.setdefault(bank, set()).add(a)
                    got += max(len(x) for x in banks.values())
                    ideal += 1
    return got / max(ideal, 1)


def _check_ptm_fast(fast, ops, sg):
    """Pauli-basis straight-line body: [signs sg[0:2]] A on (1,0) [signs sg[2:4]] B on (3,2)
    must replay the ops of the step in their order."""
    a, b = (fast - 128) // 3, (fast - 128) % 3
    slot, seq = 0, []
    for o in ops:
        seq.append((slot, o))
        slot += {FOP_SIGN: 4, FOP_DIAG: 2}.get(o["code"], 1)
    want = []
    by_slot = dict(seq)
    for s_ in sg[:2]:
        if s_ != 255:
            want.append(by_slot[s_])
    mats = [o for _, o in seq if o["code"] == FOP_MAT2]
    A = [o for o in mats if (o["j0"], o["j1"]) == (1, 0)]
    B = [o for o in mats if (o["j0"], o["j1"]) == (3, 2)]
    assert len(A) == (a != 0) and len(B) == (b != 0) and len(mats) == len(A) + len(B)
    want += A
    for s_ in sg[2:]:
        if s_ != 255:
            want.append(by_slot[s_])
    want += B
    assert [id(o) for o in want] == [id(o) for _, o in seq], "fast body replays another order"
    for o, code in ((A, a), (B, b)):
        if o:
            assert (o[0]["shape"] == 3) == (code == 2)
    assert all(by_slot[s_]["code"] == FOP_SIGN for s_ in sg if s_ != 255)


def _check_fast(fast, ops):
    """The kernel's straight-line item bodies must describe exactly the ops of the step."""
    if fast >= 128:
        return  # Pauli-basis body: _check_ptm_fast at parse time
    if fast >= 64:
        real, mask = (fast - 64) >> 4, (fast - 64) & 15
        assert all(o["code"] == FOP_MAT1 for o in ops)
        assert sorted(o["j0"] for o in ops) == [j for j in range(4) if mask >> j & 1]
        assert not real or all(o["shape"] >= 1 for o in ops)
    elif fast >= 16:
        sa, sb = ((fast - 16) >> 2) - 1, ((fast - 16) & 3) - 1
        want = {(1, 0): sa, (3, 2): sb}
        assert all(o["code"] == FOP_MAT2 for o in ops) and len(ops) <= 2
        got = {(o["j0"], o["j1"]): o["shape"] for o in ops}
        assert got == {k: v for k, v in want.items() if v >= 0}
    else:
        assert fast == 0 or fast >= 128  # the latter: _check_ptm_fast at parse time


def _deposit(w, piv):
    for b in piv:  # ascending
        low = w & ((1 << b) - 1)
        w = ((w >> b) << (b + 1)) | low
    return w


def _parity(x):
    """Bit parity of every entry of an integer array."""
    x = np.asarray(x, dtype=np.int64).copy()
    for sh in (32, 16, 8, 4, 2, 1):
        x ^= x >> sh
    return x & 1


def emulate(prog, text, args, batch):
    """States (batch, 2^n_bits) in index order after the step program (vectorised over
    the items of a step; the arithmetic per item is what ``k_frame`` does)."""
    geo, steps = parse(text)
    T, G, N = geo["tile_bits"], geo["outer_bits"], prog.n_bits
    assert T + G == N
    interp = pi.Interp(prog, args, batch)
    mats = {}

    ptm = bool(geo.get("ptm", 0))

    def matrix(index):
        if index not in mats:
            m = interp.source(prog.ops[index]["src"])
            m = np.ascontiguousarray(np.broadcast_to(m, (batch,) + m.shape[1:]))
            if ptm:  # transfer matrix of the (ket, bra) superoperator: real
                m = _T @ m @ _TINV
                assert np.abs(m.imag).max() < 1e-12
                m = m.real.astype(np.complex128)
            mats[index] = m
        return mats[index]

    st = np.zeros((batch, 1 << N), dtype=np.complex128)
    st[:, 0] = 1.0
    if ptm:  # |0..0><0..0|: coefficient 1 wherever every x bit is 0 (initial frame)
        idx = np.arange(1 << N, dtype=np.int64)
        st[:, :] = ((idx & geo["xmask"]) == 0).astype(np.float64)[None, :]
        r = _run_steps(st, steps, T, G, N, matrix, batch)
        assert np.abs(r.imag).max() == 0.0
        n = N // 2
        arr = r.reshape((batch,) + (2,) * N)
        tin = _TINV.reshape(2, 2, 2, 2)  # [k, b, x, z]
        for w in range(n):
            ax, az = 1 + w, 1 + n + w
            arr = np.moveaxis(np.tensordot(tin, arr, axes=([2, 3], [ax, az])), [0, 1], [ax, az])
        return arr.reshape(batch, 1 << N)
    if geo["passes"]:  # streamed tiles: HBM layout <-> (tile number, tile index) per pass
        idx = np.arange(1 << N, dtype=np.int64)
        for ps in geo["passes"]:
            assert ps["tp"][:geo["low_bits"]] == list(range(geo["low_bits"]))
            assert sorted(ps["tp"] + ps["opos"]) == list(range(N))
            hbm = np.zeros(1 << N, dtype=np.int64)
            for i, b in enumerate(ps["tp"] + ps["opos"]):
                hbm |= ((idx >> i) & 1) << b
            virt = st[:, hbm]
            virt = _run_steps(virt, steps[ps["first"]:ps["first"] + ps["steps"]], T, G, N,
                              matrix, batch, local_only=True)
            st = np.empty_like(virt)
            st[:, hbm] = virt
        logical = np.zeros(1 << N, dtype=np.int64)  # HBM index of logical index l
        for j, h in enumerate(geo["final_hpos"]):
            logical |= ((idx >> j) & 1) << h
        return st[:, logical]
    return _run_steps(st, steps, T, G, N, matrix, batch)


def _run_steps(st, steps, T, G, N, matrix, batch, local_only=False):
    tile_mask = (1 << T) - 1
    items = np.arange(1 << (T - R), dtype=np.int64)
    for step in steps:
        if step[0] == "relayout":
            qcol = step[1]
            src = np.zeros(1 << N, dtype=np.int64)
            d = np.arange(1 << N)
            for b in range(N):
                src ^= np.where((d >> b) & 1, qcol[b], 0)
            assert np.array_equal(np.sort(src), d), "relayout is not a permutation"
            assert step[2] or not local_only, "a streamed pass can only shuffle inside the tile"
            if step[2]:
                assert np.array_equal(src >> T, d >> T), "tile-local shuffle crosses CTAs"
            elif G > 0:
                # exchange of bit positions: 32 consecutive destinations read 32 consecutive
                # sources (coalesced distributed-shared-memory traffic)
                assert np.array_equal(src & 31, d & 31), "cluster exchange is not coalesced"
            st = st[:, src]
            continue
        _, piv, eoff, par, ops, fast = step
        _check_fast(fast, ops)
        assert sorted(piv) == piv and len(set(piv)) == R
        base = _deposit(items.copy(), piv)
        assert base.max() <= tile_mask
        new = st.copy()
        touched = np.zeros(1 << N, dtype=np.int32)
        for rank in range(1 << G):
            slots = (rank << T) | (base[:, None] ^ np.asarray(eoff, dtype=np.int64)[None, :])
            np.add.at(touched, slots.ravel(), 1)

            def par_at(pi_):
                rloc, rout, _ = par[pi_]
                return _parity(base & rloc) ^ (bin(rank & rout).count("1") & 1)

            # slot 0 would hold local value c: start the item at base ^ eoff[c] instead
            c0 = sum(par_at(j) << j for j in range(R))
            ebase = base ^ np.asarray(eoff, dtype=np.int64)[c0]
            slots = (rank << T) | (ebase[:, None] ^ np.asarray(eoff, dtype=np.int64)[None, :])
            S = new[:, slots]

            def par_at(pi_, _b=ebase):  # noqa: F811  (parities of the shifted item base)
                rloc, rout, _ = par[pi_]
                return _parity(_b & rloc) ^ (bin(rank & rout).count("1") & 1)

            cj = [par_at(j) for j in range(R)]
            assert not any(c.any() for c in cj), "shifted items must see unflipped values"
            for o in ops:
                code, k = o["code"], o["k"]
                M = matrix(o["index"]) if code != FOP_SIGN else None
                if code == FOP_MAT2 and o["shape"] == 3:  # Pauli basis: diagonal transfer matrix
                    off = M - np.einsum("bii->bi", M)[:, :, None] * np.eye(4)[None]
                    assert np.abs(off).max() < 1e-14, "transfer matrix claimed diagonal"
                elif o["shape"] >= 1 and code != FOP_SIGN:  # structural claims about the matrix
                    assert np.abs(M.imag).max() == 0.0, "matrix claimed real"
                if o["shape"] == 2 and code != FOP_SIGN:
                    for v in range(4):
                        for u in range(4):
                            if v != u and v != (u ^ 3):
                                assert np.abs(M[:, v, u]).max() == 0.0, "matrix claimed X-shaped"
                if code == FOP_MAT1:
                    j = o["j0"]
                    c = cj[j]
                    assert o["has_c"] or not cj[j].any()
                    out = S.copy()
                    for v in range(D):
                        lv = ((v >> j) & 1) ^ c
                        acc = 0
                        for u in range(2):
                            acc = acc + M[:, lv, u ^ c] * S[:, :, (v & ~(1 << j)) | (u << j)]
                        out[:, :, v] = acc
                    S = out
                elif code == FOP_MAT2:
                    ja, jb = o["j0"], o["j1"]
                    assert ja > jb
                    c = (cj[ja] << 1) | cj[jb]
                    assert o["has_c"] or not c.any()

                    def sw(x):
                        return ((x & 1) << 1) | (x >> 1) if o["flags"] & 1 else x

                    out = S.copy()
                    for v in range(D):
                        lv = ((((v >> ja) & 1) << 1) | ((v >> jb) & 1)) ^ c
                        acc = 0
                        for u in range(4):
                            slot = (v & ~((1 << ja) | (1 << jb))) | ((u >> 1) << ja) | ((u & 1) << jb)
                            acc = acc + M[:, sw(lv), sw(u ^ c)] * S[:, :, slot]
                        out[:, :, v] = acc
                    S = out
                elif code == FOP_MATK:
                    c = sum(cj[j] << j for j in range(k))
                    out = S.copy()
                    dd = 1 << k
                    for v in range(D):
                        blk, lv = v >> k, v & (dd - 1)
                        acc = 0
                        for u in range(dd):
                            acc = acc + M[:, lv ^ c, u ^ c] * S[:, :, (blk << k) | u]
                        out[:, :, v] = acc
                    S = out
                elif code == FOP_CTRL1:
                    j = o["j0"]
                    c = cj[j]
                    assert o["has_c"] or not c.any()
                    ctl = par_at(o["j1"])
                    sm = par[o["j1"]][2]
                    out = S.copy()
                    for v in range(D):
                        assert ((sm >> v) & 1) == ((sm >> (v ^ (1 << j))) & 1)
                        on = (((sm >> v) & 1) ^ ctl).astype(bool)
                        lv = ((v >> j) & 1) ^ c
                        acc = 0
                        for u in range(2):
                            acc = acc + M[:, lv, u ^ c] * S[:, :, (v & ~(1 << j)) | (u << j)]
                        out[:, :, v] = np.where(on[None, :], acc, S[:, :, v])
                    S = out
                elif code == FOP_SIGN:
                    out = S.copy()
                    mask = o["premat_off"]
                    lb = np.zeros_like(base)  # local value at slot 0: indexes the sign words
                    for a, pidx in enumerate(o["idx"]):
                        lb |= par_at(pidx) << (k - 1 - a)
                    words = np.asarray(o["table"])[lb]
                    for v in range(D):
                        loc = np.zeros_like(base)
                        for a, pidx in enumerate(o["idx"]):
                            bit = par_at(pidx) ^ ((par[pidx][2] >> v) & 1)
                            loc |= bit << (k - 1 - a)
                        neg = (mask >> loc) & 1
                        assert np.array_equal(neg, (words >> v) & 1), "planner's sign words"
                        out[:, :, v] = np.where(neg.astype(bool)[None, :], -S[:, :, v], S[:, :, v])
                    S = out
                elif code == FOP_DIAG:
                    out = S.copy()
                    for v in range(D):
                        loc = np.zeros_like(base)
                        for a, pidx in enumerate(o["idx"]):
                            bit = par_at(pidx) ^ ((par[pidx][2] >> v) & 1)
                            loc |= bit << (k - 1 - a)
                        out[:, :, v] = M[:, loc] * S[:, :, v]
                    S = out
                else:
                    raise ValueError(code)
            new[:, slots] = S
        assert (touched == 1).all(), "items do not tile the state"
        st = new
    return st


class FrameEmuExecutor:
    """Executor for ``script._set_executor_for_testing``: programs the planner assigns to
    the frame engine run through :func:`emulate`, everything else through the plain
    program interpreter."""

    name = "frame-emulator"

    def __init__(self, lib):
        self.lib = lib
        self.frame_runs = 0
        self.steps = []

    def execute(self, plan, host_args, batch, chunk=None, to_host=True):
        from qml_essentials_b200 import backend

        args = [(a[0], a[1], a[2]) if a is not None else (None, 1, 1) for a in host_args]
        text = backend.plan_describe(self.lib, plan.program, plan.out_type, plan.obs_recs,
                                     plan.obs_pool, plan.precision)
        if text.startswith(("strategy 3", "strategy 4", "strategy 5")):
            st = emulate(plan.program, text, args, batch)
            self.frame_runs += 1
            self.steps.append(text)
        else:
            st = pi.Interp(plan.program, args, batch).run()
        names = {0: "state", 1: "probs", 2: "expval", 3: "density"}
        return pi.measure(plan.program, st, names[plan.out_type], plan.obs_recs, plan.obs_pool)
