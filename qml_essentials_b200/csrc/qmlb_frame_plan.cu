// Host planner of the on-chip frame engine (strategy 3): cuts a device program into
// SUBPASS / RELAYOUT steps (qmlb_frame_types.h).  Pure host code - no CUDA call - so the
// CPU test-suite can check every schedule through qmlb_plan_describe.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "qmlb_internal.h"

namespace qmlb {

namespace {

inline int popc64(uint64_t x) { return __builtin_popcountll(x); }
inline int top_bit(uint64_t x) { return 63 - __builtin_clzll(x); }

// physical = A * logical; col[j] = A e_j (mask of logical bit j), row[j] = row j of A^-1
// (parity row: the logical value of bit j at physical index p is parity(p & row[j])).
// Invariant: every outer physical position g >= T holds exactly one un-mixed logical bit
// (row g of A is a unit vector), so "logical bit j is local" <=> col[j] has no outer bit.
struct Frame {
  int N = 0, T = 0;
  uint64_t col[FRAME_MAX_BITS] = {}, row[FRAME_MAX_BITS] = {};
  bool outer(int j) const { return (col[j] >> T) != 0; }
  void set_permutation(const std::vector<int>& pos) {  // logical j -> position pos[j]
    for (int j = 0; j < N; ++j) col[j] = row[j] = 1ull << pos[j];
  }
};

struct OpInfo {
  uint64_t bits = 0;   // every logical bit the op touches
  uint64_t need = 0;   // bits that must be local (and, for gates, in the register group)
  bool fold = false;   // GF(2)-linear permutation: folded into the frame
  int entries = 0;     // base matrix entries (premat row)
  std::vector<int> pimg, pinv_img;  // fold: images of the unit vectors under perm / perm^-1
};

int base_entries(const qmlb_program* p, const qmlb_op& o) {
  if (o.kind == QMLB_OP_PERM || o.kind == QMLB_OP_SIGN) return 0;
  const qmlb_source& s = p->sources[o.src];
  return source_entries(s.kind, s.k, s.flags);
}

// local bit b of an op (b = 0 least significant) is logical bit bits[k-1-b]
inline int op_logical(const qmlb_op& o, int b) { return o.bits[o.k - 1 - b]; }

bool analyse(const qmlb_program* p, std::vector<OpInfo>& info) {
  info.assign(p->ops.size(), OpInfo{});
  for (size_t i = 0; i < p->ops.size(); ++i) {
    const qmlb_op& o = p->ops[i];
    OpInfo& f = info[i];
    for (int j = 0; j < o.k; ++j) f.bits |= 1ull << o.bits[j];
    f.entries = base_entries(p, o);
    switch (o.kind) {
      case QMLB_OP_MAT:
        if (o.k > 4) return false;
        f.need = f.bits;
        break;
      case QMLB_OP_CTRL1:
        f.need = 1ull << o.bits[1];
        break;
      case QMLB_OP_DIAG:
      case QMLB_OP_SIGN:
        break;
      case QMLB_OP_PERM: {
        const int D = 1 << o.k;
        std::vector<int> perm(D), pinv(D);
        for (int v = 0; v < D; ++v) perm[v] = (int)p->consts[o.aux + v];
        for (int v = 0; v < D; ++v) {
          if (perm[v] < 0 || perm[v] >= D) return false;
          pinv[perm[v]] = v;
        }
        if (perm[0] != 0) return false;
        for (int v = 1; v < D; ++v) {  // linear: perm[v] = XOR of the images of its bits
          int acc = 0;
          for (int b = 0; b < o.k; ++b)
            if (v >> b & 1) acc ^= perm[1 << b];
          if (acc != perm[v]) return false;
        }
        f.fold = true;
        f.pimg.resize(o.k);
        f.pinv_img.resize(o.k);
        for (int b = 0; b < o.k; ++b) {
          f.pimg[b] = perm[1 << b];
          f.pinv_img[b] = pinv[1 << b];
        }
        // a bit whose value after the permutation depends on another bit ("target") must be
        // local: row b of L (bit b of every image) has to be the unit vector for outer bits
        for (int b = 0; b < o.k; ++b) {
          bool unit = true;
          for (int c = 0; c < o.k; ++c)
            if (((f.pimg[c] >> b) & 1) != (c == b ? 1 : 0)) unit = false;
          if (!unit) f.need |= 1ull << op_logical(o, b);
        }
        break;
      }
      default:
        return false;
    }
  }
  return true;
}

// A' = A * L for the linear permutation of op o
void fold_into(Frame& F, const qmlb_op& o, const OpInfo& f) {
  uint64_t ncol[QMLB_MAX_OP_BITS], nrow[QMLB_MAX_OP_BITS];
  for (int b = 0; b < o.k; ++b) {
    uint64_t c = 0, r = 0;
    for (int c2 = 0; c2 < o.k; ++c2) {
      if ((f.pimg[b] >> c2) & 1) c ^= F.col[op_logical(o, c2)];
      // row'[b] = sum_c2 (L^-1)[b][c2] row[c2], (L^-1)[b][c2] = bit b of pinv[e_c2]
      if ((f.pinv_img[c2] >> b) & 1) r ^= F.row[op_logical(o, c2)];
    }
    ncol[b] = c;
    nrow[b] = r;
  }
  for (int b = 0; b < o.k; ++b) {
    F.col[op_logical(o, b)] = ncol[b];
    F.row[op_logical(o, b)] = nrow[b];
  }
}

// Structure of a matrix source that the host can prove without evaluating angles.
// real: every entry has zero imaginary part; xshape (4x4): only entries with v == u or
// v == (u ^ 3) are non-zero (the 1-qubit Pauli / damping channels on (ket, bra)).
struct Shape {
  bool real = false, xshape = false;
};

bool const_block(const qmlb_program* p, int off, int n, bool* real, bool* xshape, int d) {
  *real = true;
  *xshape = d == 4;
  for (int i = 0; i < n; ++i) {
    const double re = p->consts[2 * (size_t)(off + i)], im = p->consts[2 * (size_t)(off + i) + 1];
    if (im != 0.0) *real = false;
    if (d == 4 && (re != 0.0 || im != 0.0)) {
      const int v = i >> 2, u = i & 3;
      if (v != u && v != (u ^ 3)) *xshape = false;
    }
  }
  return true;
}

Shape shape_of(const qmlb_program* p, int sid, int depth = 0) {
  Shape out;
  if (depth > 4) return out;
  const qmlb_source& s = p->sources[sid];
  const int d = 1 << s.k;
  switch (s.kind) {
    case QMLB_SRC_CONST: {
      if (s.flags & QMLB_FLAG_DIAGVEC) return out;
      bool r, x;
      const_block(p, s.a0, d * d, &r, &x, d);
      out.real = r;
      out.xshape = r && x && s.k == 2;
      return out;
    }
    case QMLB_SRC_TRIG: {
      const int axis = (s.flags >> QMLB_FLAG_ROT_SHIFT) & 3;
      if (s.k == 1 && axis != 0) {
        out.real = axis == 2;  // RY
        return out;
      }
      bool r0, r1, r2, x;
      const_block(p, s.a0, d * d, &r0, &x, 0);
      const_block(p, s.a1, d * d, &r1, &x, 0);
      const_block(p, s.a2, d * d, &r2, &x, 0);
      out.real = r0 && r1 && r2;
      return out;
    }
    case QMLB_SRC_PRE:
      return shape_of(p, p->pre[s.a2].src, depth + 1);
    case QMLB_SRC_CHAIN:
    case QMLB_SRC_SUPER: {
      out.real = true;
      out.xshape = s.kind == QMLB_SRC_SUPER;
      for (int t = 0; t < s.a1; ++t) {
        const int id = p->items[s.a0 + t];
        const Shape it = shape_of(p, id, depth + 1);
        out.real = out.real && it.real;
        // U (x) conj(U) of a 2x2 is X-shaped only for diagonal / anti-diagonal U: not tracked
        out.xshape = out.xshape && it.xshape && p->sources[id].k == 2;
      }
      out.xshape = out.xshape && out.real;
      return out;
    }
    default:
      return out;
  }
}

struct Builder {
  qmlb_program* p;
  std::vector<OpInfo> info;
  std::vector<char> done;
  std::vector<int> premat_off;     // per program op
  std::vector<int> matlist_index;  // per program op: its entry in p->stream_matlist
  Frame F;
  int N, T, G;
  int mat_cap;
  bool resident = false;   // matrices of the whole element stay in shared memory
  // streaming mode: only the matrices of the current HBM pass are staged (once per element
  // and pass), numbered compactly in the order the pass uses them
  bool pass_compact = false;
  int pass_used = 0, pass_used_max = 0;
  bool ptm = false;        // Pauli-basis engine: real state, 4x4 transfer matrices
  int swz_bits = 0;        // log2(elements per shared-memory wavefront) of the swizzled tile
  int bank_bits = 0;       // log2(elements per 128-byte wavefront): relayout lane choice
  int team_bits = 0;       // log2(threads per tile): item bits above belong to one thread
  std::vector<char> ptm_diag;  // per op: its transfer matrix is provably diagonal
  uint64_t init_xmask = 0;     // physical positions that hold x bits in the initial frame
  std::vector<FrameStep> steps;
  std::vector<std::vector<int>> step_ops;  // program-op indices per step (describe)

  // first not-done op that needs logical bit j local
  std::vector<int> next_need() const {
    std::vector<int> nn(N, 1 << 30);
    for (size_t i = 0; i < p->ops.size(); ++i) {
      if (done[i]) continue;
      for (int j = 0; j < N; ++j)
        if ((info[i].need >> j & 1) && nn[j] == (1 << 30)) nn[j] = (int)i;
    }
    return nn;
  }

  // permutation frame: the G bits needed last go outer, the rest fill the tile from the top
  // (soonest-needed bits highest: register groups then sit above the lane bits)
  std::vector<int> choose_positions() const {
    std::vector<int> nn = next_need();
    std::vector<int> order(N);
    for (int j = 0; j < N; ++j) order[j] = j;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return nn[a] > nn[b]; });
    std::vector<int> pos(N);
    for (int r = 0; r < G; ++r) pos[order[r]] = T + r;      // latest needed -> outer
    for (int r = G; r < N; ++r) pos[order[r]] = r - G;      // then ascending urgency upwards
    return pos;
  }

  void emit_relayout(const std::vector<int>& pos) {
    FrameStep s;
    std::memset(&s, 0, sizeof(s));
    s.kind = QMLB_FSTEP_RELAYOUT;
    for (int j = 0; j < N; ++j) s.qcol[pos[j]] = F.col[j];
    relayout_lanes(s);
    steps.push_back(s);
    step_ops.emplace_back();
    F.set_permutation(pos);
  }

  // Which destinations the lanes of one wavefront take.  A thread copies the elements
  // d = d_std ^ htab[d_std & (2^NB - 1)] (d_std = its consecutive indices; htab entries
  // have no bits below NB, so this is a bijection of the tile): lane bit i then moves along
  // u_i = e_i ^ h_i instead of e_i.  The h_i are chosen so that BOTH the destination banks
  // beta(u_i) and the source banks beta(M u_i) are independent - a conflict-free gather
  // and a conflict-free store for an arbitrary linear shuffle M (consecutive destinations
  // made 8-way conflicts of the gather typical: 134 M wavefronts for 17 M ideal in a
  // k_fstream pass).  The table lives in the step's unused op slots.
  void relayout_lanes(FrameStep& s) const {
    uint32_t* htab = reinterpret_cast<uint32_t*>(s.ops);
    const int NB = bank_bits;
    static const int lanes_on = [] {
      const char* v = std::getenv("QMLB_RELAYOUT_LANES");
      return v ? std::atoi(v) : 1;
    }();
    if (!lanes_on || NB <= 0 || T <= NB || T - NB > 12) return;
    const uint32_t tile_mask = (1u << T) - 1u, dm = (1u << NB) - 1u;
    auto beta = [&](uint32_t x) {
      x &= tile_mask;
      if (!swz_bits) return x & dm;
      uint32_t w = 0;
      for (; x; x >>= swz_bits) w ^= x & dm;
      return w;
    };
    auto src_of = [&](uint32_t u) {
      uint32_t acc = 0;
      for (int q = 0; q < T; ++q)
        if (u >> q & 1) acc ^= (uint32_t)s.qcol[q];
      return acc;
    };
    auto insert = [](std::vector<uint32_t>& basis, uint32_t w) {
      for (uint32_t b : basis)
        if ((w ^ b) < w) w ^= b;
      if (!w) return false;
      basis.push_back(w);
      std::sort(basis.rbegin(), basis.rend());
      return true;
    };
    std::vector<uint32_t> bd, bs;
    uint32_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < NB; ++i) {
      bool found = false;
      for (uint32_t c = 0; c < (1u << (T - NB)) && !found; ++c) {
        const uint32_t u = (1u << i) | (c << NB);
        std::vector<uint32_t> td = bd, ts = bs;
        if (insert(td, beta(u)) && insert(ts, beta(src_of(u)))) {
          bd = td, bs = ts;
          h[i] = c << NB;
          found = true;
        }
      }
      if (!found) {  // keep what independence the plain direction still offers
        insert(bd, beta(1u << i));
        insert(bs, beta(src_of(1u << i)));
      }
    }
    for (uint32_t l = 0; l < (1u << NB); ++l) {
      uint32_t m = 0;
      for (int i = 0; i < NB; ++i)
        if (l >> i & 1) m ^= h[i];
      htab[l] = m;
    }
  }

  // Layout change to the permutation frame `pos`.  Across a cluster it is split so that the
  // distributed-shared-memory traffic is COALESCED: (A) a tile-local shuffle undoes the
  // folded linear maps and parks the bits that leave next to where the arriving bits will
  // sit, (B) a pure exchange of bit POSITIONS between cluster rank and tile - consecutive
  // destinations then read consecutive remote addresses (runs of 2^t amplitudes).
  void relayout_to(const std::vector<int>& want, bool exact) {
    if (G == 0) {
      emit_relayout(want);
      steps.back().mat_entries = 1;
      return;
    }
    std::vector<int> victims, incoming;
    for (int j = 0; j < N; ++j) {
      if (!F.outer(j) && want[j] >= T) victims.push_back(j);
      if (F.outer(j) && want[j] < T) incoming.push_back(j);
    }
    // (B) lands the arriving bits on the TOP tile positions: the other tile bits keep their
    // relative order below them
    std::vector<int> posB(want);
    {
      std::vector<int> tile_bits;
      for (int j = 0; j < N; ++j)
        if (want[j] < T) tile_bits.push_back(j);
      auto is_in = [&](int j) {
        return std::find(incoming.begin(), incoming.end(), j) != incoming.end();
      };
      std::stable_sort(tile_bits.begin(), tile_bits.end(), [&](int a2, int b2) {
        if (is_in(a2) != is_in(b2)) return is_in(b2);  // arriving bits last = highest
        return want[a2] < want[b2];
      });
      for (size_t r = 0; r < tile_bits.size(); ++r) posB[tile_bits[r]] = (int)r;
    }
    std::vector<int> posA(posB);
    for (int j = 0; j < N; ++j)
      if (F.outer(j)) posA[j] = top_bit(F.col[j]);  // outer bits stay where they are
    for (size_t i = 0; i < victims.size() && i < incoming.size(); ++i)
      posA[victims[i]] = posB[incoming[i]];
    auto differs = [&](const std::vector<int>& x, const std::vector<int>& y) {
      for (int j = 0; j < N; ++j)
        if (x[j] != y[j]) return true;
      return false;
    };
    bool local_needed = false;
    for (int j = 0; j < N; ++j)
      if (F.col[j] != (1ull << posA[j])) local_needed = true;
    if (local_needed) {
      emit_relayout(posA);
      steps.back().mat_entries = 1;  // tile-local: CTA barriers suffice
    }
    if (differs(posA, posB)) emit_relayout(posB);
    if (exact && differs(posB, want)) {
      emit_relayout(want);
      steps.back().mat_entries = 1;
    }
  }

  int run();
  bool build_step();

  // ---- streaming mode: the state lives in HBM, a pass works on tiles of 2^T amplitudes ----
  // The HBM layout is always a bit permutation (hpos).  A tile holds the `L` lowest HBM
  // positions (contiguous runs: what the bulk copies move) plus T - L freely chosen ones;
  // the other N - T positions number the tiles ("outer", virtual positions T..N-1).
  int L = 0;
  std::vector<int> hpos;                 // logical bit -> HBM bit position
  std::vector<int> tp, opos;             // virtual tile index / outer index -> HBM position
  std::vector<FramePassHost> passes;
  std::vector<int> final_hpos;

  void start_pass(bool init) {
    std::vector<char> in_tile(N, 0);
    for (int i = 0; i < T; ++i) in_tile[tp[i]] = 1;
    opos.clear();
    for (int b = 0; b < N; ++b)
      if (!in_tile[b]) opos.push_back(b);
    std::vector<int> pos(N);
    for (int j = 0; j < N; ++j) {
      const int h = hpos[j];
      auto it = std::find(tp.begin(), tp.end(), h);
      pos[j] = it != tp.end() ? (int)(it - tp.begin())
                              : T + (int)(std::find(opos.begin(), opos.end(), h) - opos.begin());
    }
    F.set_permutation(pos);
    pass_used = 0;
    FramePassHost ps;
    ps.first_step = (int)steps.size();
    ps.n_steps = 0;
    ps.init = init;
    ps.tp = tp;
    ps.opos = opos;
    passes.push_back(ps);
  }

  // flush the frame of the current pass to a pure permutation that puts the logical bits in
  // `at[i]` (virtual tile index i -> logical bit) and record where everything now lives
  void close_pass(const std::vector<int>& at) {
    std::vector<int> pos(N);
    for (int j = 0; j < N; ++j)
      if (F.outer(j)) pos[j] = top_bit(F.col[j]);
    for (int i = 0; i < T; ++i) pos[at[i]] = i;
    bool needed = false;
    for (int j = 0; j < N; ++j)
      if (F.col[j] != (1ull << pos[j])) needed = true;
    if (needed) {
      emit_relayout(pos);
      steps.back().mat_entries = 1;
    }
    for (int i = 0; i < T; ++i) hpos[at[i]] = tp[i];
    passes.back().n_steps = (int)steps.size() - passes.back().first_step;
  }

  int run_stream();
};

// One SUBPASS (plus the folds that follow it).  Returns false when nothing could be done.
bool Builder::build_step() {
  // blocked: bits of an op that was skipped - everything later on them waits for the next
  // step.  folded: bits of the permutations collected in this step - later GATES on them see
  // the new frame (next step), but later permutations still fold now, in program order (a
  // CX ladder is one chain of folds: blocking them on each other applied one CX per step
  // and scattered the gates behind the ladder over single-gate sub-passes).
  uint64_t S = 0, blocked = 0, folded = 0;
  std::vector<int> picked;       // gate ops of the step, program order
  std::vector<int> folds;        // permutations folded after the step
  std::vector<int> pinned;       // k >= 3: register position j -> logical bit
  std::vector<int> par_bits;     // logical bits with a parity row beyond the register bits
  int slots = 0, entries = 0;
  for (size_t i = 0; i < p->ops.size(); ++i) {
    if (done[i]) continue;
    const qmlb_op& o = p->ops[i];
    const OpInfo& f = info[i];
    if ((f.bits & blocked) || (!f.fold && (f.bits & folded))) {
      blocked |= f.bits;
      continue;
    }
    bool outer_need = false;
    for (int j = 0; j < N; ++j)
      if ((f.need >> j & 1) && F.outer(j)) outer_need = true;
    if (outer_need) {
      blocked |= f.bits;
      continue;
    }
    if (f.fold) {
      folds.push_back((int)i);
      folded |= f.bits;
      continue;
    }
    // gate: register group, parity rows, slots, matrix area
    const uint64_t U = S | f.need;
    bool ok = popc64(U) <= FRAME_R;
    std::vector<int> pin_try = pinned;
    if (ok && o.kind == QMLB_OP_MAT && o.k >= 3) {
      std::vector<int> want(o.k);
      for (int b = 0; b < o.k; ++b) want[b] = op_logical(o, b);
      if (pin_try.empty())
        pin_try = want;
      else
        ok = want.size() <= pin_try.size() && std::equal(want.begin(), want.end(), pin_try.begin());
    }
    std::vector<int> par_try = par_bits;
    if (ok && (o.kind == QMLB_OP_CTRL1 || o.kind == QMLB_OP_DIAG || o.kind == QMLB_OP_SIGN)) {
      const int first = 0;
      const int last = o.kind == QMLB_OP_CTRL1 ? 1 : o.k;
      for (int a = first; a < last; ++a)
        if (std::find(par_try.begin(), par_try.end(), o.bits[a]) == par_try.end())
          par_try.push_back(o.bits[a]);
      ok = FRAME_R + (int)par_try.size() <= FRAME_MAX_PAR;
    }
    const int need_slots = o.kind == QMLB_OP_SIGN ? 4 : (o.kind == QMLB_OP_DIAG ? 2 : 1);
    const int e = resident ? 0 : f.entries;
    if (ok) ok = slots + need_slots <= FRAME_MAX_OPS && entries + e <= mat_cap;
    if (!ok) {
      blocked |= f.bits;
      continue;
    }
    S = U;
    pinned = pin_try;
    par_bits = par_try;
    slots += need_slots;
    entries += e;
    picked.push_back((int)i);
  }
  if (picked.empty() && folds.empty()) return false;

  if (!picked.empty()) {
    // register positions: pinned bits first, then the other group bits, then pads (local
    // bits outside the group, highest mask first so the lanes keep the low addresses)
    std::vector<int> gb(pinned);
    auto in_gb = [&](int j) { return std::find(gb.begin(), gb.end(), j) != gb.end(); };
    // 2-bit ops first, each on an adjacent register pair (1,0) / (3,2) in its own bit
    // order: the kernel's fast path for the (ket, bra) superoperators of density programs
    if (gb.empty())
      for (int i : picked) {
        const qmlb_op& o = p->ops[i];
        if (o.kind != QMLB_OP_MAT || o.k != 2) continue;
        const bool a = in_gb(o.bits[0]), b = in_gb(o.bits[1]);
        if (a || b || (gb.size() & 1) || gb.size() + 2 > (size_t)FRAME_R) continue;
        gb.push_back(o.bits[1]);  // least significant local bit on the lower register bit
        gb.push_back(o.bits[0]);
      }
    for (int j = 0; j < N; ++j)
      if ((S >> j & 1) && !in_gb(j)) gb.push_back(j);
    {
      std::vector<int> cand;
      for (int j = 0; j < N; ++j)
        if (!in_gb(j) && !F.outer(j)) cand.push_back(j);
      std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) {
        return top_bit(F.col[a]) > top_bit(F.col[b]);
      });
      for (int j : cand) {
        if ((int)gb.size() >= FRAME_R) break;
        gb.push_back(j);
      }
    }
    if ((int)gb.size() < FRAME_R) return false;  // tile smaller than a register group
    std::vector<int> regpos(N, -1);
    for (int j = 0; j < FRAME_R; ++j) regpos[gb[j]] = j;

    FrameStep s;
    std::memset(&s, 0, sizeof(s));
    s.kind = QMLB_FSTEP_SUBPASS;
    const uint64_t tile_mask = (1ull << T) - 1ull;
    // pivots: echelon form of the masks (leading bit = highest set bit)
    uint64_t ech[FRAME_R];
    int piv[FRAME_R];
    for (int j = 0; j < FRAME_R; ++j) {
      uint64_t v = F.col[gb[j]];
      bool changed = true;
      while (changed) {
        changed = false;
        for (int i = 0; i < j; ++i)
          if (v >> piv[i] & 1) {
            v ^= ech[i];
            changed = true;
          }
      }
      ech[j] = v;
      piv[j] = top_bit(v);
      // keep earlier vectors reduced against the new pivot
      for (int i = 0; i < j; ++i)
        if (ech[i] >> piv[j] & 1) ech[i] ^= v;
    }
    std::vector<int> sp(piv, piv + FRAME_R);
    std::sort(sp.begin(), sp.end());
    uint64_t pivmask = 0;
    for (int j = 0; j < FRAME_R; ++j) {
      s.pivots[j] = (uint32_t)sp[j];
      pivmask |= 1ull << sp[j];
    }
    for (int v = 0; v < FRAME_D; ++v) {
      uint64_t e = 0;
      for (int j = 0; j < FRAME_R; ++j)
        if (v >> j & 1) e ^= F.col[gb[j]];
      s.eoff[v] = (uint32_t)e;
    }
    auto make_par = [&](int logical) {
      FramePar q{};
      q.rloc = (uint32_t)(F.row[logical] & tile_mask);
      q.rout = (uint32_t)(F.row[logical] >> T);
      for (int v = 0; v < FRAME_D; ++v)
        if (popc64((uint64_t)s.eoff[v] & F.row[logical]) & 1) q.smask |= (uint16_t)(1u << v);
      return q;
    };
    for (int j = 0; j < FRAME_R; ++j) s.par[j] = make_par(gb[j]);
    for (size_t t = 0; t < par_bits.size(); ++t) s.par[FRAME_R + t] = make_par(par_bits[t]);
    s.n_par = FRAME_R + (int)par_bits.size();
    auto par_index = [&](int logical) {
      for (size_t t = 0; t < par_bits.size(); ++t)
        if (par_bits[t] == logical) return FRAME_R + (int)t;
      return -1;
    };
    // does register bit j ever see a flipped local value?
    auto has_c = [&](int j) {
      return (s.par[j].rloc & ~(uint32_t)pivmask) != 0 || s.par[j].rout != 0;
    };

    int slot = 0, used = 0;
    std::vector<int> ops_of_step;
    for (int i : picked) {
      const qmlb_op& o = p->ops[i];
      FrameOp fo;
      std::memset(&fo, 0, sizeof(fo));
      fo.k = (uint8_t)o.k;
      fo.premat_off = premat_off[i];
      fo.smem_off = pass_compact ? pass_used + used : (resident ? premat_off[i] : used);
      if (o.kind != QMLB_OP_DIAG && o.kind != QMLB_OP_SIGN) {
        const Shape sh = shape_of(p, o.src);
        fo.shape = sh.xshape && o.kind == QMLB_OP_MAT && o.k == 2
                       ? QMLB_FSHAPE_XREAL
                       : (sh.real ? QMLB_FSHAPE_REAL : QMLB_FSHAPE_FULL);
      }
      if (o.kind == QMLB_OP_MAT && o.k == 1) {
        fo.code = QMLB_FOP_MAT1;
        fo.j0 = (uint8_t)regpos[o.bits[0]];
        fo.has_c = has_c(fo.j0);
      } else if (o.kind == QMLB_OP_MAT && o.k == 2) {
        fo.code = QMLB_FOP_MAT2;
        const int p0 = regpos[o.bits[0]], p1 = regpos[o.bits[1]];
        fo.j0 = (uint8_t)std::max(p0, p1);
        fo.j1 = (uint8_t)std::min(p0, p1);
        fo.flags = p0 < p1 ? 1 : 0;
        fo.has_c = has_c(fo.j0) || has_c(fo.j1);
        // the evaluated matrix is stored with the roles of its two bits already swapped
        p->stream_matlist[matlist_index[i]].swap2 = (ptm ? 2 : 0) | (fo.flags & 1);
        if (ptm) fo.shape = ptm_diag[i] ? QMLB_FSHAPE_PDIAG : QMLB_FSHAPE_REAL;
      } else if (o.kind == QMLB_OP_MAT) {
        fo.code = QMLB_FOP_MATK;
      } else if (o.kind == QMLB_OP_CTRL1) {
        fo.code = QMLB_FOP_CTRL1;
        fo.j0 = (uint8_t)regpos[o.bits[1]];
        fo.j1 = (uint8_t)par_index(o.bits[0]);
        fo.has_c = has_c(fo.j0);
      } else if (o.kind == QMLB_OP_SIGN) {
        fo.code = QMLB_FOP_SIGN;
        fo.premat_off = o.aux;  // the sign mask
        uint8_t pi4[4];
        uint32_t ro = 0;
        for (int a = 0; a < 4; ++a) {
          pi4[a] = (uint8_t)par_index(o.bits[a]);
          ro |= (s.par[pi4[a]].rout & 255u) << (8 * a);
        }
        fo.j0 = pi4[0], fo.j1 = pi4[1], fo.shape = pi4[2], fo.has_c = pi4[3];
        fo.smem_off = (int32_t)ro;
      } else {
        fo.code = QMLB_FOP_DIAG;
      }
      used += info[i].entries;
      s.ops[slot++] = fo;
      if (o.kind == QMLB_OP_DIAG) {
        FrameOp aux;
        std::memset(&aux, 0, sizeof(aux));
        uint8_t* idx = reinterpret_cast<uint8_t*>(&aux);
        for (int a = 0; a < o.k; ++a) idx[a] = (uint8_t)par_index(o.bits[a]);
        s.ops[slot++] = aux;
      }
      if (o.kind == QMLB_OP_SIGN) {
        {
          const uint8_t idx[4] = {fo.j0, fo.j1, fo.shape, fo.has_c};
          uint32_t rl4[4];
          for (int a = 0; a < 4; ++a) rl4[a] = s.par[idx[a]].rloc;
          std::memcpy(&s.ops[slot++], rl4, sizeof(rl4));
          // sign words: entry lb = local value (first op bit most significant) read at slot
          // 0 of an item; bit v = sign of slot v, whose local value is lb ^ flip(v)
          uint16_t tab[16];
          for (int lb = 0; lb < 16; ++lb) {
            unsigned w = 0;
            for (int v = 0; v < FRAME_D; ++v) {
              int loc = lb;
              for (int a = 0; a < o.k; ++a)
                loc ^= (int)((s.par[idx[a]].smask >> v) & 1u) << (o.k - 1 - a);
              w |= (((unsigned)o.aux >> loc) & 1u) << v;
            }
            tab[lb] = (uint16_t)w;
          }
          std::memcpy(&s.ops[slot], tab, sizeof(tab));
          slot += 2;
        }
      }
      ops_of_step.push_back(i);
      done[i] = 1;
    }
    s.n_ops = slot;
    s.mat_entries = used;
    // item bit -> tile position (FrameSubX).  The address of item `it` is linear in its
    // bits: position q contributes Lq = (1 << q) ^ eoff[c_q] (c_q = the item shift its
    // parities cause).  In the swizzled tile the bank of an address is the XOR of its
    // swz_bits-wide digits, so the lanes of one wavefront hit distinct banks iff the bank
    // vectors of the lowest swz_bits item bits are independent: pick them greedily.
    {
      FrameSubX* sx = reinterpret_cast<FrameSubX*>(s.qcol);
      auto addr_of = [&](int q) {
        int c = 0;
        for (int j = 0; j < FRAME_R; ++j) c |= (int)((s.par[j].rloc >> q) & 1u) << j;
        return (uint32_t)(1u << q) ^ s.eoff[c];
      };
      std::vector<int> free_pos;
      for (int q = 0; q < T; ++q)
        if (!(pivmask >> q & 1)) free_pos.push_back(q);
      std::vector<int> lanes;
      if (swz_bits > 0) {
        const uint32_t dm = (1u << swz_bits) - 1u;
        std::vector<uint32_t> basis;
        for (int q : free_pos) {
          if ((int)lanes.size() == swz_bits) break;
          uint32_t w = 0;
          for (uint32_t h = addr_of(q); h; h >>= swz_bits) w ^= h & dm;
          for (uint32_t b : basis)
            if ((w ^ b) < w) w ^= b;
          if (!w) continue;
          basis.push_back(w);
          std::sort(basis.rbegin(), basis.rend());
          lanes.push_back(q);
        }
        sx->lanes_ok = (int)lanes.size() == std::min(swz_bits, (int)free_pos.size()) ? 1u : 0u;
      }
      std::vector<int> order(lanes);
      for (int q : free_pos)
        if (std::find(lanes.begin(), lanes.end(), q) == lanes.end()) order.push_back(q);
      for (size_t b = 0; b < order.size() && b < 16; ++b) sx->ipos[b] = (uint8_t)order[b];
      for (int b = 0; b < 4; ++b)
        if (team_bits > 0 && team_bits + b < (int)order.size())
          sx->kd[b] = addr_of(order[team_bits + b]);
    }
    // fast-path classification (see FrameStep::fast)
    {
      bool d2 = true, m1 = true, all_real = true;
      int sA = -1, sB = -1, mask = 0;
      for (int o = 0; o < slot; ++o) {
        const FrameOp& fo = s.ops[o];
        if (fo.code == QMLB_FOP_MAT2 && fo.j0 == 1 && fo.j1 == 0 && sA < 0) {
          sA = fo.shape;
          s.foff[0] = fo.smem_off;
        } else if (fo.code == QMLB_FOP_MAT2 && fo.j0 == 3 && fo.j1 == 2 && sB < 0) {
          sB = fo.shape;
          s.foff[1] = fo.smem_off;
        } else {
          d2 = false;
        }
        if (fo.code == QMLB_FOP_MAT1 && !(mask >> fo.j0 & 1)) {
          mask |= 1 << fo.j0;
          all_real = all_real && fo.shape != QMLB_FSHAPE_FULL;
        } else {
          m1 = false;
        }
      }
      if (ptm) {
        // [<= 2 signs] A on (1,0) [<= 2 signs] B on (3,2); anything else -> op interpreter
        FrameSubX* sx = reinterpret_cast<FrameSubX*>(s.qcol);
        std::memset(sx->sg, 0xff, sizeof(sx->sg));
        int a = 0, b = 0, stage = 0, ns[2] = {0, 0};
        bool ok = true;
        for (int o = 0; o < slot && ok; ++o) {
          const FrameOp& fo = s.ops[o];
          const int shp = fo.shape == QMLB_FSHAPE_PDIAG ? 2 : 1;
          if (fo.code == QMLB_FOP_SIGN) {
            if (stage < 2 && ns[stage] < 2)
              sx->sg[2 * stage + ns[stage]++] = (uint8_t)o;
            else
              ok = false;
            o += 3;
          } else if (fo.code == QMLB_FOP_MAT2 && fo.j0 == 1 && fo.j1 == 0 && stage == 0) {
            a = shp;
            s.foff[0] = fo.smem_off;
            stage = 1;
          } else if (fo.code == QMLB_FOP_MAT2 && fo.j0 == 3 && fo.j1 == 2 && stage <= 1) {
            b = shp;
            s.foff[1] = fo.smem_off;
            stage = 2;
          } else {
            ok = false;
          }
        }
        // signs met before any matrix while A is absent belong in front of B
        if (ok && a == 0 && ns[1] == 0) {
          sx->sg[2] = sx->sg[0], sx->sg[3] = sx->sg[1];
          sx->sg[0] = sx->sg[1] = 0xff;
        } else if (ok && a == 0) {
          ok = false;
        }
        s.fast = ok && (a || b) ? 128 + 3 * a + b : 0;
      } else if (d2 && slot > 0) {
        s.fast = 16 + 4 * (sA + 1) + (sB + 1);
      } else if (m1 && slot > 0) {
        s.fast = 64 + 16 * (all_real ? 1 : 0) + mask;
        for (int o = 0; o < slot; ++o) s.foff[s.ops[o].j0] = s.ops[o].smem_off;
      } else {
        s.fast = 0;
        for (int j = 0; j < 4; ++j) s.foff[j] = 0;
      }
    }
    if (pass_compact) {
      pass_used += used;
      pass_used_max = std::max(pass_used_max, pass_used);
    }
    steps.push_back(s);
    step_ops.push_back(ops_of_step);
  }
  for (int i : folds) {
    fold_into(F, p->ops[i], info[i]);
    done[i] = 1;
  }
  return true;
}

int Builder::run() {
  done.assign(p->ops.size(), 0);
  F.N = N;
  F.T = T;
  F.set_permutation(choose_positions());
  init_xmask = 0;
  for (int j = N / 2; j < N; ++j) init_xmask |= F.col[j];  // ket (= x) bits of a density program
  size_t remaining = p->ops.size();
  auto count_left = [&]() {
    size_t n = 0;
    for (char d : done) n += d ? 0 : 1;
    return n;
  };
  int guard = 0;
  while (remaining > 0) {
    if (!build_step()) {
      if (G == 0) return QMLB_ERR_UNSUPPORTED;  // cannot happen: nothing is ever outer
      std::vector<int> pos = choose_positions();
      bool same = true;
      for (int j = 0; j < N; ++j)
        if (F.col[j] != (1ull << pos[j])) same = false;
      if (same || ++guard > 4096) return QMLB_ERR_UNSUPPORTED;
      relayout_to(pos, false);
    }
    remaining = count_left();
  }
  // back to index order (logical bit j at position j) for the output
  bool ident = true;
  for (int j = 0; j < N; ++j)
    if (F.col[j] != (1ull << j)) ident = false;
  if (!ident) {
    std::vector<int> pos(N);
    for (int j = 0; j < N; ++j) pos[j] = j;
    relayout_to(pos, true);
  }
  return QMLB_OK;
}

int Builder::run_stream() {
  done.assign(p->ops.size(), 0);
  F.N = N;
  F.T = T;
  // initial HBM layout (|0..0> is invariant under bit permutations): soonest-needed bits
  // on the low positions, i.e. inside the first tile
  {
    std::vector<int> nn = next_need();
    std::vector<int> order(N);
    for (int j = 0; j < N; ++j) order[j] = j;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return nn[a] < nn[b]; });
    // ... soonest-needed on the HIGHEST tile positions: register groups then sit above the
    // lane bits and the shared-memory accesses of a warp stay contiguous (no bank conflicts)
    hpos.assign(N, 0);
    for (int r = 0; r < N; ++r) hpos[order[r]] = r < T ? T - 1 - r : r;
    tp.resize(T);
    for (int i = 0; i < T; ++i) tp[i] = i;
  }
  start_pass(true);
  auto left = [&]() {
    size_t n = 0;
    for (char d : done) n += d ? 0 : 1;
    return n;
  };
  int guard = 0;
  while (left() > 0) {
    if (build_step()) continue;
    if (++guard > 4096) return QMLB_ERR_UNSUPPORTED;
    // stuck: choose the next tile.  The T - L freely chosen positions can change, so at most
    // T - L bits arrive; the bits that stay are the soonest-needed ones of the current tile.
    std::vector<int> nn = next_need();
    std::vector<int> order(N);
    for (int j = 0; j < N; ++j) order[j] = j;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      if (nn[a] != nn[b]) return nn[a] < nn[b];
      return !F.outer(a) && F.outer(b);
    });
    std::vector<int> incoming, staying;
    for (int r = 0; r < N && (int)(incoming.size() + staying.size()) < T; ++r) {
      const int j = order[r];
      if (F.outer(j)) {
        if ((int)incoming.size() < T - L && nn[j] < (1 << 30)) incoming.push_back(j);
      } else {
        staying.push_back(j);
      }
    }
    for (int r = 0; r < N && (int)(incoming.size() + staying.size()) < T; ++r) {
      const int j = order[r];
      if (!F.outer(j) && std::find(staying.begin(), staying.end(), j) == staying.end())
        staying.push_back(j);
    }
    if (incoming.empty()) return QMLB_ERR_UNSUPPORTED;
    std::vector<int> leaving;
    for (int j = 0; j < N; ++j)
      if (!F.outer(j) && std::find(staying.begin(), staying.end(), j) == staying.end())
        leaving.push_back(j);
    // staying bits: the low run first, then the lowest high tile indices; leaving bits on
    // the highest tile indices (those HBM positions drop out of the next tile)
    // (the low run takes the staying bits that are needed LAST: gate bits on low tile
    // positions would make the lanes of a warp stride through shared memory)
    std::vector<int> at(T);
    for (size_t i = 0; i < staying.size(); ++i) at[i] = staying[staying.size() - 1 - i];
    for (size_t i = 0; i < leaving.size(); ++i) at[staying.size() + i] = leaving[i];
    close_pass(at);
    std::vector<int> ntp;
    for (size_t i = 0; i < staying.size(); ++i) ntp.push_back(tp[i]);
    for (int j : incoming) ntp.push_back(hpos[j]);
    std::sort(ntp.begin(), ntp.end());
    tp = ntp;
    start_pass(false);
  }
  {
    std::vector<int> at;
    for (int j = 0; j < N; ++j)
      if (!F.outer(j)) at.push_back(j);
    close_pass(at);
  }
  final_hpos = hpos;
  return QMLB_OK;
}

}  // namespace

// Streaming form of the frame engine (strategy 4): the state stays in HBM, every pass moves
// tiles of 2^T amplitudes through shared memory with bulk copies and runs the pass' steps on
// them.  Fills p->fstream_* ; QMLB_ERR_UNSUPPORTED -> fall back to the register-group stream.
int plan_frame_stream(qmlb_program* p) {
  const int N = p->n_bits;
  const size_t cs = p->dtype == QMLB_C128 ? 16 : 8;
  const char* et = std::getenv("QMLB_FSTREAM_TILE_BITS");
  const char* el = std::getenv("QMLB_FSTREAM_LOW_BITS");
  const int T = et ? std::atoi(et) : (p->dtype == QMLB_C128 ? 12 : 13);
  const int L = el ? std::atoi(el) : (p->dtype == QMLB_C128 ? 5 : 6);  // 512-byte runs
  if (N < T + 1 || N > 40 || T < 10 || T > 14 || L < 4 || L > T - 4) return QMLB_ERR_UNSUPPORTED;
  Builder B;
  B.p = p;
  if (!analyse(p, B.info)) return QMLB_ERR_UNSUPPORTED;
  for (const qmlb_op& o : p->ops)
    if (o.kind == QMLB_OP_MAT && o.k >= 3) return QMLB_ERR_UNSUPPORTED;
  B.N = N;
  B.bank_bits = p->dtype == QMLB_C128 ? 3 : 4;  // complex elements per 128-byte wavefront
  B.T = T;
  B.G = N - T;
  B.L = L;
  int row = 0;
  B.premat_off.assign(p->ops.size(), 0);
  B.matlist_index.assign(p->ops.size(), -1);
  p->stream_matlist.clear();
  for (size_t i = 0; i < p->ops.size(); ++i) {
    if (p->ops[i].kind == QMLB_OP_PERM) continue;
    B.premat_off[i] = row;
    B.matlist_index[i] = (int)p->stream_matlist.size();
    StreamMatOp mo{};
    mo.src = p->ops[i].src;
    mo.off = row;
    mo.swap2 = 0;
    p->stream_matlist.push_back(mo);
    row += B.info[i].entries;
  }
  p->stream_mat_row = row;
  // all matrices of an element stay in shared memory next to the tile
  const size_t tile_bytes = (size_t(1) << T) * cs;
  if (tile_bytes + (size_t)row * cs > 96 * 1024) return QMLB_ERR_UNSUPPORTED;
  B.mat_cap = std::max(row, 1);
  B.resident = true;
  B.pass_compact = true;
  const int rc = B.run_stream();
  if (rc != QMLB_OK) return rc;
  p->frame_steps = std::move(B.steps);
  p->frame_step_ops = std::move(B.step_ops);
  p->fstream_passes = std::move(B.passes);
  p->fstream_final_hpos = B.final_hpos;
  FrameProg& fp = p->frame;
  std::memset(&fp, 0, sizeof(fp));
  fp.n_steps = (int)p->frame_steps.size();
  fp.n_bits = N;
  fp.tile_bits = T;
  fp.outer_bits = N - T;
  fp.team_bits = 8;
  fp.teams = 1;
  fp.mat_cap = (std::max(B.pass_used_max, 1) + 15) & ~15;  // keeps the step records 128-byte aligned
  fp.mat_resident = 1;
  fp.premat_row = row;
  fp.density = p->density;
  fp.n_qubits = p->n_qubits;
  fp.n_obs = (int)p->obs.size();
  p->fstream_low_bits = L;
  p->frame_threads = 256;
  // [tile | matrices of the largest pass | two step records | relayout tables | mbarrier]:
  // 2^13 complex64 amplitudes + ~4 KiB leave room for three CTAs per SM
  p->frame_smem = tile_bytes + (size_t)fp.mat_cap * cs + 2 * sizeof(FrameStep) +
                  (256 + 64) * sizeof(uint32_t) + 64;
  return QMLB_OK;
}

// Fills p->frame_* ; returns QMLB_ERR_UNSUPPORTED (without touching the error string) when
// the program is outside the engine's envelope so that plan() can fall back.
int plan_frame(qmlb_program* p) {
  const int N = p->n_bits;
  const size_t cs = p->dtype == QMLB_C128 ? 16 : 8;
  // programs with a dense op on 3-4 bits run the HEAVY kernel variant (128 registers): its
  // relayout holds at most 32 elements per thread, i.e. tiles of 2^13 also in complex64
  bool heavy = false;
  for (const qmlb_op& o : p->ops)
    if (o.kind == QMLB_OP_MAT && o.k >= 3) heavy = true;
  const int maxT = (p->dtype == QMLB_C128 || heavy) ? 13 : 14;
  if (N < 6 || N > maxT + 3) return QMLB_ERR_UNSUPPORTED;
  Builder B;
  B.p = p;
  if (!analyse(p, B.info)) return QMLB_ERR_UNSUPPORTED;
  B.N = N;
  B.bank_bits = p->dtype == QMLB_C128 ? 3 : 4;  // complex elements per 128-byte wavefront
  B.G = std::max(0, N - maxT);
  B.T = N - B.G;

  // geometry
  int threads, team_bits, teams;
  if (B.T - FRAME_R >= 8) {
    // T >= 13: 256 threads with two items each at up to 255 registers (measured 4-5 % faster
    // than 512 threads x one item at 128 registers, which spills); QMLB_FRAME_THREADS=512
    // selects the latter
    const char* ev = std::getenv("QMLB_FRAME_THREADS");
    const bool wide = B.T >= 13 && ev && std::atoi(ev) == 512;
    threads = wide ? 512 : 256;
    team_bits = wide ? 9 : 8;
    teams = 1;
  } else {
    threads = 256;
    team_bits = B.T - FRAME_R;
    teams = threads >> team_bits;
  }
  // evaluated matrices of one element (premat row): every gate op, program order
  int row = 0, biggest = 1;
  for (size_t i = 0; i < p->ops.size(); ++i) {
    if (p->ops[i].kind == QMLB_OP_PERM) continue;
    row += B.info[i].entries;
    biggest = std::max(biggest, B.info[i].entries);
  }
  // Shared memory: [tiles | matrices | step records, tables].  Preferred: the element's
  // whole row of matrices stays resident (one copy per element, no per-step staging);
  // else a per-step staging area.
  const size_t budget = 216 * 1024;
  const size_t fixed = 2 * sizeof(FrameStep) + (256 + 64) * sizeof(uint32_t) + 64 * sizeof(double);
  const size_t tile_bytes = (size_t(1) << B.T) * cs;
  auto fits = [&](int tm, int cap) {
    return (size_t)tm * (tile_bytes + (size_t)cap * cs) + fixed <= budget;
  };
  int cap = std::max(row, 1);
  bool resident = true;
  // several tiles per CTA: matrices straight from global memory, all step records resident,
  // padded tile pitch (FrameProg::steps_resident); QMLB_FRAME_SMALL=0 -> round-1 geometry
  const char* es = std::getenv("QMLB_FRAME_SMALL");
  const bool small = teams > 1 && !(es && std::atoi(es) == 0);
  if (!small && !fits(teams, cap)) {
    int tm = teams;
    while (tm > 1 && (tm << team_bits) > 128 && !fits(tm, cap)) tm >>= 1;
    if (fits(tm, cap)) {
      teams = tm;
    } else {
      resident = false;
      cap = std::max(biggest, teams == 1 ? 1024 : 96);
      while (teams > 1 && !fits(teams, cap)) teams >>= 1;
      if (!fits(teams, cap)) {
        cap = (int)((budget - fixed - tile_bytes) / cs);
        if (cap < biggest) return QMLB_ERR_UNSUPPORTED;
      }
    }
  }
  if (teams > 1 || team_bits < 8) threads = teams << team_bits;
  if (threads < 32) return QMLB_ERR_UNSUPPORTED;
  B.mat_cap = cap;
  B.resident = resident;

  // evaluated matrices of one element: every gate op, program order
  B.premat_off.assign(p->ops.size(), 0);
  B.matlist_index.assign(p->ops.size(), -1);
  p->stream_matlist.clear();
  row = 0;
  for (size_t i = 0; i < p->ops.size(); ++i) {
    if (p->ops[i].kind == QMLB_OP_PERM) continue;
    B.premat_off[i] = row;
    B.matlist_index[i] = (int)p->stream_matlist.size();
    StreamMatOp mo{};
    mo.src = p->ops[i].src;
    mo.off = row;
    mo.swap2 = 0;
    p->stream_matlist.push_back(mo);
    row += B.info[i].entries;
  }
  p->stream_mat_row = row;

  const int rc = B.run();
  if (rc != QMLB_OK) return rc;
  p->frame_steps = std::move(B.steps);
  p->frame_step_ops = std::move(B.step_ops);
  FrameProg& fp = p->frame;
  std::memset(&fp, 0, sizeof(fp));
  fp.n_steps = (int)p->frame_steps.size();
  fp.n_bits = N;
  fp.tile_bits = B.T;
  fp.outer_bits = B.G;
  fp.team_bits = team_bits;
  fp.teams = teams;
  fp.mat_cap = cap;
  fp.mat_resident = small ? 2 : (resident ? 1 : 0);
  fp.premat_row = row;
  fp.density = p->density;
  fp.n_qubits = p->n_qubits;
  fp.n_obs = (int)p->obs.size();
  p->frame_threads = threads;
  p->frame_heavy = false;
  for (const qmlb_op& o : p->ops)
    if (o.kind == QMLB_OP_MAT && o.k >= 3) p->frame_heavy = true;
  // [tiles | matrices | step records | relayout tables | reduction scratch]
  if (small) {
    fp.tile_pitch = (1 << B.T) + 1;
    fp.steps_resident = fp.n_steps <= 64 ? 1 : 0;
    const size_t tiles = (((size_t)teams * fp.tile_pitch + 63) & ~size_t(63)) * cs;
    p->frame_smem = tiles + (size_t)(fp.steps_resident ? fp.n_steps : 2) * sizeof(FrameStep) +
                    (fixed - 2 * sizeof(FrameStep));
  } else {
    p->frame_smem = (((size_t)teams * (size_t(1) << B.T) + 63) & ~size_t(63)) * cs +
                    (size_t)teams * cap * cs + fixed;
  }
  return QMLB_OK;
}


// ---------------------------------------------------------------------------------------
// Pauli-basis (transfer-matrix) form of a density program.  rho = 2^-n sum_P r_P P with REAL
// coefficients r_P = Tr(P rho): half the memory of the complex 4^n vector (config 4: 512 KiB
// instead of 1 MiB -> a cluster of 4 CTAs instead of 8) and a quarter of the arithmetic (a
// 1-qubit superoperator is a real 4x4 on four real numbers instead of a complex 4x4 on four
// complex ones).  Qubit w keeps its two index bits: the ket bit now holds x_w, the bra bit
// z_w, Pauli = X^x Z^z up to phase (I = 00, Z = 01, X = 10, Y = 11), so the fused (ket, bra)
// superoperator ops become 4x4 transfer matrices R = T S T^-1 on the SAME bits (evaluated by
// k_stream_mats, swap2 = 2).  A CX is a Clifford: it permutes Paulis by the GF(2)-linear map
// x_t ^= x_c, z_c ^= z_t - folded into the frame like every linear permutation - times the
// sign (-1)^(x_c z_t (x_t ^ z_c ^ 1)) (Aaronson-Gottesman), applied as a +-1 diagonal over
// four parity rows.  Eligible: density programs made of 1-qubit (ket, bra) ops and CX / SWAP
// pairs with <Z-string> or probability output; everything else keeps the complex form.
static const int kCxFirst[4] = {0, 1, 3, 2}, kCxSecond[4] = {0, 3, 2, 1}, kSwap[4] = {0, 2, 1, 3};

static int perm_kind(const qmlb_program* p, const qmlb_op& o) {
  if (o.kind != QMLB_OP_PERM || o.k != 2) return -1;
  int t[4];
  for (int v = 0; v < 4; ++v) t[v] = (int)p->consts[o.aux + v];
  if (std::equal(t, t + 4, kCxFirst)) return 0;
  if (std::equal(t, t + 4, kCxSecond)) return 1;
  if (std::equal(t, t + 4, kSwap)) return 2;
  return -1;
}

// host value of a superoperator source made of constants only (else false)
static bool const_super(const qmlb_program* p, int sid, double out[32]) {
  const qmlb_source& s = p->sources[sid];
  auto load = [&](int off, int n, double* dst) {
    for (int i = 0; i < 2 * n; ++i) dst[i] = p->consts[2 * (size_t)off + i];
  };
  if (s.kind == QMLB_SRC_CONST && s.k == 2 && !(s.flags & QMLB_FLAG_DIAGVEC)) {
    load(s.a0, 16, out);
    return true;
  }
  if (s.kind != QMLB_SRC_SUPER) return false;
  double acc[32] = {0};
  for (int i = 0; i < 4; ++i) acc[2 * (i * 5)] = 1.0;
  for (int t = 0; t < s.a1; ++t) {
    const qmlb_source& it = p->sources[p->items[s.a0 + t]];
    if (!(it.kind == QMLB_SRC_CONST && it.k == 2 && !(it.flags & QMLB_FLAG_DIAGVEC))) return false;
    double f[32], r[32];
    load(it.a0, 16, f);
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        double re = 0, im = 0;
        for (int l = 0; l < 4; ++l) {
          const double ar = f[2 * (i * 4 + l)], ai = f[2 * (i * 4 + l) + 1];
          const double br = acc[2 * (l * 4 + j)], bi = acc[2 * (l * 4 + j) + 1];
          re += ar * br - ai * bi;
          im += ar * bi + ai * br;
        }
        r[2 * (i * 4 + j)] = re;
        r[2 * (i * 4 + j) + 1] = im;
      }
    std::memcpy(acc, r, sizeof(acc));
  }
  std::memcpy(out, acc, sizeof(acc));
  return true;
}

// R = T S T^-1 for S in (ket, bra) order [rho00, rho01, rho10, rho11] and Pauli order
// (I, Z, X, Y); is it diagonal?
static bool ptm_is_diagonal(const double S[32]) {
  // T rows: I = (1,0,0,1), Z = (1,0,0,-1), X = (0,1,1,0), Y = (0,i,-i,0); T^-1 = T^dagger / 2
  const double Tr[4][4] = {{1, 0, 0, 1}, {1, 0, 0, -1}, {0, 1, 1, 0}, {0, 0, 0, 0}};
  const double Ti[4][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 1, -1, 0}};
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      double re = 0, im = 0;
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
          // T[a][i] * S[i][j] * conj(T[b][j]) / 2
          const double tr = Tr[a][i], ti = Ti[a][i], sr = S[2 * (i * 4 + j)],
                       si = S[2 * (i * 4 + j) + 1], ur = Tr[b][j], ui = -Ti[b][j];
          const double xr = tr * sr - ti * si, xi = tr * si + ti * sr;
          re += 0.5 * (xr * ur - xi * ui);
          im += 0.5 * (xr * ui + xi * ur);
        }
      if (a != b && (std::abs(re) > 1e-15 || std::abs(im) > 1e-15)) return false;
    }
  return true;
}

int plan_frame_ptm(qmlb_program* p) {
  if (!p->density) return QMLB_ERR_UNSUPPORTED;
  const int n = p->n_qubits, N = p->n_bits;
  if (!(p->out_type == QMLB_OUT_PROBS || p->out_type == QMLB_OUT_DENSITY ||
        (p->out_type == QMLB_OUT_EXPVAL &&
         std::all_of(p->obs.begin(), p->obs.end(),
                     [](const qmlb_obs& o) { return o.kind == QMLB_OBS_ZSTRING; }))))
    return QMLB_ERR_UNSUPPORTED;
  if (p->out_type == QMLB_OUT_PROBS && n > 10) return QMLB_ERR_UNSUPPORTED;
  // ---- rewrite the op list ----------------------------------------------------------
  qmlb_program q = *p;  // shares sources / items / angles / terms; ops and consts change
  q.blob = nullptr;
  q.ops.clear();
  std::vector<char> diag_flag;
  std::vector<int> origin;  // rewritten op -> index of the original op (describe / tests)
  for (size_t i = 0; i < p->ops.size(); ++i) {
    const qmlb_op& o = p->ops[i];
    if (o.kind == QMLB_OP_MAT && o.k == 2 && o.bits[0] - o.bits[1] == n && o.bits[1] < n) {
      q.ops.push_back(o);
      origin.push_back((int)i);
      double S[32];
      diag_flag.push_back(const_super(p, o.src, S) && ptm_is_diagonal(S) ? 1 : 0);
      continue;
    }
    const int kind = perm_kind(p, o);
    if (kind < 0 || i + 1 >= p->ops.size()) return QMLB_ERR_UNSUPPORTED;
    const qmlb_op& o2 = p->ops[i + 1];
    if (perm_kind(p, o2) != kind || o.bits[0] < n || o.bits[1] < n ||
        o2.bits[0] != o.bits[0] - n || o2.bits[1] != o.bits[1] - n)
      return QMLB_ERR_UNSUPPORTED;
    ++i;
    // bit roles: ket bit of wire w = x_w, bra bit = z_w
    const int a = kind == 1 ? 1 : 0;  // index (in o.bits) of the control / first wire
    const int xc = o.bits[a], zc = o2.bits[a], xt = o.bits[1 - a], zt = o2.bits[1 - a];
    qmlb_op perm{};
    perm.kind = QMLB_OP_PERM;
    perm.k = 4;
    perm.src = -1;
    perm.aux = (int32_t)q.consts.size();
    perm.bits[0] = xc;
    perm.bits[1] = zc;
    perm.bits[2] = xt;
    perm.bits[3] = zt;
    int sign_mask = 0;
    for (int v = 0; v < 16; ++v) {  // v = x_c z_c x_t z_t
      const int vxc = v >> 3 & 1, vzc = v >> 2 & 1, vxt = v >> 1 & 1, vzt = v & 1;
      int img;
      if (kind == 2) {
        img = (vxt << 3) | (vzt << 2) | (vxc << 1) | vzc;
      } else {
        img = (vxc << 3) | ((vzc ^ vzt) << 2) | ((vxt ^ vxc) << 1) | vzt;
        if (vxc & vzt & (vxt ^ vzc ^ 1)) sign_mask |= 1 << v;
      }
      q.consts.push_back((double)img);
    }
    q.ops.push_back(perm);
    diag_flag.push_back(0);
    origin.push_back((int)i - 1);
    if (sign_mask) {
      qmlb_op sg{};
      sg.kind = QMLB_OP_SIGN;
      sg.k = 4;
      sg.src = -1;
      sg.aux = sign_mask;
      sg.bits[0] = xc;
      sg.bits[1] = zc;
      sg.bits[2] = xt;
      sg.bits[3] = zt;
      q.ops.push_back(sg);
      diag_flag.push_back(0);
      origin.push_back((int)i - 1);
    }
  }
  // ---- geometry: real elements -----------------------------------------------------
  const size_t rs = p->dtype == QMLB_C128 ? 8 : 4, cs = 2 * rs;
  const int maxT = p->dtype == QMLB_C128 ? 14 : 15;
  if (N < 6 || N > maxT + 3) return QMLB_ERR_UNSUPPORTED;
  Builder B;
  B.p = &q;
  B.ptm = true;
  B.ptm_diag = diag_flag;
  if (!analyse(&q, B.info)) return QMLB_ERR_UNSUPPORTED;
  B.N = N;
  B.G = std::max(0, N - maxT);
  B.T = N - B.G;
  int threads, team_bits, teams;
  if (B.T - FRAME_R >= 8) {
    // 512 threads with two (complex128) or four (complex64) items each: 128 registers per
    // thread hold an item, its addresses and a transfer matrix without spills
    const char* tb = std::getenv("QMLB_PTM_TEAM_BITS");
    const int cap = tb ? std::max(8, std::min(10, std::atoi(tb))) : 9;
    team_bits = std::min(cap, B.T - FRAME_R);
    // a relayout holds 2^(T - team_bits) elements per thread in registers: at most 32 doubles
    if (p->dtype == QMLB_C128) team_bits = std::max(team_bits, B.T - 5);
    threads = 1 << team_bits;
    teams = 1;
  } else {
    threads = 256;
    team_bits = B.T - FRAME_R;
    teams = threads >> team_bits;
  }
  B.swz_bits = p->dtype == QMLB_C128 ? 4 : 5;  // 128-byte wavefront / element size
  B.bank_bits = 0;  // the swizzled tile keeps the plain relayout form (measured faster)
  B.team_bits = team_bits;
  int row = 0;
  B.premat_off.assign(q.ops.size(), 0);
  B.matlist_index.assign(q.ops.size(), -1);
  q.stream_matlist.clear();
  for (size_t i = 0; i < q.ops.size(); ++i) {
    if (q.ops[i].kind != QMLB_OP_MAT) continue;
    B.premat_off[i] = row;
    B.matlist_index[i] = (int)q.stream_matlist.size();
    StreamMatOp mo{};
    mo.src = q.ops[i].src;
    mo.off = row;
    mo.swap2 = 2;
    q.stream_matlist.push_back(mo);
    row += B.info[i].entries;  // 16 complex slots; the 16 reals of R use the first half
  }
  q.stream_mat_row = row;
  const size_t budget = 216 * 1024;
  const size_t fixed = 2 * sizeof(FrameStep) + (256 + 128) * sizeof(uint32_t) + 2048 * sizeof(double);
  const size_t tile_bytes = (size_t(1) << B.T) * rs;
  while (teams > 1 && (size_t)teams * (tile_bytes + (size_t)row * cs) + fixed > budget) teams >>= 1;
  if ((size_t)teams * (tile_bytes + (size_t)row * cs) + fixed > budget) return QMLB_ERR_UNSUPPORTED;
  if (teams > 1 || team_bits < 8) threads = teams << team_bits;
  if (threads < 32) return QMLB_ERR_UNSUPPORTED;
  B.mat_cap = std::max(row, 1);
  B.resident = true;
  const int rc = B.run();
  if (rc != QMLB_OK) return rc;
  // the physical positions that hold x bits at the start (initial frame = permutation)
  p->stream_matlist = q.stream_matlist;
  p->stream_mat_row = row;
  p->frame_steps = std::move(B.steps);
  p->frame_step_ops = std::move(B.step_ops);
  for (auto& v : p->frame_step_ops)
    for (int& idx : v) idx = origin[idx];
  FrameProg& fp = p->frame;
  std::memset(&fp, 0, sizeof(fp));
  fp.n_steps = (int)p->frame_steps.size();
  fp.n_bits = N;
  fp.tile_bits = B.T;
  fp.outer_bits = B.G;
  fp.team_bits = team_bits;
  fp.teams = teams;
  fp.mat_cap = B.mat_cap;
  fp.mat_resident = 1;
  fp.premat_row = row;
  fp.density = 1;
  fp.n_qubits = n;
  fp.n_obs = (int)p->obs.size();
  fp.ptm = 1;
  p->frame_ptm_xmask = B.init_xmask;
  p->frame_threads = threads;
  p->frame_heavy = false;
  p->frame_smem = (size_t)teams * (tile_bytes + (size_t)row * cs) + fixed;
  return QMLB_OK;
}

std::string describe_frame(const qmlb_program* p) {
  std::string s;
  const FrameProg& fp = p->frame;
  if (!p->fstream_passes.empty()) {
    s += "fstream low_bits " + std::to_string(p->fstream_low_bits) + " final_hpos";
    for (int h : p->fstream_final_hpos) s += " " + std::to_string(h);
    s += "\n";
    for (const FramePassHost& ps : p->fstream_passes) {
      s += "pass init " + std::to_string(ps.init ? 1 : 0) + " first " +
           std::to_string(ps.first_step) + " steps " + std::to_string(ps.n_steps) + " tp";
      for (int b : ps.tp) s += " " + std::to_string(b);
      s += " opos";
      for (int b : ps.opos) s += " " + std::to_string(b);
      s += "\n";
    }
  }
  s += "frame tile_bits " + std::to_string(fp.tile_bits) + " outer_bits " +
       std::to_string(fp.outer_bits) + " team_bits " + std::to_string(fp.team_bits) +
       " teams " + std::to_string(fp.teams) + " threads " + std::to_string(p->frame_threads) +
       " mat_cap " + std::to_string(fp.mat_cap) + " resident " + std::to_string(fp.mat_resident) +
       " premat_row " + std::to_string(fp.premat_row) + " ptm " + std::to_string(fp.ptm) +
       " xmask " + std::to_string((unsigned long long)p->frame_ptm_xmask) +
       " smem " + std::to_string(p->frame_smem) + "\n";
  for (size_t i = 0; i < p->frame_steps.size(); ++i) {
    const FrameStep& st = p->frame_steps[i];
    if (st.kind == QMLB_FSTEP_RELAYOUT) {
      s += st.mat_entries ? "relayout local" : "relayout";
      for (int b = 0; b < fp.n_bits; ++b) s += " " + std::to_string((unsigned long long)st.qcol[b]);
      s += " htab";
      const uint32_t* htab = reinterpret_cast<const uint32_t*>(st.ops);
      for (int l = 0; l < 32; ++l) s += " " + std::to_string(htab[l]);
      s += "\n";
      continue;
    }
    s += "subpass fast " + std::to_string(st.fast);
    {
      const FrameSubX* sx = reinterpret_cast<const FrameSubX*>(st.qcol);
      s += " lanes_ok " + std::to_string(sx->lanes_ok) + " ipos";
      for (int b = 0; b < fp.tile_bits - FRAME_R && b < 16; ++b) s += " " + std::to_string(sx->ipos[b]);
      s += " kd";
      for (int b = 0; b < 4; ++b) s += " " + std::to_string(sx->kd[b]);
      s += " sg";
      for (int b = 0; b < 4; ++b) s += " " + std::to_string(sx->sg[b]);
    }
    s += " pivots";
    for (int j = 0; j < FRAME_R; ++j) s += " " + std::to_string(st.pivots[j]);
    s += " eoff";
    for (int v = 0; v < FRAME_D; ++v) s += " " + std::to_string(st.eoff[v]);
    s += " par";
    for (int j = 0; j < st.n_par; ++j)
      s += " " + std::to_string(st.par[j].rloc) + ":" + std::to_string(st.par[j].rout) + ":" +
           std::to_string(st.par[j].smask);
    s += " ops";
    size_t t = 0;
    for (int o = 0; o < st.n_ops; ++o) {
      const FrameOp& fo = st.ops[o];
      s += " " + std::to_string(p->frame_step_ops[i][t++]) + ":" + std::to_string(fo.code) + ":" +
           std::to_string(fo.k) + ":" + std::to_string(fo.j0) + ":" + std::to_string(fo.j1) + ":" +
           std::to_string(fo.has_c) + ":" + std::to_string(fo.flags) + ":" +
           std::to_string(fo.premat_off) + ":" + std::to_string(fo.smem_off) + ":" +
           std::to_string(fo.shape);
      if (fo.code == QMLB_FOP_DIAG) {
        const uint8_t* idx = reinterpret_cast<const uint8_t*>(&st.ops[o + 1]);
        s += ":";
        for (int a = 0; a < fo.k; ++a) s += (a ? "," : "") + std::to_string(idx[a]);
        ++o;
      } else if (fo.code == QMLB_FOP_SIGN) {
        const uint8_t idx[4] = {fo.j0, fo.j1, fo.shape, fo.has_c};
        s += ":";
        for (int a = 0; a < 4; ++a) s += (a ? "," : "") + std::to_string(idx[a]);
        const uint16_t* tab = reinterpret_cast<const uint16_t*>(&st.ops[o + 2]);
        s += ":";
        for (int a = 0; a < 16; ++a) s += (a ? "," : "") + std::to_string(tab[a]);
        o += 3;
      }
    }
    s += "\n";
  }
  return s;
}

}  // namespace qmlb
