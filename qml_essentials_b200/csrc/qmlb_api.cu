// C ABI of the B200 backend: program validation / scheduling / upload and the
// launch sequences.  See include/qmlb200.h.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "qmlb_internal.h"
#include "qmlb_analysis.cuh"
#include "qmlb_measure.cuh"
#include "qmlb_reg.cuh"

using namespace qmlb;

namespace qmlb {
std::atomic<unsigned long long> g_launches{0};
}

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CUDA_TRY(expr)                                                              \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess)                                                         \
      return fail(QMLB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)

constexpr int THREADS = qmlb::TILE_THREADS;
constexpr int SM_COUNT_FALLBACK = 148;

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

}  // namespace

namespace {

size_t cs_of(int dtype) { return dtype == QMLB_C128 ? 16 : 8; }
size_t rs_of(int dtype) { return dtype == QMLB_C128 ? 8 : 4; }

int validate(const qmlb_program_desc* d, qmlb_program* p) {
  if (d->n_qubits < 1 || d->n_bits != (d->density ? 2 * d->n_qubits : d->n_qubits))
    return fail(QMLB_ERR_INVALID, "n_bits does not match n_qubits/density");
  if (d->n_bits > 40) return fail(QMLB_ERR_UNSUPPORTED, "more than 40 state bits");
  if (d->dtype != QMLB_C64 && d->dtype != QMLB_C128)
    return fail(QMLB_ERR_INVALID, "dtype must be QMLB_C64 or QMLB_C128");
  if (d->out_type < 0 || d->out_type > 3) return fail(QMLB_ERR_INVALID, "bad out_type");
  if (d->out_type == QMLB_OUT_STATE && d->density)
    return fail(QMLB_ERR_INVALID,
                "Measurement type 'state' is not defined for mixed (noisy) circuits");
  for (int i = 0; i < d->n_sources; ++i) {
    const qmlb_source& s = d->sources[i];
    if (s.kind < 0 || s.kind > QMLB_SRC_PRE || s.k < 1 || s.k > QMLB_MAX_OP_BITS)
      return fail(QMLB_ERR_INVALID, "bad source record");
    if ((s.kind == QMLB_SRC_TRIG || s.kind == QMLB_SRC_DIAGPH) &&
        (s.angle < 0 || s.angle >= d->n_angles))
      return fail(QMLB_ERR_INVALID, "source references a missing angle");
    if (s.kind == QMLB_SRC_CHAIN || s.kind == QMLB_SRC_SUPER) {
      if (s.a0 < 0 || s.a1 < 1 || s.a0 + s.a1 > d->n_items)
        return fail(QMLB_ERR_INVALID, "chain items out of range");
      for (int t = 0; t < s.a1; ++t) {
        int id = d->items[s.a0 + t];
        if (id < 0 || id >= d->n_sources) return fail(QMLB_ERR_INVALID, "bad chain item");
        const qmlb_source& it = d->sources[id];
        bool elem1 = it.k == 1 && (it.kind == QMLB_SRC_CONST || it.kind == QMLB_SRC_TRIG ||
                                   it.kind == QMLB_SRC_TABLE || it.kind == QMLB_SRC_PRE);
        if (s.kind == QMLB_SRC_CHAIN && !elem1)
          return fail(QMLB_ERR_INVALID, "chain item must be an elementary 2x2 source");
        if (s.kind == QMLB_SRC_SUPER && !elem1 &&
            !(it.k == 1 && it.kind == QMLB_SRC_CHAIN) &&
            !(it.k == 2 && it.kind == QMLB_SRC_CONST))
          return fail(QMLB_ERR_INVALID, "superchain item must be a 2x2 source or const 4x4");
      }
    }
    if (s.kind == QMLB_SRC_TABLE) {
      if (s.a0 < 0 || s.a0 >= QMLB_MAX_ARGS) return fail(QMLB_ERR_INVALID, "bad table arg");
      p->max_arg = std::max(p->max_arg, s.a0);
    }
    if (s.kind == QMLB_SRC_PRE) {
      if (s.k != 1 || s.a2 < 0 || s.a2 >= d->n_pre || d->pre[s.a2].arg != s.a1 ||
          d->pre[s.a2].local != s.a0)
        return fail(QMLB_ERR_INVALID, "hoisted source does not match its pre entry");
    }
  }
  {
    int count[QMLB_MAX_ARGS] = {0};
    for (int i = 0; i < d->n_pre; ++i) {
      const qmlb_pre& e = d->pre[i];
      if (e.arg < 0 || e.arg >= QMLB_MAX_ARGS || e.src < 0 || e.src >= d->n_sources)
        return fail(QMLB_ERR_INVALID, "bad pre entry");
      if (e.local != count[e.arg]++)
        return fail(QMLB_ERR_INVALID, "pre entries of a slot must be numbered in order");
      // the hoisted source: elementary 2x2 or a chain of elementary 2x2 (no nesting)
      const qmlb_source& h = d->sources[e.src];
      auto elem = [&](const qmlb_source& it) {
        return it.k == 1 && (it.kind == QMLB_SRC_CONST || it.kind == QMLB_SRC_TRIG);
      };
      bool ok = elem(h);
      if (h.kind == QMLB_SRC_CHAIN && h.k == 1) {
        ok = h.a0 >= 0 && h.a1 >= 1 && h.a0 + h.a1 <= d->n_items;
        for (int t = 0; ok && t < h.a1; ++t) {
          int id = d->items[h.a0 + t];
          ok = id >= 0 && id < d->n_sources && elem(d->sources[id]);
        }
      }
      if (!ok) return fail(QMLB_ERR_INVALID, "pre entry must hoist elementary 2x2 sources");
      // ... whose angles read slot `arg` only
      auto slot_ok = [&](const qmlb_source& it) {
        if (it.kind != QMLB_SRC_TRIG) return true;
        if (it.angle < 0 || it.angle >= d->n_angles) return false;
        const qmlb_angle& a = d->angles[it.angle];
        for (int t = 0; t < a.n; ++t)
          if (a.first + t >= d->n_terms || d->terms[a.first + t].arg != e.arg) return false;
        return true;
      };
      ok = slot_ok(h);
      if (h.kind == QMLB_SRC_CHAIN)
        for (int t = 0; ok && t < h.a1; ++t) ok = slot_ok(d->sources[d->items[h.a0 + t]]);
      if (!ok) return fail(QMLB_ERR_INVALID, "pre entry reads another argument slot");
    }
  }
  for (int i = 0; i < d->n_terms; ++i) {
    if (d->terms[i].arg < 0 || d->terms[i].arg >= QMLB_MAX_ARGS)
      return fail(QMLB_ERR_INVALID, "term references argument slot outside 0..7");
    p->max_arg = std::max(p->max_arg, d->terms[i].arg);
  }
  for (int i = 0; i < d->n_angles; ++i)
    if (d->angles[i].first < 0 || d->angles[i].first + d->angles[i].n > d->n_terms)
      return fail(QMLB_ERR_INVALID, "angle terms out of range");
  for (int i = 0; i < d->n_ops; ++i) {
    const qmlb_op& o = d->ops[i];
    if (o.kind < 0 || o.kind > QMLB_OP_DIAG || o.k < 1 || o.k > QMLB_MAX_OP_BITS)
      return fail(QMLB_ERR_INVALID, "bad op record");
    uint64_t seen = 0;
    for (int j = 0; j < o.k; ++j) {
      if (o.bits[j] < 0 || o.bits[j] >= d->n_bits)
        return fail(QMLB_ERR_INVALID, "op bit outside the state");
      if (seen >> o.bits[j] & 1) return fail(QMLB_ERR_INVALID, "op repeats a bit");
      seen |= 1ull << o.bits[j];
    }
    if (o.kind != QMLB_OP_PERM && (o.src < 0 || o.src >= d->n_sources))
      return fail(QMLB_ERR_INVALID, "op references a missing source");
    if (o.kind == QMLB_OP_PERM && (o.aux < 0 || o.aux + (1 << o.k) > d->n_consts))
      return fail(QMLB_ERR_INVALID, "permutation table out of range");
    if ((o.kind == QMLB_OP_MAT || o.kind == QMLB_OP_PERM) && o.k > 4)
      return fail(QMLB_ERR_UNSUPPORTED, "dense / permutation ops on more than 4 bits");
    if (o.kind == QMLB_OP_CTRL1 && o.k != 2) return fail(QMLB_ERR_INVALID, "CTRL1 needs 2 bits");
  }
  for (int i = 0; i < d->n_obs; ++i) {
    const qmlb_obs& o = d->obs[i];
    if (o.kind < 0 || o.kind > QMLB_OBS_DENSE || o.k < 1 || o.k > QMLB_MAX_OP_BITS)
      return fail(QMLB_ERR_INVALID, "bad observable record");
    if (o.kind == QMLB_OBS_DENSE && o.k > 4)
      return fail(QMLB_ERR_UNSUPPORTED, "dense observables on more than 4 qubits");
    for (int j = 0; j < o.k; ++j)
      if (o.bits[j] < 0 || o.bits[j] >= d->n_qubits)
        return fail(QMLB_ERR_INVALID, "observable bit outside the register");
  }
  if (d->out_type == QMLB_OUT_EXPVAL && d->n_obs < 1)
    return fail(QMLB_ERR_INVALID, "expval needs at least one observable");
  return QMLB_OK;
}

int op_entries(const qmlb_program* p, const qmlb_op& o) {
  if (o.kind == QMLB_OP_PERM) return 0;
  const qmlb_source& s = p->sources[o.src];
  return source_entries(s.kind, s.k, s.flags);
}

bool all_zstring(const qmlb_program* p) {
  for (const auto& o : p->obs)
    if (o.kind != QMLB_OBS_ZSTRING) return false;
  return true;
}

// split a pass' ops into windows whose matrices fit `matw` entries
void make_windows(const qmlb_program* p, QmlbPassHost& ps, int matw) {
  ps.matoff.assign(ps.ops.size(), 0);
  ps.windows.clear();
  int first = 0, used = 0;
  for (size_t i = 0; i < ps.ops.size(); ++i) {
    int e = op_entries(p, ps.ops[i]);
    if (used + e > matw && (int)i > first) {
      ps.windows.push_back(make_int2(first, (int)i - first));
      first = (int)i;
      used = 0;
    }
    ps.matoff[i] = used;
    used += e;
  }
  if ((int)ps.ops.size() > first || ps.ops.empty())
    ps.windows.push_back(make_int2(first, (int)ps.ops.size() - first));
  ps.matw = matw;
}

// ---- strategy 2: pass construction for the streaming kernel -----------------------
// Every pass owns a group of R state bits.  Greedy list scheduling: scan the
// not-yet-done ops in program order; an op joins the pass if none of its bits is
// blocked by an earlier op that had to be left out and the union of group bits
// stays within R.  Diagonal ops act on global indices and need no group bits.
// Ops on 3 or 4 bits pin their bits to register positions 0..k-1.
int schedule_stream_with(qmlb_program* p, int R, bool first_fit);

// Two list schedulers (first-fit in program order; look-ahead group choice) - neither
// dominates (nearest-neighbour statevector circuits favour look-ahead, the (ket, bra)
// structure of density programs first-fit), so both run and the shorter schedule wins.
int schedule_stream(qmlb_program* p, int R) {
  const int forced = env_int("QMLB_SCHED_GREEDY", -1);
  if (forced >= 0) return schedule_stream_with(p, R, forced != 0);
  int rc = schedule_stream_with(p, R, true);
  if (rc != QMLB_OK) return rc;
  std::vector<QmlbStreamPassHost> a = std::move(p->stream_passes);
  p->stream_passes.clear();
  rc = schedule_stream_with(p, R, false);
  if (rc != QMLB_OK || a.size() <= p->stream_passes.size()) p->stream_passes = std::move(a);
  return QMLB_OK;
}

int schedule_stream_with(qmlb_program* p, int R, bool first_fit) {
  const int N = p->n_bits;
  const int max_entries = 2048;  // matrix buffer entries per pass (<= 32 KB)
  std::vector<char> done(p->ops.size(), 0);
  size_t remaining = p->ops.size();
  bool first_pass = true;
  const bool pair_rule = p->dtype != QMLB_C128 && N >= R + 2 && env_int("QMLB_PAIR_RULE", 0);
  std::vector<uint64_t> opbits(p->ops.size(), 0);
  for (size_t i = 0; i < p->ops.size(); ++i)
    for (int j = 0; j < p->ops[i].k; ++j) opbits[i] |= 1ull << p->ops[i].bits[j];
  const int look_ahead = env_int("QMLB_SCHED_WINDOW", 600);
  const int max_seeds = env_int("QMLB_SCHED_SEEDS", 64);

  // ops (within the look-ahead window) that could run in a pass owning exactly group G
  auto closure_count = [&](uint64_t G) {
    uint64_t blk = 0;
    int count = 0, seen = 0;
    for (size_t i = 0; i < p->ops.size() && seen < look_ahead; ++i) {
      if (done[i]) continue;
      ++seen;
      const uint64_t b = opbits[i];
      if (b & blk) {
        blk |= b;
        continue;
      }
      if (p->ops[i].kind == QMLB_OP_DIAG || (b & ~G) == 0) {
        count += p->ops[i].kind == QMLB_OP_PERM ? 2 : 3;  // arithmetic ops weigh a bit more
      } else {
        blk |= b;
      }
    }
    return count;
  };
  auto normalise = [&](uint64_t G) {
    if (pair_rule && (G & 2ull) && !(G & 1ull)) G |= 1ull;
    return G;
  };
  // Group choice: from every ready op, grow the group one bit at a time by the bit that
  // unlocks the most work; keep the best.  (First-fit in program order - the previous
  // scheduler - strands neighbouring qubits in different passes.)
  auto choose_group = [&]() -> uint64_t {
    uint64_t cand_bits = 0, blk = 0, best = 0;
    int best_score = -1, seeds = 0, seen = 0;
    std::vector<size_t> ready;
    for (size_t i = 0; i < p->ops.size() && seen < look_ahead; ++i) {
      if (done[i]) continue;
      ++seen;
      cand_bits |= opbits[i];
      if (!(opbits[i] & blk) && p->ops[i].kind != QMLB_OP_DIAG) ready.push_back(i);
      blk |= opbits[i];
    }
    for (size_t i : ready) {
      if (seeds++ >= max_seeds) break;
      uint64_t G = normalise(opbits[i]);
      if (__builtin_popcountll(G) > R) continue;
      while (__builtin_popcountll(G) < R) {
        uint64_t pick = 0;
        int pick_score = -1;
        for (int b = 0; b < N; ++b) {
          if (!(cand_bits >> b & 1) || (G >> b & 1)) continue;
          const uint64_t G2 = normalise(G | (1ull << b));
          if (__builtin_popcountll(G2) > R) continue;
          const int sc = closure_count(G2);
          if (sc > pick_score) {
            pick_score = sc;
            pick = G2;
          }
        }
        if (!pick) break;
        G = pick;
      }
      const int sc = closure_count(G);
      if (sc > best_score) {
        best_score = sc;
        best = G;
      }
    }
    return best;
  };

  while (remaining > 0 || first_pass) {
    uint64_t S = first_fit ? 0 : choose_group(), blocked = 0;
    const bool fixed_group = __builtin_popcountll(S) == R;
    std::vector<int> canon;  // ordered: register position j -> state bit (k >= 3 op)
    std::vector<size_t> picked;
    int entries = 0;
    for (size_t i = 0; i < p->ops.size(); ++i) {
      if (done[i]) continue;
      const qmlb_op& o = p->ops[i];
      uint64_t bits = 0;
      for (int j = 0; j < o.k; ++j) bits |= 1ull << o.bits[j];
      // complex64, optional (QMLB_PAIR_RULE=1): a group that holds state bit 1 but not bit 0
      // touches every other 8 bytes of each line; bringing bit 0 along makes the thread move
      // whole 16-byte pairs.  Measured (profiles/r1_final_probe*.jsonl): the extra group bit
      // costs more passes than the wider accesses save (config 4 complex64: 90 vs 60
      // passes, 38 k vs 49 k evals/s), so the rule is off by default.
      uint64_t grp = bits;
      if (o.kind != QMLB_OP_DIAG && pair_rule && (bits & 2ull) &&
          __builtin_popcountll(bits | 1ull) <= R)
        grp |= 1ull;
      (void)fixed_group;
      const int e = op_entries(p, o);
      if ((bits & blocked) || entries + e > max_entries ||
          (int)picked.size() >= STREAM_MAX_OPS) {
        blocked |= bits;
        continue;
      }
      if (o.kind == QMLB_OP_DIAG) {
        picked.push_back(i);
        entries += e;
        continue;
      }
      if (o.k > R) return fail(QMLB_ERR_UNSUPPORTED, "operation wider than the register group");
      const uint64_t U = S | grp;
      bool ok = __builtin_popcountll(U) <= R;
      if (ok && o.k >= 3) {
        std::vector<int> want(o.k);
        for (int j = 0; j < o.k; ++j) want[o.k - 1 - j] = o.bits[j];
        if (canon.empty()) {
          // earlier ops of this pass may sit anywhere: positions are assigned at the end
          canon = want;
        } else {
          ok = want.size() <= canon.size() &&
               std::equal(want.begin(), want.end(), canon.begin());
        }
      }
      if (!ok) {
        blocked |= bits;
        continue;
      }
      S = U;
      picked.push_back(i);
      entries += e;
    }
    if (picked.empty() && remaining > 0)
      return fail(QMLB_ERR_UNSUPPORTED, "an operation does not fit a streaming pass");

    QmlbStreamPassHost ps;
    // register positions: pinned bits first, then the other group bits ascending, then pads
    std::vector<int> gb(canon);
    auto in_gb = [&](int g) { return std::find(gb.begin(), gb.end(), g) != gb.end(); };
    // ascending: state bits 0 (and 1) land on register bits 0 (and 1) unless an op on
    // 3-4 bits pinned other bits there -> the kernel moves 16-byte pairs
    for (int g = 0; g < N; ++g)
      if ((S >> g & 1) && !in_gb(g)) gb.push_back(g);
    if (p->dtype != QMLB_C128 && in_gb(1) && !in_gb(0) && (int)gb.size() < R) gb.push_back(0);
    for (int g = N - 1; g >= 0 && (int)gb.size() < R; --g)
      if (!in_gb(g)) gb.push_back(g);
    std::vector<int> pos(N, -1);
    for (int j = 0; j < R; ++j) {
      ps.gb[j] = gb[j];
      pos[gb[j]] = j;
    }
    std::vector<int> sorted(gb.begin(), gb.begin() + R);
    std::sort(sorted.begin(), sorted.end());
    for (int j = 0; j < R; ++j) ps.sorted[j] = sorted[j];

    int used = 0;
    for (size_t i : picked) {
      qmlb_op o = p->ops[i];
      ps.matoff.push_back(used);
      used += op_entries(p, o);
      if (o.kind != QMLB_OP_DIAG)
        for (int j = 0; j < o.k; ++j) o.bits[j] = pos[o.bits[j]];
      if (o.kind == QMLB_OP_PERM) {
        const int D = 1 << o.k;
        std::vector<int> perm(D);
        for (int v = 0; v < D; ++v) perm[v] = (int)p->consts[o.aux + v];
        if (o.k == 2 && o.bits[0] < o.bits[1]) {  // canonical order JA > JB
          auto sw = [](int v) { return ((v & 1) << 1) | (v >> 1); };
          std::vector<int> q(4);
          for (int v = 0; v < 4; ++v) q[sw(v)] = sw(perm[v]);
          perm = q;
          std::swap(o.bits[0], o.bits[1]);
        }
        unsigned long long packed = 0;
        for (int v = 0; v < D; ++v) packed |= (unsigned long long)perm[v] << (o.k * v);
        o.src = (int32_t)(uint32_t)(packed & 0xffffffffull);
        o.aux = (int32_t)(uint32_t)(packed >> 32);
      }
      ps.ops.push_back(o);
      ps.src_index.push_back((int)i);
      done[i] = 1;
    }
    remaining -= picked.size();
    ps.matw = std::max(used, 1);
    ps.flags = first_pass ? QMLB_PASS_INIT : 0;
    for (const qmlb_op& o : ps.ops)
      if (o.kind != QMLB_OP_DIAG && o.k >= 3) ps.flags |= QMLB_PASS_HEAVY;
    StreamPass& pd = ps.dev;
    std::memset(&pd, 0, sizeof(pd));
    pd.n_ops = (int)ps.ops.size();
    pd.n_bits = N;
    pd.flags = ps.flags;
    pd.matw = ps.matw;
    for (int j = 0; j < R; ++j) {
      pd.gb[j] = ps.gb[j];
      pd.sorted[j] = ps.sorted[j];
    }
    for (size_t i = 0; i < ps.ops.size(); ++i) {
      const qmlb_op& o = ps.ops[i];
      StreamOp& so = pd.ops[i];
      so.kind = (uint8_t)o.kind;
      so.k = (uint8_t)o.k;
      so.b0 = (uint8_t)o.bits[0];
      so.b1 = (uint8_t)(o.k > 1 ? o.bits[1] : 0);
      so.src = o.kind == QMLB_OP_PERM ? -1 : o.src;
      so.data = 0;
      if (o.kind == QMLB_OP_PERM) {
        so.data = ((uint64_t)(uint32_t)o.aux << 32) | (uint32_t)o.src;
      } else if (o.kind == QMLB_OP_DIAG) {
        for (int j = 0; j < o.k; ++j) so.data |= (uint64_t)o.bits[j] << (6 * j);
      }
      pd.matoff[i] = (uint16_t)ps.matoff[i];
    }
    p->stream_passes.push_back(std::move(ps));
    first_pass = false;
  }
  return QMLB_OK;
}

template <typename V>
size_t place(size_t& off, const std::vector<V>& v) {
  off = (off + 15) & ~size_t(15);
  size_t at = off;
  off += v.size() * sizeof(V);
  return at;
}

int upload(qmlb_program* p) {
  size_t off = 0;
  size_t o_ops = place(off, p->ops), o_src = place(off, p->sources),
         o_items = place(off, p->items), o_ang = place(off, p->angles),
         o_terms = place(off, p->terms), o_consts = place(off, p->consts),
         o_obs = place(off, p->obs), o_oc = place(off, p->obs_consts),
         o_pre = place(off, p->pre);
  // register kernel: compact op stream (staged in shared memory by every CTA) with, per
  // op, the one / two hoisted-factor tables its matrix is read from
  std::vector<RegOp> fast;
  if (p->strategy == 0) {
    fast.assign(p->ops.size(), RegOp{});
    for (size_t i = 0; i < p->ops.size(); ++i) {
      const qmlb_op& o = p->ops[i];
      RegOp f{};
      f.kind = (uint8_t)o.kind;
      f.k = (uint8_t)o.k;
      f.b0 = (uint8_t)o.bits[0];
      f.b1 = (uint8_t)(o.k > 1 ? o.bits[1] : 0);
      f.src = o.src;
      if (o.kind == QMLB_OP_PERM && o.k == 2) {  // CX: (control, target)
        if (reg_cx_orientation(p->consts.data() + o.aux) == 1) std::swap(f.b0, f.b1);
      }
      if (o.kind == QMLB_OP_MAT || o.kind == QMLB_OP_CTRL1) {
        const qmlb_source& s = p->sources[o.src];
        if (s.kind == QMLB_SRC_PRE) {
          f.n = 1;
          f.slot0 = s.a1;
          f.local0 = s.a0;
        } else if (s.kind == QMLB_SRC_CHAIN && s.a1 == 2 &&
                   p->sources[p->items[s.a0]].kind == QMLB_SRC_PRE &&
                   p->sources[p->items[s.a0 + 1]].kind == QMLB_SRC_PRE) {
          const qmlb_source& x = p->sources[p->items[s.a0]];
          const qmlb_source& y = p->sources[p->items[s.a0 + 1]];
          f.n = 2;
          f.slot0 = x.a1;
          f.local0 = x.a0;
          f.slot1 = y.a1;
          f.local1 = y.a0;
        }
      }
      fast[i] = f;
    }
    p->reg_ops_host = fast;
  }
  const size_t o_fast = place(off, fast);
  const size_t o_matlist = place(off, p->stream_matlist);
  const size_t o_fsteps = place(off, p->frame_steps);
  size_t o_ids[QMLB_MAX_ARGS];
  for (int a = 0; a < QMLB_MAX_ARGS; ++a) o_ids[a] = place(off, p->pre_ids[a]);
  struct PO {
    size_t ops, matoff, win;
  };
  std::vector<PO> po(p->passes.size());
  for (size_t i = 0; i < p->passes.size(); ++i) {
    po[i].ops = place(off, p->passes[i].ops);
    po[i].matoff = place(off, p->passes[i].matoff);
    po[i].win = place(off, p->passes[i].windows);
  }
  off = (off + 15) & ~size_t(15);
  std::vector<unsigned char> host(off + 16, 0);
  auto put = [&](size_t at, const void* src, size_t n) {
    if (n) std::memcpy(host.data() + at, src, n);
  };
  put(o_ops, p->ops.data(), p->ops.size() * sizeof(qmlb_op));
  put(o_src, p->sources.data(), p->sources.size() * sizeof(qmlb_source));
  put(o_items, p->items.data(), p->items.size() * sizeof(int32_t));
  put(o_ang, p->angles.data(), p->angles.size() * sizeof(qmlb_angle));
  put(o_terms, p->terms.data(), p->terms.size() * sizeof(qmlb_term));
  put(o_consts, p->consts.data(), p->consts.size() * sizeof(double));
  put(o_obs, p->obs.data(), p->obs.size() * sizeof(qmlb_obs));
  put(o_oc, p->obs_consts.data(), p->obs_consts.size() * sizeof(double));
  put(o_pre, p->pre.data(), p->pre.size() * sizeof(qmlb_pre));
  put(o_fast, fast.data(), fast.size() * sizeof(RegOp));
  put(o_matlist, p->stream_matlist.data(), p->stream_matlist.size() * sizeof(StreamMatOp));
  put(o_fsteps, p->frame_steps.data(), p->frame_steps.size() * sizeof(FrameStep));
  for (int a = 0; a < QMLB_MAX_ARGS; ++a)
    put(o_ids[a], p->pre_ids[a].data(), p->pre_ids[a].size() * sizeof(int32_t));
  for (size_t i = 0; i < p->passes.size(); ++i) {
    put(po[i].ops, p->passes[i].ops.data(), p->passes[i].ops.size() * sizeof(qmlb_op));
    put(po[i].matoff, p->passes[i].matoff.data(), p->passes[i].matoff.size() * 4);
    put(po[i].win, p->passes[i].windows.data(), p->passes[i].windows.size() * sizeof(int2));
  }
  CUDA_TRY(cudaMalloc(&p->blob, host.size()));
  CUDA_TRY(cudaMemcpy(p->blob, host.data(), host.size(), cudaMemcpyHostToDevice));
  unsigned char* base = static_cast<unsigned char*>(p->blob);
  DevProg& d = p->dev;
  d.ops = reinterpret_cast<const qmlb_op*>(base + o_ops);
  d.src = reinterpret_cast<const qmlb_source*>(base + o_src);
  d.items = reinterpret_cast<const int32_t*>(base + o_items);
  d.ang = reinterpret_cast<const qmlb_angle*>(base + o_ang);
  d.terms = reinterpret_cast<const qmlb_term*>(base + o_terms);
  d.consts = reinterpret_cast<const double*>(base + o_consts);
  d.obs = reinterpret_cast<const qmlb_obs*>(base + o_obs);
  d.obs_consts = reinterpret_cast<const double*>(base + o_oc);
  d.pre = reinterpret_cast<const qmlb_pre*>(base + o_pre);
  d.n_pre = (int)p->pre.size();
  d.rops = fast.empty() ? nullptr : reinterpret_cast<const RegOp*>(base + o_fast);
  p->stream_matlist_dev = reinterpret_cast<const StreamMatOp*>(base + o_matlist);
  p->frame_steps_dev = reinterpret_cast<const FrameStep*>(base + o_fsteps);
  for (int a = 0; a < QMLB_MAX_ARGS; ++a)
    p->pre_ids_dev[a] = reinterpret_cast<const int32_t*>(base + o_ids[a]);
  d.n_ops = (int)p->ops.size();
  d.n_obs = (int)p->obs.size();
  d.n_bits = p->n_bits;
  d.n_qubits = p->n_qubits;
  d.density = p->density;
  d.out_type = p->out_type;
  for (size_t i = 0; i < p->passes.size(); ++i) {
    QmlbPassHost& ps = p->passes[i];
    PassDev& pd = ps.dev;
    pd.ops = reinterpret_cast<const qmlb_op*>(base + po[i].ops);
    pd.matoff = reinterpret_cast<const int32_t*>(base + po[i].matoff);
    pd.windows = reinterpret_cast<const int2*>(base + po[i].win);
    pd.n_windows = (int)ps.windows.size();
    pd.k_tile = (int)ps.tile_bits.size();
    pd.n_bits = p->n_bits;
    pd.flags = ps.flags;
    pd.matw = ps.matw;
    pd.identity_map = 1;
    for (size_t j = 0; j < ps.tile_bits.size(); ++j) {
      pd.tile_bits[j] = ps.tile_bits[j];
      if (ps.tile_bits[j] != (int)j) pd.identity_map = 0;
    }
  }
  return QMLB_OK;
}

bool z1_fast(const qmlb_program* p);

int plan(qmlb_program* p) {
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
        sms > 0)
      p->sm_count = sms;
  }
  const size_t cs = cs_of(p->dtype);
  const int N = p->n_bits;
  const int force = p->force_stream ? 2 : env_int("QMLB_FORCE_STRATEGY", -1);

  // ---- strategy 0: registers --------------------------------------------------
  bool reg_ok = !p->density && N <= REG_MAX_BITS;
  for (const auto& o : p->ops) reg_ok = reg_ok && reg_supports(o, p->consts.data(), N);
  if ((reg_ok && force < 0) || (reg_ok && force == 0)) {
    p->strategy = 0;
    if (p->out_type == QMLB_OUT_STATE) {
      p->direct_out = true;
      p->reg_mode = 0;
    } else if (p->out_type == QMLB_OUT_PROBS) {
      p->direct_out = true;
      p->reg_mode = 1;
    } else if (p->out_type == QMLB_OUT_EXPVAL && all_zstring(p)) {
      p->direct_out = true;
      p->reg_mode = 2;
    } else {
      p->direct_out = false;
      p->reg_mode = 0;
    }
    return QMLB_OK;
  }

  // ---- strategy 5: Pauli-basis frame engine for noisy density programs ----------------
  if ((force < 0 || force == 5) && !p->force_stream && p->density && env_int("QMLB_PTM", 1)) {
    if (plan_frame_ptm(p) == QMLB_OK) {
      p->strategy = 5;
      p->direct_out = true;
      p->frame_out_mode =
          p->out_type == QMLB_OUT_PROBS ? 1 : (p->out_type == QMLB_OUT_DENSITY ? 0 : 2);
      return QMLB_OK;
    }
    if (force == 5) return fail(QMLB_ERR_UNSUPPORTED, "program outside the Pauli-basis engine");
  }

  // ---- strategy 3: on-chip frame engine (state in shared memory / cluster DSMEM) -----
  if ((force < 0 || force == 3) && !p->force_stream && env_int("QMLB_FRAME", 1)) {
    if (plan_frame(p) == QMLB_OK) {
      p->strategy = 3;
      p->direct_out = true;
      if (p->out_type == QMLB_OUT_STATE || (p->out_type == QMLB_OUT_DENSITY && p->density)) {
        p->frame_out_mode = 0;
      } else if (p->out_type == QMLB_OUT_PROBS) {
        p->frame_out_mode = 1;
      } else if (p->out_type == QMLB_OUT_EXPVAL && all_zstring(p)) {
        p->frame_out_mode = 2;
      } else {
        p->frame_out_mode = 0;  // state to the workspace, measured by the kernels below
        p->direct_out = false;
      }
      return QMLB_OK;
    }
    if (force == 3) return fail(QMLB_ERR_UNSUPPORTED, "program outside the frame engine's envelope");
  }

  // ---- strategy 4: streaming frame engine (tiles of an HBM-resident state) ------------
  if ((force < 0 || force == 4) && !p->force_stream && z1_fast(p) &&
      env_int("QMLB_FSTREAM", 1)) {
    if (plan_frame_stream(p) == QMLB_OK) {
      p->strategy = 4;
      p->direct_out = false;  // the state sits in the workspace; <Z_q> by the sweep below
      return QMLB_OK;
    }
    if (force == 4) return fail(QMLB_ERR_UNSUPPORTED, "program outside the streaming frame engine");
  }

  // ---- strategy 1: whole state in shared memory ---------------------------------
  const int init_max = env_int("QMLB_SMEM_STATE_BITS", p->dtype == QMLB_C128 ? 13 : 14);
  p->direct_out = (p->out_type == QMLB_OUT_STATE) ||
                  (p->out_type == QMLB_OUT_DENSITY && p->density);
  const int stream_r = 4;
  if ((N <= init_max && force != 2) || force == 1 || N < stream_r) {
    if (N > QMLB_MAX_TILE_BITS) return fail(QMLB_ERR_UNSUPPORTED, "state too large for smem");
    p->strategy = 1;
    QmlbPassHost ps;
    for (int g = 0; g < N; ++g) ps.tile_bits.push_back(g);
    ps.ops = p->ops;
    ps.flags = QMLB_PASS_INIT | QMLB_PASS_STORE;
    p->warp_team = N <= 7;
    p->teams = p->warp_team ? THREADS / 32 : 1;
    const size_t tile_bytes = (size_t(1) << N) * cs;
    // matrix buffer: whole circuit if it fits next to the state in <= 96 KB per CTA
    int need = 0, biggest = 1;
    for (const auto& o : ps.ops) {
      need += op_entries(p, o);
      biggest = std::max(biggest, op_entries(p, o));
    }
    size_t budget = 96 * 1024;
    size_t per_team = budget / p->teams;
    int cap = per_team > tile_bytes ? (int)((per_team - tile_bytes) / cs) : 0;
    cap = std::max(cap, std::max(biggest, 256));
    int matw = std::max(1, std::min(need, cap));
    make_windows(p, ps, matw);
    p->smem = p->teams * (tile_bytes + (size_t)matw * cs);
    p->passes.push_back(std::move(ps));
    return QMLB_OK;
  }

  // ---- strategy 2: streamed register-group passes over HBM ------------------------
  p->strategy = 2;
  p->stream_r = 4;  // register group: 16 amplitudes per thread (see qmlb_stream.cuh)
  p->warp_team = false;
  p->teams = 1;
  p->smem = 0;
  {
    int rc = schedule_stream(p, p->stream_r);
    if (rc != QMLB_OK) return rc;
  }
  {
    int row = 0;
    for (QmlbStreamPassHost& ps : p->stream_passes) {
      ps.dev.mat_base = row;
      for (size_t i = 0; i < ps.ops.size(); ++i) {
        const qmlb_op& o = ps.ops[i];
        if (o.kind == QMLB_OP_PERM) continue;
        StreamMatOp mo{};
        mo.src = o.src;
        mo.off = row + ps.matoff[i];
        mo.swap2 = (o.kind == QMLB_OP_MAT && o.k == 2 && o.bits[0] < o.bits[1]) ? 1 : 0;
        p->stream_matlist.push_back(mo);
      }
      row += ps.matw;
    }
    p->stream_mat_row = row;
    for (QmlbStreamPassHost& ps : p->stream_passes) ps.dev.mat_row = row;
  }
  if (env_int("QMLB_DUMP_PASSES", 0)) {
    static const char* kinds[] = {"MAT", "CTRL1", "PERM", "DIAG"};
    for (size_t i = 0; i < p->stream_passes.size(); ++i) {
      const QmlbStreamPassHost& ps = p->stream_passes[i];
      std::fprintf(stderr, "pass %zu flags=%d group=[", i, ps.flags);
      for (int j = 0; j < p->stream_r; ++j) std::fprintf(stderr, "%d ", ps.gb[j]);
      std::fprintf(stderr, "] ops:");
      for (const qmlb_op& o : ps.ops) {
        std::fprintf(stderr, " %s%d(", kinds[o.kind], o.k);
        for (int j = 0; j < o.k; ++j) std::fprintf(stderr, j ? ",%d" : "%d", o.bits[j]);
        std::fprintf(stderr, ")");
      }
      std::fprintf(stderr, "\n");
    }
  }
  return QMLB_OK;
}

// every observable is Z on one qubit of a pure state with >= 256 amplitudes
bool z1_fast(const qmlb_program* p) {
  if (p->out_type != QMLB_OUT_EXPVAL || p->density || p->n_qubits < 8 || p->n_qubits > 32)
    return false;
  for (const auto& o : p->obs)
    if (o.kind != QMLB_OBS_ZSTRING || __builtin_popcountll((unsigned long long)o.zmask) != 1)
      return false;
  return true;
}

int z1_ctas(const qmlb_program* p, int64_t batch) {
  const int64_t units = int64_t(1) << (p->n_qubits - 8);
  const int64_t per = std::max<int64_t>(1, ((int64_t)p->sm_count * 8) / std::max<int64_t>(batch, 1));
  return (int)std::max<int64_t>(1, std::min<int64_t>((units + 7) / 8, per));
}

int expval_chunks(const qmlb_program* p, int64_t batch) {
  const int64_t dim = int64_t(1) << p->n_qubits;
  if (p->out_type != QMLB_OUT_EXPVAL) return 1;
  if (batch < (int64_t)p->sm_count * 4 && dim > 4096) {
    int chunks = (int)std::min<int64_t>(dim / 4096,
                                        ((int64_t)p->sm_count * 4 + batch - 1) / batch);
    return std::max(chunks, 1);
  }
  return 1;
}

template <typename T>
int launch_measure(const qmlb_program* p, const cx<T>* state, int64_t batch, void* out,
                   unsigned char* scratch, cudaStream_t st) {
  const int64_t dim = int64_t(1) << p->n_qubits;
  if (p->out_type == QMLB_OUT_PROBS) {
    int64_t total = batch * dim;
    int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)p->sm_count * 16);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    k_probs<T><<<std::max(grid, 1), 256, 0, st>>>(state, static_cast<T*>(out), batch,
                                                 p->n_qubits, p->density);
  } else if (p->out_type == QMLB_OUT_DENSITY) {
    int64_t total = batch * dim * dim;
    int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)p->sm_count * 16);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    k_outer<T><<<std::max(grid, 1), 256, 0, st>>>(state, static_cast<cx<T>*>(out), batch,
                                                 p->n_qubits);
  } else if (p->out_type == QMLB_OUT_EXPVAL && z1_fast(p)) {
    const int ctas = z1_ctas(p, batch);
    double* partial = reinterpret_cast<double*>(scratch);
    const int64_t n_out = batch * (int64_t)p->obs.size();
    g_launches.fetch_add(2, std::memory_order_relaxed);
    for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
      const int64_t nb = std::min<int64_t>(65535, batch - b0);
      k_expval_z1<T><<<dim3((unsigned)ctas, (unsigned)nb), 256, 0, st>>>(
          state + ((size_t)b0 << p->n_qubits), partial + (size_t)b0 * ctas * 33, p->n_qubits);
    }
    BitMap map;
    for (int b = 0; b < 40; ++b)
      map.pos[b] = (int8_t)(p->strategy == 4 && b < (int)p->fstream_final_hpos.size()
                                ? p->fstream_final_hpos[b]
                                : b);
    k_expval_z1_final<T><<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(
        p->dev, partial, static_cast<T*>(out), batch, ctas, map);
  } else if (p->out_type == QMLB_OUT_EXPVAL) {
    const int chunks = expval_chunks(p, batch);
    const int n_obs = (int)p->obs.size();
    if (chunks == 1) {
      g_launches.fetch_add(1, std::memory_order_relaxed);
    k_expval<T><<<(unsigned)batch, 256, 0, st>>>(p->dev, state, static_cast<T*>(out),
                                                  batch, 1);
    } else {
      T* partial = reinterpret_cast<T*>(scratch);
      g_launches.fetch_add(1, std::memory_order_relaxed);
    k_expval<T><<<(unsigned)(batch * chunks), 256, 0, st>>>(p->dev, state, partial, batch,
                                                             chunks);
      int64_t n = batch * n_obs;
      g_launches.fetch_add(1, std::memory_order_relaxed);
    k_sum_chunks<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
          partial, static_cast<T*>(out), n, chunks);
    }
  }
  CUDA_TRY(cudaGetLastError());
  return QMLB_OK;
}

// Hoisted-factor tables: slot a gets one when it has fewer distinct rows than half the
// batch (otherwise its factors are evaluated inline).  Returns the bytes used.
size_t pre_layout(const qmlb_program* p, const qmlb_arg* a, int64_t batch, bool on[QMLB_MAX_ARGS],
                  size_t off[QMLB_MAX_ARGS]) {
  size_t total = 0;
  for (int s = 0; s < QMLB_MAX_ARGS; ++s) {
    on[s] = false;
    off[s] = 0;
    const size_t n = p->pre_ids[s].size();
    if (n == 0 || a[s].mod * 2 > batch) continue;
    on[s] = true;
    off[s] = total;
    total += (n * (size_t)a[s].mod * 4 * cs_of(p->dtype) + 255) & ~size_t(255);
  }
  return total;
}

// batched streaming runs keep a table of every (element, op) matrix
size_t premats_bytes(const qmlb_program* p, int64_t batch) {
  if (!(p->strategy == 3 || p->strategy == 4 || p->strategy == 5 ||
        (p->strategy == 2 && batch > 1)))
    return 0;
  return ((size_t)batch * p->stream_mat_row * cs_of(p->dtype) + 255) & ~size_t(255);
}

size_t state_layout(const qmlb_program* p, int64_t batch, size_t* part_off) {
  size_t need = p->direct_out ? 0 : (size_t)batch * (size_t(1) << p->n_bits) * cs_of(p->dtype);
  *part_off = (need + 255) & ~size_t(255);
  const int chunks = expval_chunks(p, batch);
  if (!p->direct_out && z1_fast(p))
    need = *part_off + (size_t)batch * z1_ctas(p, batch) * 33 * sizeof(double);
  else if (!p->direct_out && chunks > 1)
    need = *part_off + (size_t)batch * p->obs.size() * chunks * rs_of(p->dtype);
  return ((need + 255) & ~size_t(255)) + premats_bytes(p, batch);
}

// fills the hoisted-factor tables at the start of `workspace`; returns their size
template <typename T>
int prepare_tables(const qmlb_program* p, RunArgs& R, void* workspace, size_t ws_bytes,
                   cudaStream_t st, size_t* tab_bytes_out) {
  bool on[QMLB_MAX_ARGS];
  size_t toff[QMLB_MAX_ARGS];
  const size_t tab_bytes = pre_layout(p, R.a, R.batch, on, toff);
  if (tab_bytes > ws_bytes) return fail(QMLB_ERR_WORKSPACE, "workspace too small");
  unsigned char* wsb = static_cast<unsigned char*>(workspace);
  for (int s = 0; s < QMLB_MAX_ARGS; ++s) {
    R.pre_on[s] = on[s];
    R.pre_tab[s] = on[s] ? wsb + toff[s] : nullptr;
  }
  for (int s = 0; s < QMLB_MAX_ARGS; ++s) {
    if (!on[s]) continue;
    const int n_ids = (int)p->pre_ids[s].size();
    const int64_t total = (int64_t)n_ids * R.a[s].mod;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    k_pre<T><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(
        p->dev, R, s, p->pre_ids_dev[s], n_ids, reinterpret_cast<cx<T>*>(wsb + toff[s]));
  }
  *tab_bytes_out = tab_bytes;
  return QMLB_OK;
}

// the streamed gate passes over `state`; init_mode 1: |0..0>, 2: zero vector, 0: continue
template <typename T>
int evolve_stream(const qmlb_program* p, const RunArgs& R, void* state, int init_mode,
                  void* premats, cudaStream_t st, const StreamPeers* peers = nullptr) {
  if (premats)
    CUDA_TRY((std::is_same<T, double>::value ? launch_stream_mats_f64 : launch_stream_mats_f32)(
        p, R, premats, st));
  const int64_t items = int64_t(1) << (p->n_bits - p->stream_r);
  const int64_t ctas_x = (items + STREAM_THREADS - 1) / STREAM_THREADS;
  const int64_t want = (int64_t)p->sm_count * 16;  // persistent: CTAs loop over items / elements
  dim3 grid;
  if (R.batch == 1) {
    grid = dim3((unsigned)std::max<int64_t>(1, std::min(ctas_x, want)), 1, 1);
  } else {
    // every CTA evaluates the pass' matrices for its element once: give it enough items
    // to amortise that (>= 8 trips) as long as the batch alone fills the GPU
    int64_t gx = std::max<int64_t>(1, std::min<int64_t>(ctas_x, 64));
    while (gx > 1 && R.batch * (gx / 2) >= want && ctas_x / gx < 8) gx /= 2;
    const int64_t gy =
        std::max<int64_t>(1, std::min<int64_t>(R.batch, std::max<int64_t>(1, want / gx)));
    grid = dim3((unsigned)gx, (unsigned)std::min<int64_t>(gy, 65535), 1);
  }
  for (const QmlbStreamPassHost& ps : p->stream_passes) {
    StreamPass pass = ps.dev;
    if (pass.flags & QMLB_PASS_INIT) {
      if (init_mode == 0) pass.flags &= ~QMLB_PASS_INIT;
      if (init_mode == 2) pass.flags |= QMLB_PASS_INIT_ZERO;
    }
    CUDA_TRY((std::is_same<T, double>::value ? launch_stream_f64 : launch_stream_f32)(
        p, R, pass, grid, state, premats, peers, st));
    peers = nullptr;  // only the first pass of the epoch pulls from the peers
  }
  return QMLB_OK;
}

template <typename T>
int run_typed(const qmlb_program* p, RunArgs& R, void* out, void* workspace, size_t ws_bytes,
              cudaStream_t st) {
  size_t tab_bytes = 0, part_off = 0;
  {
    bool on[QMLB_MAX_ARGS];
    size_t toff[QMLB_MAX_ARGS];
    const size_t need =
        pre_layout(p, R.a, R.batch, on, toff) + state_layout(p, R.batch, &part_off);
    if (need > ws_bytes) return fail(QMLB_ERR_WORKSPACE, "workspace too small");
  }
  {
    int rc = prepare_tables<T>(p, R, workspace, ws_bytes, st, &tab_bytes);
    if (rc != QMLB_OK) return rc;
  }
  unsigned char* ws_state = static_cast<unsigned char*>(workspace) + tab_bytes;
  cx<T>* state = p->direct_out ? static_cast<cx<T>*>(out) : reinterpret_cast<cx<T>*>(ws_state);

  if (p->strategy == 0) {
    void* dst = p->direct_out ? out : static_cast<void*>(ws_state);
    CUDA_TRY((std::is_same<T, double>::value ? launch_reg_f64 : launch_reg_f32)(p, R, dst, st));
  } else if (p->strategy == 2) {
    // [tables | state + partials | premats]
    void* premats = nullptr;
    if (premats_bytes(p, R.batch))
      premats = static_cast<unsigned char*>(workspace) + tab_bytes +
                (state_layout(p, R.batch, &part_off) - premats_bytes(p, R.batch));
    int rc = evolve_stream<T>(p, R, state, 1, premats, st);
    if (rc != QMLB_OK) return rc;
  } else if (p->strategy == 4) {
    unsigned char* premats = static_cast<unsigned char*>(workspace) + tab_bytes +
                             (state_layout(p, R.batch, &part_off) - premats_bytes(p, R.batch));
    const bool f64 = std::is_same<T, double>::value;
    CUDA_TRY((f64 ? launch_stream_mats_f64 : launch_stream_mats_f32)(p, R, premats, st));
    CUDA_TRY((f64 ? launch_fstream_f64 : launch_fstream_f32)(p, R, state, premats, 1, st));
  } else if (p->strategy == 5) {
    unsigned char* premats = static_cast<unsigned char*>(workspace) + tab_bytes +
                             (state_layout(p, R.batch, &part_off) - premats_bytes(p, R.batch));
    const bool f64 = std::is_same<T, double>::value;
    CUDA_TRY((f64 ? launch_stream_mats_f64 : launch_stream_mats_f32)(p, R, premats, st));
    CUDA_TRY((f64 ? launch_frame_ptm_f64 : launch_frame_ptm_f32)(p, R, premats, out,
                                                                   p->frame_out_mode, st));
  } else if (p->strategy == 3) {
    // [tables | state + partials | evaluated matrices]: one launch evaluates every matrix
    // of every element, one launch runs the whole tape on chip
    unsigned char* premats = static_cast<unsigned char*>(workspace) + tab_bytes +
                             (state_layout(p, R.batch, &part_off) - premats_bytes(p, R.batch));
    const bool f64 = std::is_same<T, double>::value;
    CUDA_TRY((f64 ? launch_stream_mats_f64 : launch_stream_mats_f32)(p, R, premats, st));
    CUDA_TRY((f64 ? launch_frame_f64 : launch_frame_f32)(
        p, R, premats, p->direct_out ? out : static_cast<void*>(ws_state), p->frame_out_mode, st));
  } else {
    for (const QmlbPassHost& ps : p->passes) {
      const int kt = (int)ps.tile_bits.size();
      const int64_t tiles = R.batch << (p->n_bits - kt);
      const int64_t blocks_needed = (tiles + p->teams - 1) / p->teams;
      const unsigned grid = (unsigned)std::max<int64_t>(
          1, std::min<int64_t>(blocks_needed, (int64_t)p->sm_count * 8));
      CUDA_TRY((std::is_same<T, double>::value ? launch_tile_f64 : launch_tile_f32)(
          p, R, ps.dev, grid, state, st));
    }
  }
  if (!p->direct_out)
    return launch_measure<T>(p, state, R.batch, out, ws_state + part_off, st);
  return QMLB_OK;
}

int set_smem_attr(const qmlb_program* p) {
  if (p->strategy != 1 || p->smem <= 48 * 1024) return QMLB_OK;
  CUDA_TRY(p->dtype == QMLB_C128 ? tile_set_smem_f64(p->smem) : tile_set_smem_f32(p->smem));
  return QMLB_OK;
}

}  // namespace

extern "C" {

int qmlb_version(void) { return QMLB_VERSION; }

unsigned long long qmlb_launch_count(void) { return g_launches.load(); }

const char* qmlb_last_error(void) { return g_err.c_str(); }

// validate + copy + plan (host only, no CUDA allocation)
static int build_host_program(const qmlb_program_desc* d, qmlb_program* p);

int qmlb_plan_describe(const qmlb_program_desc* d, char* buf, size_t buflen) {
  if (!d || !buf || buflen < 2) return fail(QMLB_ERR_INVALID, "null argument");
  qmlb_program prog;
  int rc = build_host_program(d, &prog);
  if (rc != QMLB_OK) return rc;
  std::string s = "strategy " + std::to_string(prog.strategy) + "\n";
  if (prog.strategy == 2) {
    for (const QmlbStreamPassHost& ps : prog.stream_passes) {
      s += "pass flags " + std::to_string(ps.flags) + " group";
      for (int j = 0; j < prog.stream_r; ++j) s += " " + std::to_string(ps.gb[j]);
      s += " ops";
      for (size_t i = 0; i < ps.ops.size(); ++i) {
        const qmlb_op& o = ps.ops[i];
        s += " " + std::to_string(ps.src_index[i]) + ":" + std::to_string(o.kind) + ":";
        for (int j = 0; j < o.k; ++j) s += (j ? "," : "") + std::to_string(o.bits[j]);
      }
      s += "\n";
    }
  } else if (prog.strategy == 3 || prog.strategy == 4 || prog.strategy == 5) {
    s += describe_frame(&prog);
  } else if (prog.strategy == 1) {
    s += "smem_bytes " + std::to_string(prog.smem) + " teams " + std::to_string(prog.teams) + "\n";
  }
  if (s.size() + 1 > buflen) return fail(QMLB_ERR_WORKSPACE, "description buffer too small");
  std::memcpy(buf, s.c_str(), s.size() + 1);
  return QMLB_OK;
}

int qmlb_program_create(const qmlb_program_desc* d, qmlb_program** out) {
  if (!d || !out) return fail(QMLB_ERR_INVALID, "null argument");
  *out = nullptr;
  qmlb_program* p = new qmlb_program();
  int rc = build_host_program(d, p);
  if (rc == QMLB_OK) rc = upload(p);
  if (rc == QMLB_OK) rc = set_smem_attr(p);
  if (rc != QMLB_OK) {
    if (p->blob) cudaFree(p->blob);
    delete p;
    return rc;
  }
  *out = p;
  return QMLB_OK;
}

static int build_host_program(const qmlb_program_desc* d, qmlb_program* p) {
  int rc = validate(d, p);
  if (rc != QMLB_OK) return rc;
  p->n_qubits = d->n_qubits;
  p->n_bits = d->n_bits;
  p->density = d->density;
  p->dtype = d->dtype;
  p->out_type = d->out_type;
  p->force_stream = (d->reserved & QMLB_DESC_FORCE_STREAM) != 0;
  p->ops.assign(d->ops, d->ops + d->n_ops);
  p->sources.assign(d->sources, d->sources + d->n_sources);
  p->items.assign(d->items, d->items + d->n_items);
  p->angles.assign(d->angles, d->angles + d->n_angles);
  p->terms.assign(d->terms, d->terms + d->n_terms);
  p->consts.assign(d->consts, d->consts + d->n_consts);
  p->obs.assign(d->obs, d->obs + d->n_obs);
  p->obs_consts.assign(d->obs_consts, d->obs_consts + d->n_obs_consts);
  if (d->n_pre > 0) p->pre.assign(d->pre, d->pre + d->n_pre);
  for (int i = 0; i < (int)p->pre.size(); ++i) p->pre_ids[p->pre[i].arg].push_back(i);
  return plan(p);
}

int qmlb_program_destroy(qmlb_program* p) {
  if (!p) return QMLB_OK;
  if (p->blob) cudaFree(p->blob);
  delete p;
  return QMLB_OK;
}

int qmlb_program_info(const qmlb_program* p, int32_t* strategy, int32_t* n_passes,
                      int32_t* n_device_ops) {
  if (!p) return fail(QMLB_ERR_INVALID, "null program");
  if (strategy) *strategy = p->strategy;
  if (n_passes)
    *n_passes = p->strategy == 0   ? 1
                : p->strategy == 2 ? (int32_t)p->stream_passes.size()
                : (p->strategy == 3 || p->strategy == 5) ? (int32_t)p->frame_steps.size()
                : p->strategy == 4 ? (int32_t)p->fstream_passes.size()
                                   : (int32_t)p->passes.size();
  if (n_device_ops) *n_device_ops = (int32_t)p->ops.size();
  return QMLB_OK;
}

size_t qmlb_workspace_bytes(const qmlb_program* p, const qmlb_arg* args, int32_t n_args,
                            int64_t batch) {
  if (!p || batch <= 0) return 0;
  qmlb_arg a[QMLB_MAX_ARGS];
  for (int i = 0; i < QMLB_MAX_ARGS; ++i) {
    a[i].ptr = nullptr;
    a[i].stride = 0;
    a[i].div = 1;
    a[i].mod = (i < n_args && args && args[i].mod >= 1) ? args[i].mod : 1;
  }
  bool on[QMLB_MAX_ARGS];
  size_t toff[QMLB_MAX_ARGS], part_off = 0;
  return pre_layout(p, a, batch, on, toff) + state_layout(p, batch, &part_off);
}

static int fill_run_args(const qmlb_program* p, const qmlb_arg* args, int32_t n_args,
                         int64_t batch, int64_t batch_offset, RunArgs& R) {
  if (n_args < 0 || n_args > QMLB_MAX_ARGS) return fail(QMLB_ERR_INVALID, "bad n_args");
  if (p->max_arg >= n_args) return fail(QMLB_ERR_INVALID, "program needs more arguments");
  std::memset(&R, 0, sizeof(R));
  for (int i = 0; i < QMLB_MAX_ARGS; ++i) {
    R.a[i].div = 1;
    R.a[i].mod = 1;
  }
  for (int i = 0; i < n_args; ++i) {
    R.a[i] = args[i];
    if (R.a[i].div < 1 || R.a[i].mod < 1) return fail(QMLB_ERR_INVALID, "arg div/mod < 1");
    if (R.a[i].mod > 0x7fffffffLL) return fail(QMLB_ERR_UNSUPPORTED, "argument with >= 2^31 rows");
  }
  for (const auto& t : p->terms)
    if (!R.a[t.arg].ptr) return fail(QMLB_ERR_INVALID, "program reads a NULL argument");
  R.batch = batch;
  R.batch_offset = batch_offset;
  return QMLB_OK;
}

int qmlb_evolve(const qmlb_program* p, const qmlb_arg* args, int32_t n_args, int64_t batch,
                int64_t batch_offset, void* state, int32_t init_mode, void* workspace,
                size_t workspace_bytes, void* stream) {
  if (!p || !state) return fail(QMLB_ERR_INVALID, "null program or state");
  if (p->strategy != 2)
    return fail(QMLB_ERR_INVALID, "qmlb_evolve needs a streaming program (QMLB_DESC_FORCE_STREAM)");
  if (init_mode < 0 || init_mode > 2) return fail(QMLB_ERR_INVALID, "bad init_mode");
  if (batch <= 0) return QMLB_OK;
  RunArgs R;
  int rc = fill_run_args(p, args, n_args, batch, batch_offset, R);
  if (rc != QMLB_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t tab = 0;
  if (p->dtype == QMLB_C128) {
    rc = prepare_tables<double>(p, R, workspace, workspace_bytes, st, &tab);
    return rc != QMLB_OK ? rc : evolve_stream<double>(p, R, state, init_mode, nullptr, st);
  }
  rc = prepare_tables<float>(p, R, workspace, workspace_bytes, st, &tab);
  return rc != QMLB_OK ? rc : evolve_stream<float>(p, R, state, init_mode, nullptr, st);
}

int qmlb_evolve_peer(const qmlb_program* p, const qmlb_arg* args, int32_t n_args, void* dst_state,
                     const void* const* peer_src, int32_t n_peers, int32_t rank,
                     void* workspace, size_t workspace_bytes, void* stream) {
  if (!p || !dst_state || !peer_src) return fail(QMLB_ERR_INVALID, "null argument");
  if (p->strategy != 2)
    return fail(QMLB_ERR_INVALID, "qmlb_evolve_peer needs a streaming program");
  if (n_peers < 2 || n_peers > 8 || (n_peers & (n_peers - 1)) || rank < 0 || rank >= n_peers)
    return fail(QMLB_ERR_INVALID, "peer count must be 2, 4 or 8 and rank inside it");
  int g = 0;
  while ((1 << g) < n_peers) ++g;
  if (p->n_bits < 2 * g || p->n_bits - g < 1)
    return fail(QMLB_ERR_INVALID, "shard too small for the exchange");
  if (p->stream_passes.empty() || (p->stream_passes[0].ops.empty()))
    return fail(QMLB_ERR_INVALID, "epoch without operations");
  RunArgs R;
  int rc = fill_run_args(p, args, n_args, 1, 0, R);
  if (rc != QMLB_OK) return rc;
  StreamPeers peers{};
  for (int i = 0; i < n_peers; ++i) {
    if (!peer_src[i]) return fail(QMLB_ERR_INVALID, "null peer pointer");
    peers.ptr[i] = peer_src[i];
  }
  peers.enabled = 1;
  peers.cshift = p->n_bits - g;
  peers.rank = rank;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t tab = 0;
  if (p->dtype == QMLB_C128) {
    rc = prepare_tables<double>(p, R, workspace, workspace_bytes, st, &tab);
    return rc != QMLB_OK ? rc : evolve_stream<double>(p, R, dst_state, 0, nullptr, st, &peers);
  }
  rc = prepare_tables<float>(p, R, workspace, workspace_bytes, st, &tab);
  return rc != QMLB_OK ? rc : evolve_stream<float>(p, R, dst_state, 0, nullptr, st, &peers);
}

size_t qmlb_zsums_workspace_bytes(int64_t batch, int32_t n_bits) {
  if (batch <= 0 || n_bits < 8) return 0;
  const int64_t units = int64_t(1) << (n_bits - 8);
  const int64_t per = std::max<int64_t>(1, (148 * 8) / batch);
  const int64_t ctas = std::max<int64_t>(1, std::min<int64_t>((units + 7) / 8, per));
  return (size_t)batch * ctas * 33 * sizeof(double);
}

int qmlb_zsums(const void* state, int dtype, int64_t batch, int32_t n_bits, double* out,
               void* workspace, size_t workspace_bytes, void* stream) {
  if (!state || !out) return fail(QMLB_ERR_INVALID, "null argument");
  if (n_bits < 8 || n_bits > 32) return fail(QMLB_ERR_UNSUPPORTED, "qmlb_zsums needs 8..32 bits");
  if (batch <= 0) return QMLB_OK;
  if (batch > 65535) return fail(QMLB_ERR_UNSUPPORTED, "qmlb_zsums batch > 65535");
  const size_t need = qmlb_zsums_workspace_bytes(batch, n_bits);
  if (need > workspace_bytes || !workspace) return fail(QMLB_ERR_WORKSPACE, "workspace too small");
  const int ctas = (int)(need / ((size_t)batch * 33 * sizeof(double)));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(workspace);
  g_launches.fetch_add(2, std::memory_order_relaxed);
  if (dtype == QMLB_C128)
    k_expval_z1<double><<<dim3((unsigned)ctas, (unsigned)batch), 256, 0, st>>>(
        static_cast<const cx<double>*>(state), partial, n_bits);
  else
    k_expval_z1<float><<<dim3((unsigned)ctas, (unsigned)batch), 256, 0, st>>>(
        static_cast<const cx<float>*>(state), partial, n_bits);
  k_zsums_final<<<(unsigned)((batch * 33 + 255) / 256), 256, 0, st>>>(partial, out, batch, ctas);
  CUDA_TRY(cudaGetLastError());
  return QMLB_OK;
}

int qmlb_run(const qmlb_program* p, const qmlb_arg* args, int32_t n_args, int64_t batch,
             int64_t batch_offset, void* out, void* workspace, size_t workspace_bytes,
             void* stream) {
  if (!p || !out) return fail(QMLB_ERR_INVALID, "null program or output");
  if (batch <= 0) return QMLB_OK;
  RunArgs R;
  {
    int rc = fill_run_args(p, args, n_args, batch, batch_offset, R);
    if (rc != QMLB_OK) return rc;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->dtype == QMLB_C128) return run_typed<double>(p, R, out, workspace, workspace_bytes, st);
  return run_typed<float>(p, R, out, workspace, workspace_bytes, st);
}

int qmlb_sample(const void* probs, int dtype, const double* uniforms, int64_t batch,
                int32_t n_qubits, int64_t shots, int32_t* counts, void* stream) {
  if (!probs || !uniforms || !counts) return fail(QMLB_ERR_INVALID, "null argument");
  if (n_qubits < 1 || n_qubits > 14) return fail(QMLB_ERR_UNSUPPORTED, "shots need n <= 14");
  if (batch <= 0) return QMLB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t dim = size_t(1) << n_qubits;
  CUDA_TRY(cudaMemsetAsync(counts, 0, (size_t)batch * dim * sizeof(int32_t), st));
  if (dtype == QMLB_C128) {
    size_t smem = dim * sizeof(double);
    if (smem > 48 * 1024)
      CUDA_TRY(cudaFuncSetAttribute(k_sample<double>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    k_sample<double><<<(unsigned)batch, 256, smem, st>>>(
        static_cast<const double*>(probs), uniforms, n_qubits, shots, counts);
  } else {
    size_t smem = dim * sizeof(float);
    if (smem > 48 * 1024)
      CUDA_TRY(cudaFuncSetAttribute(k_sample<float>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    k_sample<float><<<(unsigned)batch, 256, smem, st>>>(static_cast<const float*>(probs),
                                                        uniforms, n_qubits, shots, counts);
  }
  CUDA_TRY(cudaGetLastError());
  return QMLB_OK;
}

int qmlb_purity(const void* states, int dtype, int is_density, int64_t batch,
                int32_t n_qubits, void* out, void* stream) {
  if (!states || !out) return fail(QMLB_ERR_INVALID, "null argument");
  if (batch <= 0) return QMLB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (dtype == QMLB_C128)
    k_purity<double><<<(unsigned)batch, 256, 0, st>>>(
        static_cast<const cx<double>*>(states), is_density, n_qubits, static_cast<double*>(out));
  else
    k_purity<float><<<(unsigned)batch, 256, 0, st>>>(
        static_cast<const cx<float>*>(states), is_density, n_qubits, static_cast<float*>(out));
  CUDA_TRY(cudaGetLastError());
  return QMLB_OK;
}

int qmlb_overlap_fidelity(const void* states, int dtype, int64_t half, int32_t n_qubits,
                          void* out, void* stream) {
  if (!states || !out) return fail(QMLB_ERR_INVALID, "null argument");
  if (half <= 0) return QMLB_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (dtype == QMLB_C128)
    k_overlap<double><<<(unsigned)half, 256, 0, st>>>(
        static_cast<const cx<double>*>(states), half, n_qubits, static_cast<double*>(out));
  else
    k_overlap<float><<<(unsigned)half, 256, 0, st>>>(
        static_cast<const cx<float>*>(states), half, n_qubits, static_cast<float*>(out));
  CUDA_TRY(cudaGetLastError());
  return QMLB_OK;
}

int qmlb_grid_dft(const void* ev, int dtype, int32_t n_x, int64_t n_p, int32_t n_obs,
                  const int32_t* row_of, void* out, void* stream) {
  if (!ev || !out) return fail(QMLB_ERR_INVALID, "null argument");
  if (n_x < 1 || n_p < 1 || n_obs < 1) return fail(QMLB_ERR_INVALID, "empty grid");
  const size_t smem = (size_t)n_x * DFT_PCOLS * sizeof(double) + (size_t)n_x * sizeof(double2);
  if (smem > 200 * 1024) return fail(QMLB_ERR_UNSUPPORTED, "grid too long for the on-chip DFT");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((n_p + DFT_PCOLS - 1) / DFT_PCOLS);
  // one pass over the frequencies: (n_x / 2 + 1) x (column groups) threads, whole warps
  const unsigned dft_threads = (unsigned)std::min<int64_t>(
      1024, std::max<int64_t>(64, (((int64_t)(n_x / 2 + 1) * (DFT_PCOLS / DFT_TC) + 31) / 32) * 32));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (dtype == QMLB_C128) {
    if (smem > 48 * 1024)
      CUDA_TRY(cudaFuncSetAttribute(k_grid_dft<double>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_grid_dft<double><<<grid, dft_threads, smem, st>>>(static_cast<const double*>(ev), n_x, n_p, n_obs,
                                                row_of, static_cast<cx<double>*>(out));
  } else {
    if (smem > 48 * 1024)
      CUDA_TRY(cudaFuncSetAttribute(k_grid_dft<float>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_grid_dft<float><<<grid, dft_threads, smem, st>>>(static_cast<const float*>(ev), n_x, n_p, n_obs,
                                               row_of, static_cast<cx<float>*>(out));
  }
  CUDA_TRY(cudaGetLastError());
  return QMLB_OK;
}

int qmlb_coef_moments(const void* coef, int dtype, const int32_t* rows, int32_t K, int64_t n_p,
                      void* out, void* stream) {
  if (!coef || !rows || !out) return fail(QMLB_ERR_INVALID, "null argument");
  if (K < 1 || n_p < 1) return fail(QMLB_ERR_INVALID, "empty selection");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t total = (int64_t)K * K + 2 * K;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (dtype == QMLB_C128)
    k_coef_moments<double><<<(unsigned)total, 128, 0, st>>>(
        static_cast<const cx<double>*>(coef), rows, K, n_p, static_cast<double2*>(out));
  else
    k_coef_moments<float><<<(unsigned)total, 128, 0, st>>>(
        static_cast<const cx<float>*>(coef), rows, K, n_p, static_cast<double2*>(out));
  CUDA_TRY(cudaGetLastError());
  return QMLB_OK;
}

static int make_bitsel(int32_t n, const int32_t* keep, int32_t k, BitSel* sel) {
  if (!keep || n < 1 || n > 16 || k < 1 || k > n) return fail(QMLB_ERR_INVALID, "bad qubit subset");
  std::memset(sel, 0, sizeof(*sel));
  uint32_t seen = 0;
  for (int j = 0; j < k; ++j) {
    if (keep[j] < 0 || keep[j] >= n || (seen >> keep[j] & 1u) || (j && keep[j] <= keep[j - 1]))
      return fail(QMLB_ERR_INVALID, "kept wires must be distinct, ascending and inside the register");
    seen |= 1u << keep[j];
    sel->keep[j] = (int8_t)(n - 1 - keep[j]);  // wire q is bit n-1-q
  }
  int g = 0;
  for (int q = n - 1; q >= 0; --q)
    if (!(seen >> q & 1u)) sel->gone[g++] = (int8_t)(n - 1 - q);
  sel->k = k;
  sel->g = g;
  return QMLB_OK;
}

int qmlb_partial_trace(const void* rho, int dtype, int64_t batch, int32_t n, const int32_t* keep,
                       int32_t k, void* out, void* stream) {
  if (!rho || !out) return fail(QMLB_ERR_INVALID, "null argument");
  BitSel sel;
  int rc = make_bitsel(n, keep, k, &sel);
  if (rc != QMLB_OK) return rc;
  if (batch <= 0) return QMLB_OK;
  const int64_t total = batch << (2 * k);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (dtype == QMLB_C128)
    k_partial_trace<double><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        static_cast<const cx<double>*>(rho), batch, n, sel, static_cast<cx<double>*>(out));
  else
    k_partial_trace<float><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        static_cast<const cx<float>*>(rho), batch, n, sel, static_cast<cx<float>*>(out));
  CUDA_TRY(cudaGetLastError());
  return QMLB_OK;
}

int qmlb_marginal_probs(const void* probs, int dtype, int64_t batch, int32_t n,
                        const int32_t* keep, int32_t k, void* out, void* stream) {
  if (!probs || !out) return fail(QMLB_ERR_INVALID, "null argument");
  BitSel sel;
  int rc = make_bitsel(n, keep, k, &sel);
  if (rc != QMLB_OK) return rc;
  if (batch <= 0) return QMLB_OK;
  const int64_t total = batch << k;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (dtype == QMLB_C128)
    k_marginal_probs<double><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        static_cast<const double*>(probs), batch, n, sel, static_cast<double*>(out));
  else
    k_marginal_probs<float><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        static_cast<const float*>(probs), batch, n, sel, static_cast<float*>(out));
  CUDA_TRY(cudaGetLastError());
  return QMLB_OK;
}

size_t qmlb_allreduce_buffer_bytes(int64_t n) {
  return 256 + 3 * 8 * (size_t)std::max<int64_t>(n, 0) * sizeof(double);
}

int qmlb_allreduce_peer(const void* const* peer_buf, int32_t n_peers, int32_t rank, int64_t n,
                        const double* in, double* out, int32_t mode, void* stream) {
  if (!peer_buf || !in || !out) return fail(QMLB_ERR_INVALID, "null argument");
  if (n_peers < 2 || n_peers > 8 || rank < 0 || rank >= n_peers || n < 1 || mode < 0 || mode > 2)
    return fail(QMLB_ERR_INVALID, "peer count must be 2..8, rank inside it, n >= 1, mode 0..2");
  PeerReduce R{};
  for (int i = 0; i < n_peers; ++i) {
    if (!peer_buf[i]) return fail(QMLB_ERR_INVALID, "null peer pointer");
    R.buf[i] = const_cast<void*>(peer_buf[i]);
  }
  R.n_peers = n_peers;
  R.rank = rank;
  R.mode = mode;
  R.n = n;
  static const int light = [] {
    const char* v = std::getenv("QMLB_AR_FENCE");
    return (v && std::string(v) == "all") ? 0 : 1;
  }();
  R.pad = light;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  k_allreduce_oneshot<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(R, in, out);
  CUDA_TRY(cudaGetLastError());
  return QMLB_OK;
}

int qmlb_fma_peak(int dtype, double* tflops) {
  if (!tflops) return fail(QMLB_ERR_INVALID, "null argument");
  int dev = 0, sms = SM_COUNT_FALLBACK;
  CUDA_TRY(cudaGetDevice(&dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms * 8, threads = 256, iters = 1 << 15;
  void* buf = nullptr;
  CUDA_TRY(cudaMalloc(&buf, (size_t)grid * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CUDA_TRY(cudaEventRecord(e0));
    if (dtype == QMLB_C128)
      k_fma_peak<double><<<grid, threads>>>(static_cast<double*>(buf), iters);
    else
      k_fma_peak<float><<<grid, threads>>>(static_cast<float*>(buf), iters);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  const double flops = (double)grid * threads * (double)iters * 8.0 * 2.0;
  *tflops = flops / (best * 1e-3) / 1e12;
  return QMLB_OK;
}

}  // extern "C"
