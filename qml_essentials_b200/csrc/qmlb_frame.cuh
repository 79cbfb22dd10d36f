// On-chip frame engine (strategy 3): ONE launch evolves every element of the batch through
// the whole tape with the state resident in shared memory - one CTA per state (several
// small states per CTA), or a thread-block CLUSTER whose CTAs each hold 2^T amplitudes of
// one state in their shared memory and exchange bits through distributed shared memory
// (BASELINE config 4: the 256 x 256 complex128 density matrix of 8 qubits, 1 MiB, lives in
// a cluster of 8 CTAs for all 472 tape operations and is written to HBM once).
// See qmlb_frame_types.h for the step program and qmlb_frame_plan.cu for the planner.
#pragma once

#include <cooperative_groups.h>

#include "qmlb_frame_types.h"
#include "qmlb_stream.cuh"

namespace qmlb {

namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t frame_deposit(uint32_t w, const uint32_t (&piv)[FRAME_R]) {
#pragma unroll
  for (int j = 0; j < FRAME_R; ++j) w = insert0(w, (int)piv[j]);
  return w;
}

// dense 2^K x 2^K on register bits K-1..0 where slot u holds logical local value u ^ c
template <typename T, int K>
__device__ __forceinline__ void frame_matk(RegState<T, FRAME_R>& S, const cx<T>* __restrict__ m,
                                           int c) {
  constexpr int D = 1 << K;
#pragma unroll
  for (int blk = 0; blk < (1 << (FRAME_R - K)); ++blk) {
    T ar[D], ai[D];
#pragma unroll
    for (int u = 0; u < D; ++u) {
      ar[u] = S.re((blk << K) | u);
      ai[u] = S.im((blk << K) | u);
    }
#pragma unroll 1
    for (int v = 0; v < D; ++v) {
      T xr = (T)0, xi = (T)0;
#pragma unroll
      for (int u = 0; u < D; ++u) {
        const cx<T> e = m[((v ^ c) << K) | (u ^ c)];
        xr = fma(e.x, ar[u], xr);
        xr = fma(-e.y, ai[u], xr);
        xi = fma(e.x, ai[u], xi);
        xi = fma(e.y, ar[u], xi);
      }
#pragma unroll
      for (int w = 0; w < D; ++w)
        if (w == v) {
          S.re((blk << K) | w) = xr;
          S.im((blk << K) | w) = xi;
        }
    }
  }
}

// ---- register-group updates -----------------------------------------------------------
// SHAPE: QMLB_FSHAPE_FULL / _REAL (imaginary parts known to be zero) / _XREAL (4x4: only
// v == u and v == u ^ 3, real).

template <typename T, int BIT, int SHAPE>
__device__ __forceinline__ void frame_mat1(RegState<T, FRAME_R>& S, const cx<T>* __restrict__ m) {
  if constexpr (SHAPE == QMLB_FSHAPE_FULL) {
    cx<T> mm[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) mm[i] = m[i];
    reg_mat1<T, FRAME_R, BIT>(S, mm);
  } else {
    T r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = m[i].x;
#pragma unroll
    for (int g = 0; g < (1 << (FRAME_R - 1)); ++g) {
      const int i0 = pair_i0<FRAME_R, BIT>(g), i1 = i0 | (1 << BIT);
      const T ar = S.re(i0), ai = S.im(i0), br = S.re(i1), bi = S.im(i1);
      S.re(i0) = r[0] * ar + r[1] * br;
      S.im(i0) = r[0] * ai + r[1] * bi;
      S.re(i1) = r[2] * ar + r[3] * br;
      S.im(i1) = r[2] * ai + r[3] * bi;
    }
  }
}

template <typename T, int JA, int JB, int SHAPE>
__device__ __forceinline__ void frame_mat2(RegState<T, FRAME_R>& S, const cx<T>* __restrict__ m) {
  static_assert(JA > JB, "canonical order");
  if constexpr (SHAPE == QMLB_FSHAPE_XREAL) {
    // new[v] = d[v] * old[v] + e[v] * old[v ^ 3]
    T d[4], e[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      d[v] = m[v * 5].x;
      e[v] = m[v * 4 + (v ^ 3)].x;
    }
#pragma unroll
    for (int g = 0; g < (1 << (FRAME_R - 2)); ++g) {
      const int t = ((g >> JB) << (JB + 1)) | (g & ((1 << JB) - 1));
      const int i00 = ((t >> JA) << (JA + 1)) | (t & ((1 << JA) - 1));
      const int idx[4] = {i00, i00 | (1 << JB), i00 | (1 << JA), i00 | (1 << JA) | (1 << JB)};
      T ar[4], ai[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ar[u] = S.re(idx[u]);
        ai[u] = S.im(idx[u]);
      }
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        S.re(idx[v]) = fma(d[v], ar[v], e[v] * ar[v ^ 3]);
        S.im(idx[v]) = fma(d[v], ai[v], e[v] * ai[v ^ 3]);
      }
    }
  } else {
#pragma unroll
    for (int g = 0; g < (1 << (FRAME_R - 2)); ++g) {
      const int t = ((g >> JB) << (JB + 1)) | (g & ((1 << JB) - 1));
      const int i00 = ((t >> JA) << (JA + 1)) | (t & ((1 << JA) - 1));
      const int idx[4] = {i00, i00 | (1 << JB), i00 | (1 << JA), i00 | (1 << JA) | (1 << JB)};
      T ar[4], ai[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ar[u] = S.re(idx[u]);
        ai[u] = S.im(idx[u]);
      }
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        T xr = (T)0, xi = (T)0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const cx<T> c = m[v * 4 + u];
          xr = fma(c.x, ar[u], xr);
          xi = fma(c.x, ai[u], xi);
          if constexpr (SHAPE == QMLB_FSHAPE_FULL) {
            xr = fma(-c.y, ai[u], xr);
            xi = fma(c.y, ar[u], xi);
          }
        }
        S.re(idx[v]) = xr;
        S.im(idx[v]) = xi;
      }
    }
  }
}

// 2x2 on register bit TB of the pairs whose control value is 1.  The control value of
// slot i is ctl_base ^ bit i of smask (it never depends on the target bit itself).
template <typename T, int TB>
__device__ __forceinline__ void frame_ctrl1(RegState<T, FRAME_R>& S, const cx<T>* __restrict__ m,
                                            unsigned smask, int ctl_base) {
  cx<T> mm[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) mm[i] = m[i];
#pragma unroll
  for (int g = 0; g < (1 << (FRAME_R - 1)); ++g) {
    const int i0 = pair_i0<FRAME_R, TB>(g), i1 = i0 | (1 << TB);
    const bool on = (((smask >> i0) & 1u) ^ (unsigned)ctl_base) != 0u;
    const T ar = S.re(i0), ai = S.im(i0), br = S.re(i1), bi = S.im(i1);
    const T xr = mm[0].x * ar - mm[0].y * ai + mm[1].x * br - mm[1].y * bi;
    const T xi = mm[0].x * ai + mm[0].y * ar + mm[1].x * bi + mm[1].y * br;
    const T yr = mm[2].x * ar - mm[2].y * ai + mm[3].x * br - mm[3].y * bi;
    const T yi = mm[2].x * ai + mm[2].y * ar + mm[3].x * bi + mm[3].y * br;
    S.re(i0) = on ? xr : ar;
    S.im(i0) = on ? xi : ai;
    S.re(i1) = on ? yr : br;
    S.im(i1) = on ? yi : bi;
  }
}


// ---- items of a SUBPASS ------------------------------------------------------------------
// The parity rows of the register bits tell which logical value slot 0 of an item holds
// (c_j = parity(base & rloc_j) ^ parity(rank & rout_j)).  eoff is linear in the slot number,
// so starting the item at base ^ eoff[c] instead makes slot v hold logical value v exactly:
// every matrix is then read with plain compile-time entry offsets.
__device__ __forceinline__ uint32_t frame_item_base(const FrameStep& st, uint32_t it,
                                                    const uint32_t (&piv)[FRAME_R], unsigned rank) {
  uint32_t base = frame_deposit(it, piv);
  int c = 0;
#pragma unroll
  for (int j = 0; j < FRAME_R; ++j)
    c |= ((__popc(base & st.par[j].rloc) ^ __popc(rank & st.par[j].rout)) & 1) << j;
  return base ^ st.eoff[c];
}

template <typename T>
__device__ __forceinline__ void frame_load(RegState<T, FRAME_R>& S, const cx<T>* tile,
                                           uint32_t base, const FrameStep& st) {
#pragma unroll
  for (int v = 0; v < FRAME_D; ++v) {
    const cx<T> a = tile[base ^ st.eoff[v]];
    S.re(v) = a.x;
    S.im(v) = a.y;
  }
}

template <typename T>
__device__ __forceinline__ void frame_store(RegState<T, FRAME_R>& S, cx<T>* tile, uint32_t base,
                                            const FrameStep& st) {
#pragma unroll
  for (int v = 0; v < FRAME_D; ++v) tile[base ^ st.eoff[v]] = mk<T>(S.re(v), S.im(v));
}

// eoff is GF(2)-linear in the slot number, so six words in registers - the four masks and the
// two pair sums - give every slot address with ONE three-input XOR: no eoff word is read
// from shared memory inside the item loop (the compiler had to re-read them after every
// tile store, which may alias the step record).
struct FrameLin {
  uint32_t lo[4], hi[4];  // lo[v & 3] ^ hi[v >> 2] = eoff[v]
};
__device__ __forceinline__ FrameLin frame_lin(const FrameStep& st) {
  FrameLin L;
  L.lo[0] = 0u, L.lo[1] = st.eoff[1], L.lo[2] = st.eoff[2], L.lo[3] = st.eoff[3];
  L.hi[0] = 0u, L.hi[1] = st.eoff[4], L.hi[2] = st.eoff[8], L.hi[3] = st.eoff[12];
  return L;
}
template <typename T>
__device__ __forceinline__ void frame_load_lin(RegState<T, FRAME_R>& S, const cx<T>* tile,
                                               uint32_t base, const FrameLin& L) {
#pragma unroll
  for (int v = 0; v < FRAME_D; ++v) {
    const cx<T> a = tile[base ^ L.lo[v & 3] ^ L.hi[v >> 2]];
    S.re(v) = a.x;
    S.im(v) = a.y;
  }
}
template <typename T>
__device__ __forceinline__ void frame_store_lin(RegState<T, FRAME_R>& S, cx<T>* tile,
                                                uint32_t base, const FrameLin& L) {
#pragma unroll
  for (int v = 0; v < FRAME_D; ++v)
    tile[base ^ L.lo[v & 3] ^ L.hi[v >> 2]] = mk<T>(S.re(v), S.im(v));
}

// fast path: at most one 4x4 on register pair (1,0) (shape SA) and one on (3,2) (shape SB);
// -1 = absent.  Straight-line code: no op decoding, no dispatch inside the item loop.
template <typename T, int SA, int SB>
__device__ __noinline__ void frame_items_d2(cx<T>* tile, const cx<T>* mats, const FrameStep& st,
                                            unsigned rank, uint32_t tlane, uint32_t tsize,
                                            uint32_t n_items) {
  uint32_t piv[FRAME_R];
#pragma unroll
  for (int j = 0; j < FRAME_R; ++j) piv[j] = st.pivots[j];
  const cx<T>* ma = mats + st.foff[0];
  const cx<T>* mb = mats + st.foff[1];
  for (uint32_t it = tlane; it < n_items; it += tsize) {
    const uint32_t base = frame_item_base(st, it, piv, rank);
    RegState<T, FRAME_R> S;
    frame_load<T>(S, tile, base, st);
    if constexpr (SA >= 0) frame_mat2<T, 1, 0, SA>(S, ma);
    if constexpr (SB >= 0) frame_mat2<T, 3, 2, SB>(S, mb);
    frame_store<T>(S, tile, base, st);
  }
}

// The same for exactly TWO items per thread (256 threads on a 2^13-amplitude tile, up to
// 255 registers): both items are loaded first, so the shared-memory latency of the second
// hides under the arithmetic of the first and the stores of the first under the arithmetic
// of the second - only half of the shared-memory time of the step stays exposed.
template <typename T, int SA, int SB>
__device__ __noinline__ void frame_items_d2_x2(cx<T>* tile, const cx<T>* mats,
                                               const FrameStep& st, unsigned rank, uint32_t tlane,
                                               uint32_t tsize) {
  uint32_t piv[FRAME_R];
#pragma unroll
  for (int j = 0; j < FRAME_R; ++j) piv[j] = st.pivots[j];
  const cx<T>* ma = mats + st.foff[0];
  const cx<T>* mb = mats + st.foff[1];
  const uint32_t base0 = frame_item_base(st, tlane, piv, rank);
  const uint32_t base1 = frame_item_base(st, tlane + tsize, piv, rank);
  RegState<T, FRAME_R> S0, S1;
  frame_load<T>(S0, tile, base0, st);
  frame_load<T>(S1, tile, base1, st);
  if constexpr (SA >= 0) frame_mat2<T, 1, 0, SA>(S0, ma);
  if constexpr (SB >= 0) frame_mat2<T, 3, 2, SB>(S0, mb);
  frame_store<T>(S0, tile, base0, st);
  if constexpr (SA >= 0) frame_mat2<T, 1, 0, SA>(S1, ma);
  if constexpr (SB >= 0) frame_mat2<T, 3, 2, SB>(S1, mb);
  frame_store<T>(S1, tile, base1, st);
}

// fast path: only 2x2 ops, at most one per register bit (MASK)
template <typename T, int MASK, bool REAL>
__device__ __noinline__ void frame_items_m1(cx<T>* tile, const cx<T>* mats, const FrameStep& st,
                                            unsigned rank, uint32_t tlane, uint32_t tsize,
                                            uint32_t n_items) {
  constexpr int SH = REAL ? QMLB_FSHAPE_REAL : QMLB_FSHAPE_FULL;
  uint32_t piv[FRAME_R];
#pragma unroll
  for (int j = 0; j < FRAME_R; ++j) piv[j] = st.pivots[j];
  const FrameLin L = frame_lin(st);
  for (uint32_t it = tlane; it < n_items; it += tsize) {
    const uint32_t base = frame_item_base(st, it, piv, rank);
    RegState<T, FRAME_R> S;
    frame_load_lin<T>(S, tile, base, L);
    if constexpr (MASK & 1) frame_mat1<T, 0, SH>(S, mats + st.foff[0]);
    if constexpr (MASK & 2) frame_mat1<T, 1, SH>(S, mats + st.foff[1]);
    if constexpr (MASK & 4) frame_mat1<T, 2, SH>(S, mats + st.foff[2]);
    if constexpr (MASK & 8) frame_mat1<T, 3, SH>(S, mats + st.foff[3]);
    frame_store_lin<T>(S, tile, base, L);
  }
}

template <typename T, bool X2, int I = 0>
__device__ __forceinline__ void frame_dispatch_d2(int code, cx<T>* tile, const cx<T>* mats,
                                                  const FrameStep& st, unsigned rank,
                                                  uint32_t tlane, uint32_t tsize, uint32_t n) {
  if constexpr (I < 16) {
    if (code == I) {
      if constexpr (I > 0) {
        if (X2 && n == 2 * tsize)
          frame_items_d2_x2<T, (I >> 2) - 1, (I & 3) - 1>(tile, mats, st, rank, tlane, tsize);
        else
          frame_items_d2<T, (I >> 2) - 1, (I & 3) - 1>(tile, mats, st, rank, tlane, tsize, n);
      }
    } else {
      frame_dispatch_d2<T, X2, I + 1>(code, tile, mats, st, rank, tlane, tsize, n);
    }
  }
}

template <typename T, int I = 1>
__device__ __forceinline__ void frame_dispatch_m1(int code, cx<T>* tile, const cx<T>* mats,
                                                  const FrameStep& st, unsigned rank,
                                                  uint32_t tlane, uint32_t tsize, uint32_t n) {
  if constexpr (I < 32) {
    if (code == I) {
      if constexpr ((I & 15) != 0)
        frame_items_m1<T, (I & 15), (I >> 4) != 0>(tile, mats, st, rank, tlane, tsize, n);
    } else {
      frame_dispatch_m1<T, I + 1>(code, tile, mats, st, rank, tlane, tsize, n);
    }
  }
}


// generic interpreter of a SUBPASS: any mix of 2x2 / 4x4 / dense 3-4 bit / controlled /
// diagonal ops (steps the straight-line fast paths do not cover)
template <typename T, bool HEAVY>
__device__ __noinline__ void frame_items_generic(cx<T>* tile, const cx<T>* mats,
                                                 const FrameStep& st, unsigned rank,
                                                 uint32_t tlane, uint32_t tsize,
                                                 uint32_t n_items) {
        uint32_t piv[FRAME_R];
#pragma unroll
  for (int j = 0; j < FRAME_R; ++j) piv[j] = st.pivots[j];
  for (uint32_t it = tlane; it < n_items; it += tsize) {
    const uint32_t base = frame_item_base(st, it, piv, rank);
    RegState<T, FRAME_R> S;
    frame_load<T>(S, tile, base, st);
    auto parity_at = [&](int pi) -> int {
      return (__popc(base & st.par[pi].rloc) ^ __popc(rank & st.par[pi].rout)) & 1;
    };
#pragma unroll 1
    for (int o = 0; o < st.n_ops; ++o) {
      const FrameOp fo = st.ops[o];
      const cx<T>* m = mats + fo.smem_off;
      switch (fo.code) {
        case QMLB_FOP_MAT1:
          dispatch1<T, FRAME_R>(fo.j0, [&](auto B) {
            constexpr int BIT = decltype(B)::value;
            if (fo.shape != QMLB_FSHAPE_FULL)
              frame_mat1<T, BIT, QMLB_FSHAPE_REAL>(S, m);
            else
              frame_mat1<T, BIT, QMLB_FSHAPE_FULL>(S, m);
          });
          break;
        case QMLB_FOP_MAT2: {
          auto on_pair = [&](auto JA, auto JB) {
            constexpr int A_ = decltype(JA)::value, B_ = decltype(JB)::value;
            if constexpr (A_ > B_) {
              if (fo.shape == QMLB_FSHAPE_XREAL)
                frame_mat2<T, A_, B_, QMLB_FSHAPE_XREAL>(S, m);
              else if (fo.shape == QMLB_FSHAPE_REAL)
                frame_mat2<T, A_, B_, QMLB_FSHAPE_REAL>(S, m);
              else
                frame_mat2<T, A_, B_, QMLB_FSHAPE_FULL>(S, m);
            }
          };
          // the planner seats 2-bit ops on the register pairs (3,2) / (1,0)
          if (fo.j0 == 3 && fo.j1 == 2)
            on_pair(std::integral_constant<int, 3>{}, std::integral_constant<int, 2>{});
          else if (fo.j0 == 1 && fo.j1 == 0)
            on_pair(std::integral_constant<int, 1>{}, std::integral_constant<int, 0>{});
          else
            dispatch2<T, FRAME_R>(fo.j0, fo.j1, on_pair);
          break;
        }
        case QMLB_FOP_MATK:
          if constexpr (HEAVY) {
            if (fo.k == 3)
              frame_matk<T, 3>(S, m, 0);
            else
              frame_matk<T, 4>(S, m, 0);
          }
          break;
        case QMLB_FOP_CTRL1: {
          const int ctl = parity_at(fo.j1);
          const unsigned sm = st.par[fo.j1].smask;
          dispatch1<T, FRAME_R>(fo.j0, [&](auto B) {
            constexpr int BIT = decltype(B)::value;
            frame_ctrl1<T, BIT>(S, m, sm, ctl);
          });
          break;
        }
        case QMLB_FOP_DIAG: {
          const uint8_t* idx = reinterpret_cast<const uint8_t*>(&st.ops[o + 1]);
          int lb = 0;  // local value of slot 0
          for (int a = 0; a < fo.k; ++a) lb |= parity_at(idx[a]) << (fo.k - 1 - a);
#pragma unroll
          for (int v = 0; v < FRAME_D; ++v) {
            int loc = lb;
            for (int a = 0; a < fo.k; ++a)
              loc ^= (int)((st.par[idx[a]].smask >> v) & 1u) << (fo.k - 1 - a);
            const cx<T> d = m[loc];
            const T r = S.re(v), q = S.im(v);
            S.re(v) = d.x * r - d.y * q;
            S.im(v) = d.x * q + d.y * r;
          }
          ++o;
          break;
        }
      }
    }

    frame_store<T>(S, tile, base, st);
  }
}

template <typename T, bool HEAVY, bool X2 = false>
__device__ __forceinline__ void frame_subpass(cx<T>* tile, const cx<T>* mats, const FrameStep& st,
                                              unsigned outer, uint32_t tlane, uint32_t tsize,
                                              uint32_t n_items) {
  if (!HEAVY && st.fast >= 64)
    frame_dispatch_m1<T>(st.fast - 64, tile, mats, st, outer, tlane, tsize, n_items);
  else if (!HEAVY && st.fast >= 16)
    frame_dispatch_d2<T, X2>(st.fast - 16, tile, mats, st, outer, tlane, tsize, n_items);
  else
    frame_items_generic<T, HEAVY>(tile, mats, st, outer, tlane, tsize, n_items);
}

// RELAYOUT gather: destination d of this thread <- source (rank, index) through the GF(2)
// map; PER amplitudes per thread held in registers across the cluster barrier
// NB > 0: the tile is swizzled (qmlb_frame_ptm.cuh) - the tables then hold swizzled source
// indices and the destination index is swizzled here
// htab (the step's lane table, 2^LB entries without bits below LB, LB = log2(elements per
// 128-byte wavefront)): the thread copies d = d_std ^ htab[d_std & (2^LB - 1)] - the
// planner's choice of lane directions that keeps gather AND store free of bank conflicts
template <typename V, int PER, int NB = 0, typename Cluster>
__device__ __forceinline__ void frame_relayout_v(V* tv, Cluster& cluster, bool clustered,
                                                 unsigned rank, int Tb, int team_bits, int tlane,
                                                 uint32_t cmine, const uint32_t* tab_lo,
                                                 const uint32_t* tab_hi,
                                                 const uint32_t* htab = nullptr) {
  constexpr int hi_shift = 8;
  constexpr int LB = sizeof(V) == 16 ? 3 : (sizeof(V) == 8 ? 4 : 5);
  const uint32_t tile_mask = (1u << Tb) - 1u;
  if (htab == nullptr) {
    // plain form (consecutive destinations per warp): the Pauli-basis kernel, whose swizzled
    // tile already spreads the gather - there the lane table measured 8 % slower
    V hold[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const uint32_t d = (uint32_t)tlane + ((uint32_t)k << team_bits);
      const uint32_t src = cmine ^ tab_lo[d & 255u] ^ tab_hi[d >> hi_shift];
      const uint32_t r = src >> Tb, loc = src & tile_mask;
      const V* from = tv;
      if (clustered && r != rank) from = cluster.map_shared_rank(tv, r);
      hold[k] = from[loc];
    }
    if (clustered)
      cluster.sync();
    else
      __syncthreads();
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      uint32_t d = (uint32_t)tlane + ((uint32_t)k << team_bits);
      if constexpr (NB > 0)
        d ^= ((d >> NB) ^ (d >> (2 * NB)) ^ (d >> (3 * NB))) & ((1u << NB) - 1u);
      tv[d] = hold[k];
    }
    return;
  }
  // Every map here is GF(2)-linear, and tlane + (k << team_bits) = tlane ^ (k << team_bits):
  // the thread's part (lane table included) is folded into per-thread constants once, the
  // k part is uniform - one XOR per element and side.  team_bits >= LB: the low LB bits of
  // every index of this thread are those of its lane (small tiles look the table up per
  // element instead).
  const bool hoist = htab != nullptr && team_bits >= LB;
  const uint32_t base = (uint32_t)tlane ^ (hoist ? htab[tlane & ((1 << LB) - 1)] : 0u);
  auto swz = [](uint32_t d) {
    if constexpr (NB > 0)
      d ^= ((d >> NB) ^ (d >> (2 * NB)) ^ (d >> (3 * NB))) & ((1u << NB) - 1u);
    return d;
  };
  const uint32_t base_sw = swz(base);
  const uint32_t src0 = cmine ^ tab_lo[base & 255u] ^ tab_hi[base >> hi_shift];
  V hold[PER];
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const uint32_t dk = (uint32_t)k << team_bits;  // uniform
    uint32_t src = src0 ^ tab_lo[dk & 255u] ^ tab_hi[dk >> hi_shift];
    if (htab != nullptr && !hoist) {
      const uint32_t d = ((uint32_t)tlane ^ dk);
      const uint32_t dd = d ^ htab[d & ((1u << LB) - 1u)];
      src = cmine ^ tab_lo[dd & 255u] ^ tab_hi[dd >> hi_shift];
    }
    const uint32_t r = src >> Tb, loc = src & tile_mask;
    const V* from = tv;
    if (clustered && r != rank) from = cluster.map_shared_rank(tv, r);
    hold[k] = from[loc];
  }
  if (clustered)
    cluster.sync();
  else
    __syncthreads();
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const uint32_t dk = (uint32_t)k << team_bits;
    uint32_t d = base_sw ^ swz(dk);
    if (htab != nullptr && !hoist) {
      const uint32_t d0 = ((uint32_t)tlane ^ dk);
      d = swz(d0 ^ htab[d0 & ((1u << LB) - 1u)]);
    }
    tv[d] = hold[k];
  }
}

template <typename T, int PER, typename Cluster>
__device__ __forceinline__ void frame_relayout(cx<T>* tile, Cluster& cluster, bool clustered,
                                               unsigned rank, int Tb, int team_bits, int tlane,
                                               uint32_t cmine, const uint32_t* tab_lo,
                                               const uint32_t* tab_hi,
                                               const uint32_t* htab = nullptr) {
  using V = typename std::conditional<sizeof(T) == 8, double2, float2>::type;  // one access
  frame_relayout_v<V, PER>(reinterpret_cast<V*>(tile), cluster, clustered, rank, Tb, team_bits,
                           tlane, cmine, tab_lo, tab_hi, htab);
}

// HEAVY = the program holds a dense op on 3 or 4 bits (rare: 2-qubit channels, CCX as a
// matrix); the lean variant keeps the register pressure of the common steps low.
// WIDE = one CTA per SM even at 256 threads (255 registers per thread: two items per thread
// without spills when the tile fills the CTA's shared memory anyway).
template <typename T, int THREADS, bool HEAVY, bool WIDE = false>
__global__ void __launch_bounds__(THREADS, (THREADS == 512 || WIDE) ? 1 : 2)
    k_frame(DevProg P, RunArgs A, const FrameProg F, const cx<T>* __restrict__ premats,
            void* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char fsm[];
  const int Tb = F.tile_bits;
  const uint32_t tile_n = 1u << Tb;
  const int teams = F.teams;
  const int tsize = 1 << F.team_bits;
  const int team = threadIdx.x >> F.team_bits;
  const int tlane = threadIdx.x & (tsize - 1);
  // [tiles | matrices | 2 step records | relayout tables | reduction scratch]
  const uint32_t pitch = F.tile_pitch ? (uint32_t)F.tile_pitch : tile_n;
  const bool direct = F.mat_resident == 2;  // matrices read from global memory
  cx<T>* tiles = reinterpret_cast<cx<T>*>(fsm);
  cx<T>* mats_all = tiles + (((size_t)teams * pitch + 63) & ~size_t(63));
  FrameStep* sstep =
      reinterpret_cast<FrameStep*>(mats_all + (direct ? 0 : (size_t)teams * F.mat_cap));
  uint32_t* tab_lo = reinterpret_cast<uint32_t*>(sstep + (F.steps_resident ? F.n_steps : 2));
  uint32_t* tab_hi = tab_lo + 256;
  double* red = reinterpret_cast<double*>(tab_hi + 64);
  cx<T>* tile = tiles + (size_t)team * pitch;
  const cx<T>* mats = mats_all + (size_t)team * F.mat_cap;

  const bool clustered = F.outer_bits > 0;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned csize = 1u << F.outer_bits;
  const unsigned rank = clustered ? cluster.block_rank() : 0u;
  const int64_t cluster_id = blockIdx.x >> F.outer_bits;
  const int64_t n_clusters = gridDim.x >> F.outer_bits;
  auto sync_all = [&]() {
    if (clustered)
      cluster.sync();
    else
      __syncthreads();
  };

  const int64_t per_round = n_clusters * teams;
  const int64_t rounds = (A.batch + per_round - 1) / per_round;
  const uint32_t n_items = 1u << (Tb - FRAME_R);

  if (F.steps_resident) {  // the whole step program, once per CTA
    const int4* src = reinterpret_cast<const int4*>(F.steps);
    int4* dst = reinterpret_cast<int4*>(sstep);
    for (int i = threadIdx.x; i < F.n_steps * 64; i += THREADS) dst[i] = src[i];
  }

  for (int64_t rd = 0; rd < rounds; ++rd) {
    const int64_t bl = rd * per_round + cluster_id * teams + team;
    const bool valid = bl < A.batch;
    const cx<T>* prow = premats + (size_t)(valid ? bl : 0) * F.premat_row;
    if (direct) mats = prow;

    // |0..0>: the frame is linear, so logical index 0 sits at physical index 0
    for (uint32_t i = tlane; i < tile_n; i += tsize)
      tile[i] = mk<T>((i == 0 && rank == 0) ? (T)1 : (T)0, (T)0);
    if (F.mat_resident == 1) {  // every matrix of this element, once
      cx<T>* mdst = mats_all + (size_t)team * F.mat_cap;
      for (int i = tlane; i < F.premat_row; i += tsize) mdst[i] = prow[i];
    }
    if (!F.steps_resident)
      for (int i = threadIdx.x; i < 256; i += THREADS)
        reinterpret_cast<uint32_t*>(&sstep[0])[i] =
            reinterpret_cast<const uint32_t*>(&F.steps[0])[i];
    sync_all();

    for (int si = 0; si < F.n_steps; ++si) {
      const FrameStep& st = sstep[F.steps_resident ? si : (si & 1)];
      if (!F.steps_resident && si + 1 < F.n_steps)
        for (int i = threadIdx.x; i < 256; i += THREADS)
          reinterpret_cast<uint32_t*>(&sstep[(si + 1) & 1])[i] =
              reinterpret_cast<const uint32_t*>(&F.steps[si + 1])[i];

      if (st.kind == QMLB_FSTEP_RELAYOUT) {
        // source index of destination d: XOR of qcol over the set bits of (rank, d)
        for (int i = threadIdx.x; i < 256 + 64; i += THREADS) {
          uint32_t acc = 0;
          if (i < 256) {
            for (int b = 0; b < 8; ++b)
              if (i >> b & 1) acc ^= (uint32_t)st.qcol[b];
            tab_lo[i] = acc;
          } else {
            const int h = i - 256;
            for (int b = 0; b < 6; ++b)
              if ((h >> b & 1) && 8 + b < Tb) acc ^= (uint32_t)st.qcol[8 + b];
            tab_hi[h] = acc;
          }
        }
        uint32_t cmine = 0;
        for (int g = 0; g < F.outer_bits; ++g)
          if (rank >> g & 1) cmine ^= (uint32_t)st.qcol[Tb + g];
        const bool across = clustered && st.mat_entries == 0;  // else a tile-local shuffle
        const uint32_t* htab = reinterpret_cast<const uint32_t*>(st.ops);
        if (across)
          cluster.sync();  // tables visible; every CTA of the cluster finished its previous step
        else
          __syncthreads();
        const int per = (int)(tile_n >> F.team_bits);
        if (per == 16) {
          frame_relayout<T, 16>(tile, cluster, across, rank, Tb, F.team_bits, tlane, cmine,
                                tab_lo, tab_hi, htab);
        } else if (per == 32) {
          frame_relayout<T, 32>(tile, cluster, across, rank, Tb, F.team_bits, tlane, cmine,
                                tab_lo, tab_hi, htab);
        } else {
          if constexpr (WIDE && sizeof(T) == 4)  // complex64, 2^14 amplitudes, 256 threads
            frame_relayout<T, 64>(tile, cluster, across, rank, Tb, F.team_bits, tlane, cmine,
                                  tab_lo, tab_hi, htab);
        }
        __syncthreads();  // the next reader of this tile is this CTA (or a later relayout)
        continue;
      }

      // ---- SUBPASS ------------------------------------------------------------------------
      if (!F.mat_resident) {  // stage the matrices of this step
        if (valid) {
          cx<T>* mdst = mats_all + (size_t)team * F.mat_cap;
          for (int o = 0; o < st.n_ops; ++o) {
            const FrameOp fo = st.ops[o];
            const int n = fo.code == QMLB_FOP_DIAG    ? (1 << fo.k)
                          : fo.code == QMLB_FOP_CTRL1 ? 4  // the 2x2 alone, not 4^k
                                                      : (1 << (2 * fo.k));
            for (int e = tlane; e < n; e += tsize) mdst[fo.smem_off + e] = prow[fo.premat_off + e];
            if (fo.code == QMLB_FOP_DIAG) ++o;
          }
        }
        __syncthreads();
      }

      if (valid) frame_subpass<T, HEAVY, WIDE>(tile, mats, st, rank, tlane, tsize, n_items);
      __syncthreads();
    }

    // ---- result: the tile is in index order (logical index = rank << T | i) -----------------
    const int nq = F.n_qubits;
    if (F.out_mode == 0) {
      if (valid) {
        cx<T>* o = reinterpret_cast<cx<T>*>(out) + ((size_t)bl << F.n_bits) + ((size_t)rank << Tb);
        for (uint32_t i = tlane; i < tile_n; i += tsize) o[i] = tile[i];
      }
    } else if (F.out_mode == 1) {
      if (valid) {
        T* o = reinterpret_cast<T*>(out) + ((size_t)bl << nq);
        if (!F.density) {
          for (uint32_t i = tlane; i < tile_n; i += tsize) {
            const cx<T> a = tile[i];
            o[((size_t)rank << Tb) + i] = a.x * a.x + a.y * a.y;
          }
        } else {
          // rho[k][k] at index k * 2^n + k: this CTA owns the k whose top bits equal its rank
          const uint32_t kper = 1u << (nq - F.outer_bits);
          for (uint32_t kk = tlane; kk < kper; kk += tsize) {
            const uint32_t ket = (rank << (nq - F.outer_bits)) | kk;
            o[ket] = tile[((size_t)kk << nq) | ket].x;
          }
        }
      }
    } else {
      // Z-string expectation values, double accumulation in a fixed order: thread partial ->
      // xor-shuffle tree -> warp partials -> (cluster) rank order
      const uint32_t cnt = F.density ? (1u << (nq - F.outer_bits)) : tile_n;
      for (int j = 0; j < F.n_obs; ++j) {
        const uint32_t zm = (uint32_t)P.obs[j].zmask;
        double acc = 0.0;
        if (valid) {
          for (uint32_t i = tlane; i < cnt; i += tsize) {
            uint32_t idx;
            double pr;
            if (F.density) {
              idx = (rank << (nq - F.outer_bits)) | i;
              pr = (double)tile[((size_t)i << nq) | idx].x;
            } else {
              idx = (rank << Tb) | i;
              const cx<T> a = tile[i];
              pr = (double)a.x * (double)a.x + (double)a.y * (double)a.y;
            }
            acc += (__popc(idx & zm) & 1) ? -pr : pr;
          }
        }
        const int seg = tsize < 32 ? tsize : 32;
        for (int off = seg >> 1; off > 0; off >>= 1)
          acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (tsize <= 32) {
          if (tlane == 0 && valid)
            reinterpret_cast<T*>(out)[(size_t)bl * F.n_obs + j] = (T)acc;
        } else {
          // a team of several warps: warp partials, summed in warp order by the team's lane 0
          if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
          __syncthreads();
          double s = 0.0;
          if (tlane == 0) {
            const int wpt = tsize >> 5;
            for (int w = 0; w < wpt; ++w) s += red[team * wpt + w];
          }
          __syncthreads();
          if (!clustered) {
            if (tlane == 0 && valid)
              reinterpret_cast<T*>(out)[(size_t)bl * F.n_obs + j] = (T)s;
          } else {
            if (threadIdx.x == 0) red[32] = s;
            cluster.sync();
            if (rank == 0 && threadIdx.x == 0 && valid) {
              double tot = 0.0;
              for (unsigned r = 0; r < csize; ++r) tot += *cluster.map_shared_rank(&red[32], r);
              reinterpret_cast<T*>(out)[(size_t)bl * F.n_obs + j] = (T)tot;
            }
            cluster.sync();
          }
        }
      }
    }
    sync_all();
  }
}

}  // namespace qmlb
