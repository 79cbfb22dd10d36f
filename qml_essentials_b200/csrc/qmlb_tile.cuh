// Shared-memory tile interpreter: applies a run of program ops to a 2^kt-amplitude
// tile held in shared memory by one team (a warp or the whole CTA).
//
//   * strategy 1 (state fits on chip): the tile IS the state; it starts as |0..0>,
//     every op of the circuit is applied in place and the state is written once.
//   * strategy 2 (streamed): each pass gathers tiles from HBM (the low `m` bits are
//     contiguous -> coalesced 128-bit accesses, the other tile bits are strided:
//     the shared-memory transpose for high-order qubits), applies every op that
//     fits the tile's bit set, and scatters the tile back.
#pragma once

#include "qmlb_device.cuh"
#include "qmlb_tile_types.h"

namespace qmlb {

template <bool WARP_TEAM>
__device__ __forceinline__ void team_sync() {
  if (WARP_TEAM)
    __syncwarp();
  else
    __syncthreads();
}

// positions sorted ascending -> base index of group g with zeros at those positions
template <int K>
__device__ __forceinline__ uint32_t group_base(uint32_t g, const int (&sorted)[K]) {
#pragma unroll
  for (int j = 0; j < K; ++j) g = insert0(g, sorted[j]);
  return g;
}

template <int K>
__device__ __forceinline__ void sort_bits(const int32_t* bits, int (&sorted)[K]) {
#pragma unroll
  for (int j = 0; j < K; ++j) sorted[j] = bits[j];
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j + 1 < K - i; ++j)
      if (sorted[j] > sorted[j + 1]) {
        int t = sorted[j];
        sorted[j] = sorted[j + 1];
        sorted[j + 1] = t;
      }
}

// offset of local value v (bits[0] = MSB of v)
template <int K>
__device__ __forceinline__ uint32_t value_offset(int v, const int32_t* bits) {
  uint32_t o = 0;
#pragma unroll
  for (int j = 0; j < K; ++j) o |= (uint32_t)((v >> (K - 1 - j)) & 1) << bits[j];
  return o;
}

// dense 2^K x 2^K matrix `m` (shared memory, row-major) on tile `s`
template <typename T, int K>
__device__ __forceinline__ void tile_mat(cx<T>* s, int kt, const int32_t* bits,
                                         const cx<T>* m, int tlane, int tsize) {
  constexpr int D = 1 << K;
  int sorted[K];
  sort_bits<K>(bits, sorted);
  uint32_t off[D];
#pragma unroll
  for (int v = 0; v < D; ++v) off[v] = value_offset<K>(v, bits);
  const uint32_t groups = 1u << (kt - K);
  for (uint32_t g = tlane; g < groups; g += tsize) {
    uint32_t base = group_base<K>(g, sorted);
    cx<T> a[D];
#pragma unroll
    for (int v = 0; v < D; ++v) a[v] = s[base | off[v]];
    if (K <= 2) {
#pragma unroll
      for (int v = 0; v < D; ++v) {
        cx<T> acc = mk<T>((T)0, (T)0);
#pragma unroll
        for (int u = 0; u < D; ++u) {
          cx<T> c = m[v * D + u];
          if (c.x != (T)0 || c.y != (T)0) cfma(acc, c, a[u]);  // uniform: skips zeros
        }
        s[base | off[v]] = acc;
      }
    } else {
      // rows as a rolled loop: keeps register pressure (and code size) down for 8x8 / 16x16
#pragma unroll 1
      for (int v = 0; v < D; ++v) {
        cx<T> acc = mk<T>((T)0, (T)0);
#pragma unroll
        for (int u = 0; u < D; ++u) {
          cx<T> c = m[v * D + u];
          if (c.x != (T)0 || c.y != (T)0) cfma(acc, c, a[u]);
        }
        s[base | value_offset<K>(v, bits)] = acc;
      }
    }
  }
}

// 2x2 `m` on bits[1] where bits[0] is set
template <typename T>
__device__ __forceinline__ void tile_ctrl1(cx<T>* s, int kt, const int32_t* bits,
                                           const cx<T>* m, int tlane, int tsize) {
  int sorted[2];
  sort_bits<2>(bits, sorted);
  const uint32_t cb = 1u << bits[0], tb = 1u << bits[1];
  const cx<T> m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
  const uint32_t groups = 1u << (kt - 2);
  for (uint32_t g = tlane; g < groups; g += tsize) {
    uint32_t i0 = group_base<2>(g, sorted) | cb;
    uint32_t i1 = i0 | tb;
    cx<T> a0 = s[i0], a1 = s[i1];
    cx<T> r0 = cmul(m00, a0);
    cfma(r0, m01, a1);
    cx<T> r1 = cmul(m10, a0);
    cfma(r1, m11, a1);
    s[i0] = r0;
    s[i1] = r1;
  }
}

template <typename T, int K>
__device__ __forceinline__ void tile_perm(cx<T>* s, int kt, const int32_t* bits,
                                          const double* perm, int tlane, int tsize) {
  constexpr int D = 1 << K;
  int sorted[K];
  sort_bits<K>(bits, sorted);
  uint32_t off[D];
  int p[D];
#pragma unroll
  for (int v = 0; v < D; ++v) {
    off[v] = value_offset<K>(v, bits);
    p[v] = (int)perm[v];
  }
  const uint32_t groups = 1u << (kt - K);
  for (uint32_t g = tlane; g < groups; g += tsize) {
    uint32_t base = group_base<K>(g, sorted);
    cx<T> a[D];
#pragma unroll
    for (int v = 0; v < D; ++v) a[v] = s[base | off[v]];
#pragma unroll
    for (int v = 0; v < D; ++v) {
      // new[v] = old[p[v]] ; p[v] is uniform, select without dynamic register indexing
      cx<T> r = a[0];
#pragma unroll
      for (int u = 1; u < D; ++u)
        if (p[v] == u) r = a[u];
      if (p[v] != v) s[base | off[v]] = r;
    }
  }
}

// diagonal d (2^k entries, shared memory); bits are GLOBAL positions, gidx maps a
// tile-local index to the global state index
template <typename T, typename G>
__device__ __forceinline__ void tile_diag(cx<T>* s, int kt, int k, const int32_t* bits,
                                          const cx<T>* d, int tlane, int tsize, G gidx) {
  const uint32_t n = 1u << kt;
  for (uint32_t i = tlane; i < n; i += tsize) {
    uint64_t gi = gidx(i);
    int v = 0;
    for (int j = 0; j < k; ++j) v |= (int)((gi >> bits[j]) & 1ull) << (k - 1 - j);
    s[i] = cmul(d[v], s[i]);
  }
}

template <typename T, bool WARP_TEAM, typename G>
__device__ __forceinline__ void tile_apply_op(const DevProg& P, const qmlb_op& op,
                                              cx<T>* st, const cx<T>* mat, int kt,
                                              int tlane, int tsize, G gidx) {
  switch (op.kind) {
    case QMLB_OP_MAT:
      switch (op.k) {
        case 1: tile_mat<T, 1>(st, kt, op.bits, mat, tlane, tsize); break;
        case 2: tile_mat<T, 2>(st, kt, op.bits, mat, tlane, tsize); break;
        case 3: tile_mat<T, 3>(st, kt, op.bits, mat, tlane, tsize); break;
        case 4: tile_mat<T, 4>(st, kt, op.bits, mat, tlane, tsize); break;
      }
      break;
    case QMLB_OP_CTRL1: tile_ctrl1<T>(st, kt, op.bits, mat, tlane, tsize); break;
    case QMLB_OP_PERM:
      switch (op.k) {
        case 1: tile_perm<T, 1>(st, kt, op.bits, P.consts + op.aux, tlane, tsize); break;
        case 2: tile_perm<T, 2>(st, kt, op.bits, P.consts + op.aux, tlane, tsize); break;
        case 3: tile_perm<T, 3>(st, kt, op.bits, P.consts + op.aux, tlane, tsize); break;
        case 4: tile_perm<T, 4>(st, kt, op.bits, P.consts + op.aux, tlane, tsize); break;
      }
      break;
    case QMLB_OP_DIAG: tile_diag<T>(st, kt, op.k, op.bits, mat, tlane, tsize, gidx); break;
  }
}

// One pass over every tile of every element of the launch.
template <typename T, bool WARP_TEAM>
__global__ void __launch_bounds__(256) k_tile(DevProg P, RunArgs R, PassDev pass,
                                              cx<T>* __restrict__ gstate) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int kt = pass.k_tile;
  const uint32_t tile = 1u << kt;
  const int teams = WARP_TEAM ? (blockDim.x >> 5) : 1;
  const int team = WARP_TEAM ? (threadIdx.x >> 5) : 0;
  const int tsize = WARP_TEAM ? 32 : blockDim.x;
  const int tlane = WARP_TEAM ? (threadIdx.x & 31) : threadIdx.x;
  cx<T>* st = reinterpret_cast<cx<T>*>(smem_raw) + (size_t)team * tile;
  cx<T>* mb = reinterpret_cast<cx<T>*>(smem_raw) + (size_t)teams * tile +
              (size_t)team * pass.matw;

  const int rest_bits = pass.n_bits - kt;
  const int64_t tiles_per_elem = (int64_t)1 << rest_bits;
  const int64_t total = R.batch * tiles_per_elem;
  const int64_t stride = (int64_t)gridDim.x * teams;
  // CTA-wide teams must run the same trip count (they use __syncthreads)
  const int64_t trips = (total + stride - 1) / stride;

  for (int64_t it = 0; it < trips; ++it) {
    const int64_t tid = it * stride + (int64_t)blockIdx.x * teams + team;
    const bool valid = tid < total;
    if (WARP_TEAM && !valid) break;
    const int64_t bl = valid ? tid / tiles_per_elem : 0;
    const uint64_t t = valid ? (uint64_t)(tid % tiles_per_elem) : 0;
    const int64_t b = bl + R.batch_offset;

    // deposit the bits of t into the non-tile positions
    uint64_t base = 0;
    if (!pass.identity_map || rest_bits > 0) {
      int tb = 0, src = 0;
      for (int g = 0; g < pass.n_bits; ++g) {
        if (tb < kt && pass.tile_bits[tb] == g) {
          ++tb;
        } else {
          base |= ((t >> src) & 1ull) << g;
          ++src;
        }
      }
    }
    auto gidx = [&](uint32_t i) -> uint64_t {
      if (pass.identity_map) return base | i;
      uint64_t o = 0;
      for (int j = 0; j < kt; ++j) o |= (uint64_t)((i >> j) & 1u) << pass.tile_bits[j];
      return base | o;
    };
    cx<T>* gs = gstate ? gstate + (size_t)bl * ((size_t)1 << pass.n_bits) : nullptr;

    if (pass.flags & QMLB_PASS_INIT) {
      for (uint32_t i = tlane; i < tile; i += tsize)
        st[i] = mk<T>((i == 0 && base == 0) ? (T)1 : (T)0, (T)0);
    } else if (valid) {
      for (uint32_t i = tlane; i < tile; i += tsize) st[i] = gs[gidx(i)];
    }
    team_sync<WARP_TEAM>();

    for (int w = 0; w < pass.n_windows; ++w) {
      const int2 win = pass.windows[w];
      // prologue: one thread per op evaluates that op's matrix for element b
      for (int j = tlane; j < win.y; j += tsize) {
        const qmlb_op& op = pass.ops[win.x + j];
        if (op.src >= 0)
          eval_source_mem<T>(P, R, RowsDirect{R, b}, op.src, mb + pass.matoff[win.x + j]);
      }
      team_sync<WARP_TEAM>();
      for (int j = 0; j < win.y; ++j) {
        const qmlb_op& op = pass.ops[win.x + j];
        tile_apply_op<T, WARP_TEAM>(P, op, st, mb + pass.matoff[win.x + j], kt, tlane,
                                    tsize, gidx);
        team_sync<WARP_TEAM>();
      }
    }

    if ((pass.flags & QMLB_PASS_STORE) && valid) {
      for (uint32_t i = tlane; i < tile; i += tsize) gs[gidx(i)] = st[i];
    }
    team_sync<WARP_TEAM>();
  }
}

}  // namespace qmlb
