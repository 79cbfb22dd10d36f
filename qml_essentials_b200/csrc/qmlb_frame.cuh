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

// 2x2 on register bit TB of the pairs whose control value is 1.  The control value of
// slot i is ctl_base ^ bit i of smask (it never depends on the target bit itself).
template <typename T, int TB>
__device__ __forceinline__ void frame_ctrl1(RegState<T, FRAME_R>& S, const cx<T> (&m)[4],
                                            unsigned smask, int ctl_base) {
#pragma unroll
  for (int g = 0; g < (1 << (FRAME_R - 1)); ++g) {
    const int i0 = pair_i0<FRAME_R, TB>(g), i1 = i0 | (1 << TB);
    const bool on = (((smask >> i0) & 1u) ^ (unsigned)ctl_base) != 0u;
    const T ar = S.re(i0), ai = S.im(i0), br = S.re(i1), bi = S.im(i1);
    const T xr = m[0].x * ar - m[0].y * ai + m[1].x * br - m[1].y * bi;
    const T xi = m[0].x * ai + m[0].y * ar + m[1].x * bi + m[1].y * br;
    const T yr = m[2].x * ar - m[2].y * ai + m[3].x * br - m[3].y * bi;
    const T yi = m[2].x * ai + m[2].y * ar + m[3].x * bi + m[3].y * br;
    S.re(i0) = on ? xr : ar;
    S.im(i0) = on ? xi : ai;
    S.re(i1) = on ? yr : br;
    S.im(i1) = on ? yi : bi;
  }
}

// HEAVY = the program holds a dense op on 3 or 4 bits (rare: 2-qubit channels, CCX as a
// matrix); the lean variant keeps the register pressure of the common steps low.
template <typename T, int THREADS, bool HEAVY>
__global__ void __launch_bounds__(THREADS, THREADS == 512 ? 1 : 2)
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
  cx<T>* tiles = reinterpret_cast<cx<T>*>(fsm);
  cx<T>* mats_all = tiles + (size_t)teams * tile_n;
  FrameStep* sstep = reinterpret_cast<FrameStep*>(mats_all + (size_t)teams * F.mat_cap);
  uint32_t* tab_lo = reinterpret_cast<uint32_t*>(sstep + 2);
  uint32_t* tab_hi = tab_lo + 256;
  double* red = reinterpret_cast<double*>(tab_hi + 64);
  cx<T>* tile = tiles + (size_t)team * tile_n;
  cx<T>* mats = mats_all + (size_t)team * F.mat_cap;

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

  for (int64_t rd = 0; rd < rounds; ++rd) {
    const int64_t bl = rd * per_round + cluster_id * teams + team;
    const bool valid = bl < A.batch;
    const cx<T>* prow = premats + (size_t)(valid ? bl : 0) * F.premat_row;

    // |0..0>: the frame is linear, so logical index 0 sits at physical index 0
    for (uint32_t i = tlane; i < tile_n; i += tsize)
      tile[i] = mk<T>((i == 0 && rank == 0) ? (T)1 : (T)0, (T)0);
    for (int i = threadIdx.x; i < 256; i += THREADS)
      reinterpret_cast<uint32_t*>(&sstep[0])[i] =
          reinterpret_cast<const uint32_t*>(&F.steps[0])[i];
    sync_all();

    for (int si = 0; si < F.n_steps; ++si) {
      const FrameStep& st = sstep[si & 1];
      if (si + 1 < F.n_steps)
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
        sync_all();  // tables visible; every CTA of the cluster has finished its previous step
        constexpr int PER = sizeof(T) == 8 ? 16 : 32;
        const int per = (int)(tile_n >> F.team_bits);
        cx<T> hold[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) {
          if (k < per) {
            const uint32_t d = (uint32_t)tlane + ((uint32_t)k << F.team_bits);
            const uint32_t src = cmine ^ tab_lo[d & 255u] ^ tab_hi[d >> 8];
            const uint32_t r = src >> Tb, loc = src & (tile_n - 1u);
            const cx<T>* from = tile;
            if (clustered && r != rank) from = cluster.map_shared_rank(tile, r);
            hold[k] = from[loc];
          }
        }
        sync_all();
#pragma unroll
        for (int k = 0; k < PER; ++k)
          if (k < per) tile[(uint32_t)tlane + ((uint32_t)k << F.team_bits)] = hold[k];
        sync_all();
        continue;
      }

      // ---- SUBPASS: matrices (with their XOR variants) into shared memory -----------------
      if (valid) {
        for (int o = 0; o < st.n_ops; ++o) {
          const FrameOp fo = st.ops[o];
          cx<T>* dst = mats + fo.smem_off;
          const cx<T>* srcm = prow + fo.premat_off;
          if (fo.code == QMLB_FOP_DIAG) {
            for (int e = tlane; e < (1 << fo.k); e += tsize) dst[e] = srcm[e];
            ++o;  // parity-row indices
          } else if (fo.code == QMLB_FOP_MATK) {
            for (int e = tlane; e < (1 << (2 * fo.k)); e += tsize) dst[e] = srcm[e];
          } else {
            const int k = (fo.code == QMLB_FOP_MAT2) ? 2 : 1, dd = 1 << k, ee = dd * dd;
            for (int e = tlane; e < ee * fo.nvar; e += tsize) {
              const int c = e / ee, idx = e % ee;
              int v = (idx >> k) ^ c, u = (idx & (dd - 1)) ^ c;
              if (fo.flags & 1) {  // logical (bits[0], bits[1]) sit at register bits (j1, j0)
                v = ((v & 1) << 1) | (v >> 1);
                u = ((u & 1) << 1) | (u >> 1);
              }
              dst[e] = srcm[v * dd + u];
            }
          }
        }
      }
      __syncthreads();

      if (valid) {
        uint32_t piv[FRAME_R];
#pragma unroll
        for (int j = 0; j < FRAME_R; ++j) piv[j] = st.pivots[j];
        for (uint32_t it = tlane; it < n_items; it += tsize) {
          const uint32_t base = frame_deposit(it, piv);
          RegState<T, FRAME_R> S;
#pragma unroll
          for (int v = 0; v < FRAME_D; ++v) {
            const cx<T> a = tile[base ^ st.eoff[v]];
            S.re(v) = a.x;
            S.im(v) = a.y;
          }
          int cj[FRAME_R];
#pragma unroll
          for (int j = 0; j < FRAME_R; ++j)
            cj[j] = (__popc(base & st.par[j].rloc) ^ __popc(rank & st.par[j].rout)) & 1;
          auto parity_at = [&](int pi) -> int {
            return (__popc(base & st.par[pi].rloc) ^ __popc(rank & st.par[pi].rout)) & 1;
          };

#pragma unroll 1
          for (int o = 0; o < st.n_ops; ++o) {
            const FrameOp fo = st.ops[o];
            const cx<T>* m = mats + fo.smem_off;
            switch (fo.code) {
              case QMLB_FOP_MAT1: {
                cx<T> mm[4];
                dispatch1<T, FRAME_R>(fo.j0, [&](auto B) {
                  constexpr int BIT = decltype(B)::value;
                  const cx<T>* mv = m + (fo.nvar > 1 ? 4 * cj[BIT] : 0);
#pragma unroll
                  for (int i = 0; i < 4; ++i) mm[i] = mv[i];
                  reg_mat1<T, FRAME_R, BIT>(S, mm);
                });
                break;
              }
              case QMLB_FOP_MAT2:
                dispatch2<T, FRAME_R>(fo.j0, fo.j1, [&](auto JA, auto JB) {
                  constexpr int A_ = decltype(JA)::value, B_ = decltype(JB)::value;
                  if constexpr (A_ > B_) {
                    const int c = fo.nvar > 1 ? ((cj[A_] << 1) | cj[B_]) : 0;
                    reg_mat2<T, FRAME_R, A_, B_>(S, m + 16 * c);
                  }
                });
                break;
              case QMLB_FOP_MATK: if constexpr (HEAVY) {
                int c = 0;
#pragma unroll
                for (int j = 0; j < FRAME_R; ++j)
                  if (j < fo.k) c |= cj[j] << j;
                if (fo.k == 3)
                  frame_matk<T, 3>(S, m, c);
                else
                  frame_matk<T, 4>(S, m, c);
              } break;
              case QMLB_FOP_CTRL1: {
                cx<T> mm[4];
                const int ctl = parity_at(fo.j1);
                const unsigned sm = st.par[fo.j1].smask;
                dispatch1<T, FRAME_R>(fo.j0, [&](auto B) {
                  constexpr int BIT = decltype(B)::value;
                  const cx<T>* mv = m + (fo.nvar > 1 ? 4 * cj[BIT] : 0);
#pragma unroll
                  for (int i = 0; i < 4; ++i) mm[i] = mv[i];
                  frame_ctrl1<T, BIT>(S, mm, sm, ctl);
                });
                break;
              }
              case QMLB_FOP_DIAG: {
                const uint8_t* idx = reinterpret_cast<const uint8_t*>(&st.ops[o + 1]);
                int lb = 0;        // local value of slot 0
                unsigned flip[8];  // per op bit: which slots see it flipped
                for (int a = 0; a < fo.k; ++a) {
                  lb |= parity_at(idx[a]) << (fo.k - 1 - a);
                  flip[a] = st.par[idx[a]].smask;
                }
#pragma unroll
                for (int v = 0; v < FRAME_D; ++v) {
                  int loc = lb;
                  for (int a = 0; a < fo.k; ++a)
                    loc ^= (int)((flip[a] >> v) & 1u) << (fo.k - 1 - a);
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

#pragma unroll
          for (int v = 0; v < FRAME_D; ++v) tile[base ^ st.eoff[v]] = mk<T>(S.re(v), S.im(v));
        }
      }
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
