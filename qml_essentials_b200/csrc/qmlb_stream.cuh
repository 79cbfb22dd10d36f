// Streaming gate-pass kernel for states that live in HBM (strategy 2): the large-n
// statevector regime (BASELINE config 5) and batched density matrices that do not
// fit on chip (config 4).
//
// One launch = one fused gate pass = the state is read once and written once.
// A pass owns a GROUP of R state bits (R = 5 for complex64, 4 for complex128).
// Every thread takes work item w (the other N - R bits), loads the 2^R amplitudes
// of its group into registers, applies EVERY op of the pass that acts inside the
// group (fused 1-qubit chains, controlled 2x2, CX / 2-bit permutations, dense 4x4
// such as 1-qubit superoperators on (ket, bra), dense 8x8 / 16x16, diagonals on
// any bits) with compile-time register indices, and stores them back.
//
// Memory behaviour: lanes run over consecutive w, i.e. over the lowest bits that
// are NOT in the group, so for a group of high-order qubits every warp-wide access
// is 32 consecutive amplitudes (256 / 512 contiguous bytes, fully coalesced); the
// 2^R loads of a thread are independent (no dependent address), which puts
// 256 B per thread = 64 KB per CTA in flight.  Groups that contain low-order
// bits make each lane walk its own 32 B sectors; the host scheduler keeps bits
// 0/1 together so that every fetched sector is fully used by one thread.
// Matrices are evaluated once per (CTA, element) into shared memory and read by
// broadcast.  The first pass of a run does not read the state (|0..0> is implied).
#pragma once

#include "qmlb_reg.cuh"
#include "qmlb_stream_types.h"

namespace qmlb {

// cp.async (LDGSTS): global -> shared without staging registers
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem)),
               "l"(gmem));
}
template <typename T>
__device__ __forceinline__ void cp_async_elem(cx<T>* smem, const cx<T>* gmem) {
  if constexpr (sizeof(T) == 8) {
    cp_async16(smem, gmem);
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem));
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n"); }
__device__ __forceinline__ void cp_async_wait_all_but_one() {
  asm volatile("cp.async.wait_group 1;\n" ::: "memory");
}

// dense 4x4 (row-major, local value v = (bit JA << 1) | bit JB) on register bits JA > JB
template <typename T, int N, int JA, int JB>
__device__ __forceinline__ void reg_mat2(RegState<T, N>& S, const cx<T>* __restrict__ m) {
  static_assert(JA > JB, "canonical order");
  if constexpr (std::is_same<T, float>::value && JB >= 1) {
    // packed FP32: both bits >= 1, so amplitudes 2k / 2k+1 share their role (see reg_mat1)
    constexpr int A = JA - 1, B = JB - 1;
#pragma unroll
    for (int g = 0; g < (1 << (N - 3)); ++g) {
      const int t = ((g >> B) << (B + 1)) | (g & ((1 << B) - 1));
      const int k00 = ((t >> A) << (A + 1)) | (t & ((1 << A) - 1));
      const int idx[4] = {k00, k00 | (1 << B), k00 | (1 << A), k00 | (1 << A) | (1 << B)};
      float2 ar[4], ai[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ar[u] = S.r2[idx[u]];
        ai[u] = S.i2[idx[u]];
      }
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        float2 xr = make_float2(0.f, 0.f), xi = make_float2(0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const cx<float> c = m[v * 4 + u];  // shared memory, broadcast
          const float2 cr = make_float2(c.x, c.x), ci = make_float2(c.y, c.y);
          const float2 nci = make_float2(-c.y, -c.y);
          xr = __ffma2_rn(cr, ar[u], xr);
          xr = __ffma2_rn(nci, ai[u], xr);
          xi = __ffma2_rn(cr, ai[u], xi);
          xi = __ffma2_rn(ci, ar[u], xi);
        }
        S.r2[idx[v]] = xr;
        S.i2[idx[v]] = xi;
      }
    }
  } else {
#pragma unroll
    for (int g = 0; g < (1 << (N - 2)); ++g) {
      // insert zeros at JB then JA
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
          const cx<T> c = m[v * 4 + u];  // shared memory, broadcast
          xr = fma(c.x, ar[u], xr);
          xr = fma(-c.y, ai[u], xr);
          xi = fma(c.x, ai[u], xi);
          xi = fma(c.y, ar[u], xi);
        }
        S.re(idx[v]) = xr;
        S.im(idx[v]) = xi;
      }
    }
  }
}

// dense 2^K x 2^K on register bits K-1..0 (local value = low K bits of the register index)
template <typename T, int N, int K>
__device__ __forceinline__ void reg_matk(RegState<T, N>& S,
                                         const cx<T>* __restrict__ m) {
  constexpr int D = 1 << K;
#pragma unroll
  for (int blk = 0; blk < (1 << (N - K)); ++blk) {
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
        const cx<T> c = m[v * D + u];
        xr = fma(c.x, ar[u], xr);
        xr = fma(-c.y, ai[u], xr);
        xi = fma(c.x, ai[u], xi);
        xi = fma(c.y, ar[u], xi);
      }
      // v is a runtime loop variable: scatter through a uniform select chain
#pragma unroll
      for (int w = 0; w < D; ++w)
        if (w == v) {
          S.re((blk << K) | w) = xr;  // safe: row v only reads the ar/ai copies
          S.im((blk << K) | w) = xi;
        }
    }
  }
}

// permutation of the low K register bits: new[v] = old[p[v]], p packed K bits per entry
template <typename T, int N, int K>
__device__ __forceinline__ void reg_permk(RegState<T, N>& S,
                                          unsigned long long packed) {
  constexpr int D = 1 << K;
#pragma unroll
  for (int blk = 0; blk < (1 << (N - K)); ++blk) {
    T ar[D], ai[D];
#pragma unroll
    for (int u = 0; u < D; ++u) {
      ar[u] = S.re((blk << K) | u);
      ai[u] = S.im((blk << K) | u);
    }
#pragma unroll
    for (int v = 0; v < D; ++v) {
      const int pv = (int)((packed >> (K * v)) & (D - 1));
      T xr = ar[0], xi = ai[0];
#pragma unroll
      for (int u = 1; u < D; ++u)
        if (pv == u) {
          xr = ar[u];
          xi = ai[u];
        }
      S.re((blk << K) | v) = xr;
      S.im((blk << K) | v) = xi;
    }
  }
}

// 2-bit permutation on register bits JA > JB (v = (bit JA << 1) | bit JB), p packed 2 bits/entry
template <typename T, int N, int JA, int JB>
__device__ __forceinline__ void reg_perm2(RegState<T, N>& S, unsigned packed) {
  if (packed == 0xB4u) {  // (0,1,3,2): CX, control JA, target JB
    reg_cx<T, N, JA, JB>(S);
    return;
  }
  if (packed == 0x6Cu) {  // (0,3,2,1): CX, control JB, target JA
    reg_cx<T, N, JB, JA>(S);
    return;
  }
#pragma unroll
  for (int g = 0; g < (1 << (N - 2)); ++g) {
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
      const int pv = (packed >> (2 * v)) & 3;
      T xr = ar[0], xi = ai[0];
#pragma unroll
      for (int u = 1; u < 4; ++u)
        if (pv == u) {
          xr = ar[u];
          xi = ai[u];
        }
      S.re(idx[v]) = xr;
      S.im(idx[v]) = xi;
    }
  }
}

// Batched streaming runs (density matrices, config 4): one thread per (element, matrix)
// evaluates the matrix sources of ALL passes once into a table, so the gate-pass CTAs only
// copy their few hundred bytes instead of each re-deriving sincos + 4x4 chain products.
// PLAIN: no entry needs the role swap or the Pauli transform - the light variant (64
// registers, 16 warps per SM instead of 10 resident: the kernel is a chain of dependent
// loads through the source / angle / argument records, i.e. latency bound; config 3:
// 70 -> ~25 us for 960 000 matrices)
template <typename T, bool PLAIN>
__global__ void __launch_bounds__(128, PLAIN ? 8 : 2)
    k_stream_mats(DevProg P, RunArgs A, const StreamMatOp* __restrict__ list, int n_list,
                  int mat_row, cx<T>* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.batch * n_list) return;
  const int64_t bl = t / n_list;
  const StreamMatOp mo = list[t % n_list];
  cx<T>* dst = out + (size_t)bl * mat_row + mo.off;
  const RowsDirect rows{A, bl + A.batch_offset};
  if constexpr (PLAIN) {
    // 2x2 sources (every statevector gate): registers only - the generic path below keeps
    // its matrices in local memory
    const qmlb_source s0 = P.src[mo.src];
    if (s0.k == 1 && (s0.kind == QMLB_SRC_TRIG || s0.kind == QMLB_SRC_CHAIN ||
                      s0.kind == QMLB_SRC_PRE)) {
      cx<T> m[4];
      eval_2x2<T>(P, A, rows, mo.src, m);
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[i] = m[i];
      return;
    }
    eval_source_mem<T>(P, A, rows, mo.src, dst);
    return;
  }
  if (mo.swap2 >= 2) {
    // Pauli transfer matrix of a 1-qubit superoperator: R = T S T^-1 with S in (ket, bra)
    // order [rho00, rho01, rho10, rho11], Pauli order (I, Z, X, Y) and T^-1 = T^dagger / 2,
    //   T = [[1,0,0,1],[1,0,0,-1],[0,1,1,0],[0,i,-i,0]].
    // R is real for every completely positive map; its 16 reals go to the first half of
    // the op's 16 complex slots (row-major; swap2 == 3: roles of the two bits exchanged).
    cx<T> S[16];
    eval_source_mem<T>(P, A, rows, mo.src, S);
    T* rdst = reinterpret_cast<T*>(dst);
    // X = T S: rows I, Z, X, Y
    cx<T> X[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      X[0 * 4 + j] = cadd(S[0 * 4 + j], S[3 * 4 + j]);
      X[1 * 4 + j] = mk<T>(S[0 * 4 + j].x - S[3 * 4 + j].x, S[0 * 4 + j].y - S[3 * 4 + j].y);
      X[2 * 4 + j] = cadd(S[1 * 4 + j], S[2 * 4 + j]);
      // i * (S1 - S2)
      X[3 * 4 + j] = mk<T>(-(S[1 * 4 + j].y - S[2 * 4 + j].y), S[1 * 4 + j].x - S[2 * 4 + j].x);
    }
    // R = X T^dagger / 2: columns I: (x0 + x3)/2, Z: (x0 - x3)/2, X: (x1 + x2)/2,
    // Y: (conj(i) x1 + conj(-i) x2)/2 = (-i x1 + i x2)/2 -> real part = (x1.y - x2.y)/2
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const cx<T> x0 = X[a * 4 + 0], x1 = X[a * 4 + 1], x2 = X[a * 4 + 2], x3 = X[a * 4 + 3];
      T r[4];
      r[0] = (T)0.5 * (x0.x + x3.x);
      r[1] = (T)0.5 * (x0.x - x3.x);
      r[2] = (T)0.5 * (x1.x + x2.x);
      r[3] = (T)0.5 * (x1.y - x2.y);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        int ra = a, rb = b;
        if (mo.swap2 == 3) {
          ra = ((a & 1) << 1) | (a >> 1);
          rb = ((b & 1) << 1) | (b >> 1);
        }
        rdst[ra * 4 + rb] = r[b];
      }
    }
  } else if (mo.swap2) {
    cx<T> tmp[16];
    eval_source_mem<T>(P, A, rows, mo.src, tmp);
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int sv = ((v & 1) << 1) | (v >> 1), su = ((u & 1) << 1) | (u >> 1);
        dst[sv * 4 + su] = tmp[v * 4 + u];
      }
  } else {
    eval_source_mem<T>(P, A, rows, mo.src, dst);
  }
}

// HEAVY = the pass holds a dense / permutation op on 3 or 4 bits (rare: CCX, CSWAP,
// 2-qubit channels); the lean variant keeps the register count of the common passes low.
// IDX = uint32_t when every amplitude index of the launch fits 32 bits (one element of at
// most 32 state bits), else uint64_t.
template <typename T, int R, bool HEAVY, typename IDX>
__global__ void __launch_bounds__(STREAM_THREADS, HEAVY ? 1 : (sizeof(T) == 4 ? 4 : STREAM_MIN_CTAS))
    k_stream(DevProg P, RunArgs A, const __grid_constant__ StreamPass pass,
             cx<T>* __restrict__ gstate, const cx<T>* __restrict__ premats,
             const StreamPeers peers) {
  constexpr int D = 1 << R;
  // complex64: amplitudes v and v^1 are moved as one 16-byte access when register bit 0
  // is state bit 0 (the scheduler then also puts state bit 1 at register bit 1, so a
  // thread owns whole 32-byte sectors)
  constexpr bool CAN_PAIR = sizeof(T) == 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [2 stages x 2^R x STREAM_THREADS amplitudes | matrices of the pass]
  cx<T>* stg = reinterpret_cast<cx<T>*>(smem_raw);
  cx<T>* mb = stg + 2 * D * STREAM_THREADS;
  const bool init_pass = pass.flags & QMLB_PASS_INIT;

  const int N = pass.n_bits;
  const IDX items = (IDX)1 << (N - R);
  IDX ob[R];  // offset of register bit j in the state index
#pragma unroll
  for (int j = 0; j < R; ++j) ob[j] = (IDX)1 << pass.gb[j];
  auto offv = [&](int v) -> IDX {
    IDX o = 0;
#pragma unroll
    for (int j = 0; j < R; ++j)
      if ((v >> j) & 1) o |= ob[j];
    return o;
  };
  const bool paired = CAN_PAIR && pass.gb[0] == 0;
  // element offsets of the 2^R group members: uniform for the whole launch, computed once;
  // an amplitude address is then (state + base) + eoff[v] - one wide multiply-add
  IDX eoff[D];
#pragma unroll
  for (int v = 0; v < D; ++v) eoff[v] = offv(v);

  for (int64_t bl = blockIdx.y; bl < A.batch; bl += gridDim.y) {
    const int64_t b = bl + A.batch_offset;
    __syncthreads();  // previous element's matrices are no longer read
    if (premats != nullptr) {
      // batched run: the matrices of every (element, op) were evaluated by k_stream_mats
      const cx<T>* row = premats + (size_t)bl * pass.mat_row + pass.mat_base;
      for (int e = threadIdx.x; e < pass.matw; e += blockDim.x) mb[e] = row[e];
    } else
    for (int j = threadIdx.x; j < pass.n_ops; j += blockDim.x) {
      const StreamOp op = pass.ops[j];
      if (op.kind == QMLB_OP_PERM) continue;
      cx<T>* dst = mb + pass.matoff[j];
      if (op.kind == QMLB_OP_MAT && op.k == 2 && op.b0 < op.b1) {
        // canonical register order (JA > JB): swap the roles of the two local bits
        cx<T> tmp[16];
        eval_source_mem<T>(P, A, RowsDirect{A, b}, op.src, tmp);
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int sv = ((v & 1) << 1) | (v >> 1), su = ((u & 1) << 1) | (u >> 1);
            dst[sv * 4 + su] = tmp[v * 4 + u];
          }
      } else {
        eval_source_mem<T>(P, A, RowsDirect{A, b}, op.src, dst);
      }
    }
    __syncthreads();

    cx<T>* gs = gstate + (size_t)bl * ((size_t)1 << N);
    const IDX stride = (IDX)gridDim.x * blockDim.x;
    auto base_of = [&](IDX w) -> IDX {
      IDX base = w;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int pbit = pass.sorted[j];
        base = ((base >> pbit) << (pbit + 1)) | (base & (((IDX)1 << pbit) - (IDX)1));
      }
      return base;
    };
    // Software pipeline: the 2^R amplitudes of the NEXT work item travel global -> shared
    // with cp.async (LDGSTS, no registers held, thread-private slots so no barrier) while
    // the current item is in the arithmetic; with ~12 resident warps per SM that is what
    // keeps HBM requests in flight during the compute phase.
    // where amplitude idx is read from: this GPU's state, or (fused exchange) the peer
    // that holds it before the global<->local swap
    auto src_of = [&](IDX idx) -> const cx<T>* {
      if (!peers.enabled) return gs + idx;
      const IDX chunk = idx >> peers.cshift;
      const IDX within = idx & (((IDX)1 << peers.cshift) - (IDX)1);
      return static_cast<const cx<T>*>(peers.ptr[chunk]) +
             (((IDX)peers.rank << peers.cshift) | within);
    };
    auto prefetch = [&](IDX w, int stage) {
      if (w < items && !init_pass) {
        const IDX base = base_of(w);
        cx<T>* slot = stg + (size_t)stage * D * STREAM_THREADS + threadIdx.x;
        if (peers.enabled) {  // fused exchange: general addressing through the peer table
          if (paired) {
            if constexpr (CAN_PAIR) {
              float4* slot4 = reinterpret_cast<float4*>(stg) +
                              (size_t)stage * (D / 2) * STREAM_THREADS + threadIdx.x;
#pragma unroll
              for (int v = 0; v < D; v += 2)
                cp_async16(slot4 + (v >> 1) * STREAM_THREADS, src_of(base | eoff[v]));
            }
          } else {
#pragma unroll
            for (int v = 0; v < D; ++v)
              cp_async_elem<T>(slot + v * STREAM_THREADS, src_of(base | eoff[v]));
          }
        } else {
          const cx<T>* pb = gs + base;
          if (paired) {
            if constexpr (CAN_PAIR) {
              float4* slot4 = reinterpret_cast<float4*>(stg) +
                              (size_t)stage * (D / 2) * STREAM_THREADS + threadIdx.x;
#pragma unroll
              for (int v = 0; v < D; v += 2)
                cp_async16(slot4 + (v >> 1) * STREAM_THREADS, pb + eoff[v]);
            }
          } else {
#pragma unroll
            for (int v = 0; v < D; ++v) cp_async_elem<T>(slot + v * STREAM_THREADS, pb + eoff[v]);
          }
        }
      }
      cp_async_commit();
    };

    const IDX w_first = (IDX)blockIdx.x * blockDim.x + threadIdx.x;
    prefetch(w_first, 0);
    int stage = 0;
    for (IDX w = w_first; w < items; w += stride, stage ^= 1) {
      prefetch(w + stride < w ? items : w + stride, stage ^ 1);  // (overflow-safe)
      cp_async_wait_all_but_one();
      const IDX base = base_of(w);
      RegState<T, R> S;
      if (init_pass) {
#pragma unroll
        for (int v = 0; v < D; ++v) {
          S.re(v) = (base == 0 && v == 0 && !(pass.flags & QMLB_PASS_INIT_ZERO)) ? (T)1 : (T)0;
          S.im(v) = (T)0;
        }
      } else if (paired) {
        if constexpr (CAN_PAIR) {
          const float4* slot4 = reinterpret_cast<const float4*>(stg) +
                                (size_t)stage * (D / 2) * STREAM_THREADS + threadIdx.x;
#pragma unroll
          for (int v = 0; v < D; v += 2) {
            const float4 a = slot4[(v >> 1) * STREAM_THREADS];
            S.r2[v >> 1] = make_float2(a.x, a.z);
            S.i2[v >> 1] = make_float2(a.y, a.w);
          }
        }
      } else {
        const cx<T>* slot = stg + (size_t)stage * D * STREAM_THREADS + threadIdx.x;
#pragma unroll
        for (int v = 0; v < D; ++v) {
          const cx<T> a = slot[v * STREAM_THREADS];
          S.re(v) = a.x;
          S.im(v) = a.y;
        }
      }

      for (int o = 0; o < pass.n_ops; ++o) {
        const StreamOp op = pass.ops[o];
        const cx<T>* m = mb + pass.matoff[o];
        switch (op.kind) {
          case QMLB_OP_MAT:
            if (op.k == 1) {
              cx<T> mm[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) mm[i] = m[i];
              dispatch1<T, R>(op.b0, [&](auto B) {
                reg_mat1<T, R, decltype(B)::value>(S, mm);
              });
            } else if (op.k == 2) {
              const int ja = max(op.b0, op.b1), jb = min(op.b0, op.b1);
              dispatch2<T, R>(ja, jb, [&](auto JA, auto JB) {
                if constexpr (decltype(JA)::value > decltype(JB)::value)
                  reg_mat2<T, R, decltype(JA)::value, decltype(JB)::value>(S, m);
              });
            } else if (HEAVY && op.k == 3) {
              reg_matk<T, R, 3>(S, m);
            } else if (HEAVY) {
              reg_matk<T, R, 4>(S, m);
            }
            break;
          case QMLB_OP_CTRL1: {
            cx<T> mm[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) mm[i] = m[i];
            dispatch2<T, R>(op.b0, op.b1, [&](auto CB, auto TB) {
              reg_ctrl1<T, R, decltype(CB)::value, decltype(TB)::value>(S, mm);
            });
            break;
          }
          case QMLB_OP_PERM: {
            if (op.k == 1) {
              dispatch1<T, R>(op.b0, [&](auto B) {
                constexpr int BIT = decltype(B)::value;
#pragma unroll
                for (int g = 0; g < (1 << (R - 1)); ++g) {
                  const int i0 = pair_i0<R, BIT>(g), i1 = i0 | (1 << BIT);
                  swap_renamed(S.re(i0), S.re(i1));
                  swap_renamed(S.im(i0), S.im(i1));
                }
              });
            } else if (op.k == 2) {
              dispatch2<T, R>(op.b0, op.b1, [&](auto JA, auto JB) {
                if constexpr (decltype(JA)::value > decltype(JB)::value)
                  reg_perm2<T, R, decltype(JA)::value, decltype(JB)::value>(S,
                                                                            (unsigned)op.data);
              });
            } else if (HEAVY && op.k == 3) {
              reg_permk<T, R, 3>(S, op.data);
            } else if (HEAVY) {
              reg_permk<T, R, 4>(S, op.data);
            }
            break;
          }
          case QMLB_OP_DIAG: {
            // diagonal on GLOBAL bits (any position): d[v] from shared memory
#pragma unroll
            for (int v = 0; v < D; ++v) {
              const uint64_t gi = (uint64_t)(base | eoff[v]);
              int loc = 0;
              for (int j = 0; j < op.k; ++j)
                loc |= (int)((gi >> ((op.data >> (6 * j)) & 63)) & 1ull) << (op.k - 1 - j);
              const cx<T> d = m[loc];
              const T r = S.re(v), q = S.im(v);
              S.re(v) = d.x * r - d.y * q;
              S.im(v) = d.x * q + d.y * r;
            }
            break;
          }
        }
      }

      cx<T>* pw = gs + base;
      if (paired) {
        if constexpr (CAN_PAIR) {
#pragma unroll
          for (int v = 0; v < D; v += 2)
            *reinterpret_cast<float4*>(pw + eoff[v]) =
                make_float4(S.r2[v >> 1].x, S.i2[v >> 1].x, S.r2[v >> 1].y, S.i2[v >> 1].y);
        }
      } else {
#pragma unroll
        for (int v = 0; v < D; ++v) pw[eoff[v]] = mk<T>(S.re(v), S.im(v));
      }
    }
  }
}

}  // namespace qmlb
