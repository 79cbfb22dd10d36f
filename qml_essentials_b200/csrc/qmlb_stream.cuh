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

// dense 4x4 (row-major, local value v = (bit JA << 1) | bit JB) on register bits JA > JB
template <typename T, int N, int JA, int JB>
__device__ __forceinline__ void reg_mat2(T (&re)[1 << N], T (&im)[1 << N],
                                         const cx<T>* __restrict__ m) {
  static_assert(JA > JB, "canonical order");
#pragma unroll
  for (int g = 0; g < (1 << (N - 2)); ++g) {
    // insert zeros at JB then JA
    const int t = ((g >> JB) << (JB + 1)) | (g & ((1 << JB) - 1));
    const int i00 = ((t >> JA) << (JA + 1)) | (t & ((1 << JA) - 1));
    const int idx[4] = {i00, i00 | (1 << JB), i00 | (1 << JA), i00 | (1 << JA) | (1 << JB)};
    T ar[4], ai[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ar[u] = re[idx[u]];
      ai[u] = im[idx[u]];
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
      re[idx[v]] = xr;
      im[idx[v]] = xi;
    }
  }
}

// dense 2^K x 2^K on register bits K-1..0 (local value = low K bits of the register index)
template <typename T, int N, int K>
__device__ __forceinline__ void reg_matk(T (&re)[1 << N], T (&im)[1 << N],
                                         const cx<T>* __restrict__ m) {
  constexpr int D = 1 << K;
#pragma unroll
  for (int blk = 0; blk < (1 << (N - K)); ++blk) {
    T ar[D], ai[D];
#pragma unroll
    for (int u = 0; u < D; ++u) {
      ar[u] = re[(blk << K) | u];
      ai[u] = im[(blk << K) | u];
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
          re[(blk << K) | w] = xr;  // safe: row v only reads the ar/ai copies
          im[(blk << K) | w] = xi;
        }
    }
  }
}

// permutation of the low K register bits: new[v] = old[p[v]], p packed K bits per entry
template <typename T, int N, int K>
__device__ __forceinline__ void reg_permk(T (&re)[1 << N], T (&im)[1 << N],
                                          unsigned long long packed) {
  constexpr int D = 1 << K;
#pragma unroll
  for (int blk = 0; blk < (1 << (N - K)); ++blk) {
    T ar[D], ai[D];
#pragma unroll
    for (int u = 0; u < D; ++u) {
      ar[u] = re[(blk << K) | u];
      ai[u] = im[(blk << K) | u];
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
      re[(blk << K) | v] = xr;
      im[(blk << K) | v] = xi;
    }
  }
}

// 2-bit permutation on register bits JA > JB (v = (bit JA << 1) | bit JB), p packed 2 bits/entry
template <typename T, int N, int JA, int JB>
__device__ __forceinline__ void reg_perm2(T (&re)[1 << N], T (&im)[1 << N], unsigned packed) {
  if (packed == 0xB4u) {  // (0,1,3,2): CX, control JA, target JB
    reg_cx<T, N, JA, JB>(re, im);
    return;
  }
  if (packed == 0x6Cu) {  // (0,3,2,1): CX, control JB, target JA
    reg_cx<T, N, JB, JA>(re, im);
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
      ar[u] = re[idx[u]];
      ai[u] = im[idx[u]];
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
      re[idx[v]] = xr;
      im[idx[v]] = xi;
    }
  }
}

// HEAVY = the pass holds a dense / permutation op on 3 or 4 bits (rare: CCX, CSWAP,
// 2-qubit channels); the lean variant keeps the register count of the common passes low.
// IDX = uint32_t when every amplitude index of the launch fits 32 bits (one element of at
// most 32 state bits), else uint64_t.
template <typename T, int R, bool HEAVY, typename IDX>
__global__ void __launch_bounds__(STREAM_THREADS, HEAVY ? 1 : STREAM_MIN_CTAS)
    k_stream(DevProg P, RunArgs A, const __grid_constant__ StreamPass pass,
             cx<T>* __restrict__ gstate) {
  constexpr int D = 1 << R;
  // complex64: amplitudes v and v^1 are moved as one 16-byte access when register bit 0
  // is state bit 0 (the scheduler then also puts state bit 1 at register bit 1, so a
  // thread owns whole 32-byte sectors)
  constexpr bool CAN_PAIR = sizeof(T) == 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T>* mb = reinterpret_cast<cx<T>*>(smem_raw);

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

  for (int64_t bl = blockIdx.y; bl < A.batch; bl += gridDim.y) {
    const int64_t b = bl + A.batch_offset;
    __syncthreads();  // previous element's matrices are no longer read
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
    for (IDX w = (IDX)blockIdx.x * blockDim.x + threadIdx.x; w < items;
         w += (IDX)gridDim.x * blockDim.x) {
      IDX base = w;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int pbit = pass.sorted[j];
        base = ((base >> pbit) << (pbit + 1)) | (base & (((IDX)1 << pbit) - (IDX)1));
      }
      T re[D], im[D];
      if (pass.flags & QMLB_PASS_INIT) {
#pragma unroll
        for (int v = 0; v < D; ++v) {
          re[v] = (base == 0 && v == 0 && !(pass.flags & QMLB_PASS_INIT_ZERO)) ? (T)1 : (T)0;
          im[v] = (T)0;
        }
      } else if (paired) {
        if constexpr (CAN_PAIR) {
#pragma unroll
          for (int v = 0; v < D; v += 2) {
            const float4 a = *reinterpret_cast<const float4*>(gs + (base | offv(v)));
            re[v] = a.x;
            im[v] = a.y;
            re[v + 1] = a.z;
            im[v + 1] = a.w;
          }
        }
      } else {
#pragma unroll
        for (int v = 0; v < D; ++v) {
          const cx<T> a = gs[base | offv(v)];
          re[v] = a.x;
          im[v] = a.y;
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
                reg_mat1<T, R, decltype(B)::value>(re, im, mm);
              });
            } else if (op.k == 2) {
              const int ja = max(op.b0, op.b1), jb = min(op.b0, op.b1);
              dispatch2<T, R>(ja, jb, [&](auto JA, auto JB) {
                if constexpr (decltype(JA)::value > decltype(JB)::value)
                  reg_mat2<T, R, decltype(JA)::value, decltype(JB)::value>(re, im, m);
              });
            } else if (HEAVY && op.k == 3) {
              reg_matk<T, R, 3>(re, im, m);
            } else if (HEAVY) {
              reg_matk<T, R, 4>(re, im, m);
            }
            break;
          case QMLB_OP_CTRL1: {
            cx<T> mm[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) mm[i] = m[i];
            dispatch2<T, R>(op.b0, op.b1, [&](auto CB, auto TB) {
              reg_ctrl1<T, R, decltype(CB)::value, decltype(TB)::value>(re, im, mm);
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
                  const T r = re[i0], q = im[i0];
                  re[i0] = re[i1];
                  im[i0] = im[i1];
                  re[i1] = r;
                  im[i1] = q;
                }
              });
            } else if (op.k == 2) {
              dispatch2<T, R>(op.b0, op.b1, [&](auto JA, auto JB) {
                if constexpr (decltype(JA)::value > decltype(JB)::value)
                  reg_perm2<T, R, decltype(JA)::value, decltype(JB)::value>(re, im,
                                                                            (unsigned)op.data);
              });
            } else if (HEAVY && op.k == 3) {
              reg_permk<T, R, 3>(re, im, op.data);
            } else if (HEAVY) {
              reg_permk<T, R, 4>(re, im, op.data);
            }
            break;
          }
          case QMLB_OP_DIAG: {
            // diagonal on GLOBAL bits (any position): d[v] from shared memory
#pragma unroll
            for (int v = 0; v < D; ++v) {
              const uint64_t gi = (uint64_t)(base | offv(v));
              int loc = 0;
              for (int j = 0; j < op.k; ++j)
                loc |= (int)((gi >> ((op.data >> (6 * j)) & 63)) & 1ull) << (op.k - 1 - j);
              const cx<T> d = m[loc];
              const T r = re[v], q = im[v];
              re[v] = d.x * r - d.y * q;
              im[v] = d.x * q + d.y * r;
            }
            break;
          }
        }
      }

      if (paired) {
        if constexpr (CAN_PAIR) {
#pragma unroll
          for (int v = 0; v < D; v += 2)
            *reinterpret_cast<float4*>(gs + (base | offv(v))) =
                make_float4(re[v], im[v], re[v + 1], im[v + 1]);
        }
      } else {
#pragma unroll
        for (int v = 0; v < D; ++v) gs[base | offv(v)] = mk<T>(re[v], im[v]);
      }
    }
  }
}

}  // namespace qmlb
