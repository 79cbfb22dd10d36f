// Register-resident statevector kernel: ONE THREAD PER CIRCUIT EVALUATION.
//
// The small-n data-reuploading regime (BASELINE configs 1 and 2: n = 2, 4) has
// 2^n <= 32 amplitudes - the whole state lives in registers for the entire tape,
// there is no shared-memory or HBM traffic for the state at all, and the only
// global reads are the element's parameters / inputs.  The program is uniform
// across the batch, so the per-op `switch` on (kind, bit) never diverges; each
// case is fully unrolled with compile-time amplitude indices.  The expectation
// values are reduced in registers and written once.
#pragma once

#include <type_traits>

#include "qmlb_device.cuh"

namespace qmlb {

template <int N, int BIT>
__device__ __forceinline__ constexpr int pair_i0(int g) {
  return ((g >> BIT) << (BIT + 1)) | (g & ((1 << BIT) - 1));
}

template <typename T, int N, int BIT>
__device__ __forceinline__ void reg_mat1(T (&re)[1 << N], T (&im)[1 << N], const cx<T> (&m)[4]) {
#pragma unroll
  for (int g = 0; g < (1 << (N - 1)); ++g) {
    const int i0 = pair_i0<N, BIT>(g), i1 = i0 | (1 << BIT);
    const T ar = re[i0], ai = im[i0], br = re[i1], bi = im[i1];
    re[i0] = m[0].x * ar - m[0].y * ai + m[1].x * br - m[1].y * bi;
    im[i0] = m[0].x * ai + m[0].y * ar + m[1].x * bi + m[1].y * br;
    re[i1] = m[2].x * ar - m[2].y * ai + m[3].x * br - m[3].y * bi;
    im[i1] = m[2].x * ai + m[2].y * ar + m[3].x * bi + m[3].y * br;
  }
}

// 2x2 on TB where CB is set
template <typename T, int N, int CB, int TB>
__device__ __forceinline__ void reg_ctrl1(T (&re)[1 << N], T (&im)[1 << N],
                                          const cx<T> (&m)[4]) {
#pragma unroll
  for (int g = 0; g < (1 << (N - 1)); ++g) {
    const int i0 = pair_i0<N, TB>(g), i1 = i0 | (1 << TB);
    if (i0 & (1 << CB)) {
      const T ar = re[i0], ai = im[i0], br = re[i1], bi = im[i1];
      re[i0] = m[0].x * ar - m[0].y * ai + m[1].x * br - m[1].y * bi;
      im[i0] = m[0].x * ai + m[0].y * ar + m[1].x * bi + m[1].y * br;
      re[i1] = m[2].x * ar - m[2].y * ai + m[3].x * br - m[3].y * bi;
      im[i1] = m[2].x * ai + m[2].y * ar + m[3].x * bi + m[3].y * br;
    }
  }
}

// CX: swap the target pair where the control bit is set (pure register moves)
template <typename T, int N, int CB, int TB>
__device__ __forceinline__ void reg_cx(T (&re)[1 << N], T (&im)[1 << N]) {
#pragma unroll
  for (int g = 0; g < (1 << (N - 1)); ++g) {
    const int i0 = pair_i0<N, TB>(g), i1 = i0 | (1 << TB);
    if (i0 & (1 << CB)) {
      const T r = re[i0], q = im[i0];
      re[i0] = re[i1];
      im[i0] = im[i1];
      re[i1] = r;
      im[i1] = q;
    }
  }
}

// ---- compile-time dispatch over runtime (uniform) bit positions ---------------
template <typename T, int N, int BIT>
struct D1 {
  template <typename F>
  static __device__ __forceinline__ void run(int bit, F&& f) {
    if (bit == BIT)
      f(std::integral_constant<int, BIT>{});
    else
      D1<T, N, BIT - 1>::run(bit, f);
  }
};
template <typename T, int N>
struct D1<T, N, -1> {
  template <typename F>
  static __device__ __forceinline__ void run(int, F&&) {}
};

template <typename T, int N, typename F>
__device__ __forceinline__ void dispatch1(int bit, F&& f) {
  D1<T, N, N - 1>::run(bit, f);
}
template <typename T, int N, typename F>
__device__ __forceinline__ void dispatch2(int b0, int b1, F&& f) {
  dispatch1<T, N>(b0, [&](auto B0) {
    dispatch1<T, N>(b1, [&](auto B1) {
      if constexpr (decltype(B0)::value != decltype(B1)::value) f(B0, B1);
    });
  });
}

// Can the register kernel run this op?  `consts` is the host copy of the pool.
// Supported: fused 1-qubit chains, controlled 2x2 (CRX/CRY/CRZ/CPhase/CY/CZ), X, CX,
// 1-bit diagonals and the all-qubit diagonal (Golomb encoding).  Anything else
// (dense 2-qubit matrices, SWAP, 3-qubit gates) runs in the shared-memory kernel.
__host__ inline int reg_cx_orientation(const double* perm) {
  // perm over local value v = (bits[0], bits[1]); returns 0: control = bits[0],
  // 1: control = bits[1], -1: not a CX
  const int p[4] = {(int)perm[0], (int)perm[1], (int)perm[2], (int)perm[3]};
  if (p[0] == 0 && p[1] == 1 && p[2] == 3 && p[3] == 2) return 0;
  if (p[0] == 0 && p[1] == 3 && p[2] == 2 && p[3] == 1) return 1;
  return -1;
}

__host__ inline bool reg_supports(const qmlb_op& op, const double* consts, int n_bits) {
  switch (op.kind) {
    case QMLB_OP_MAT: return op.k == 1;
    case QMLB_OP_CTRL1: return true;
    case QMLB_OP_PERM:
      if (op.k == 1) return true;
      return op.k == 2 && reg_cx_orientation(consts + op.aux) >= 0;
    case QMLB_OP_DIAG: {
      if (op.k == 1) return true;
      if (op.k != n_bits) return false;
      for (int j = 0; j < op.k; ++j)
        if (op.bits[j] != n_bits - 1 - j) return false;
      return true;
    }
  }
  return false;
}

// rows cached in shared memory: slot a of this thread at base[a * 128]
struct RowsShared {
  const int32_t* base;
  __device__ __forceinline__ int64_t operator()(int a) const { return base[a * 128]; }
};

// mode: 0 -> write state, 1 -> probs, 2 -> Z-string expectation values
template <typename T, int N>
__global__ void __launch_bounds__(128) k_reg(DevProg P, RunArgs R, int mode, int n_args,
                                             void* __restrict__ out) {
  constexpr int D = 1 << N;
  const int64_t bl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (bl >= R.batch) return;
  const int64_t b = bl + R.batch_offset;

  // rows of the argument slots this element reads, computed once (64-bit div/mod)
  __shared__ int32_t s_rows[QMLB_MAX_ARGS][128];
#pragma unroll
  for (int a = 0; a < QMLB_MAX_ARGS; ++a)
    if (a < n_args) s_rows[a][threadIdx.x] = (int32_t)((b / R.a[a].div) % R.a[a].mod);
  const RowsShared rows{&s_rows[0][threadIdx.x]};

  T re[D], im[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    re[i] = (T)0;
    im[i] = (T)0;
  }
  re[0] = (T)1;

  for (int o = 0; o < P.n_ops; ++o) {
    const qmlb_op op = P.ops[o];
    if (op.kind == QMLB_OP_PERM) {
      if (op.k == 1) {  // the only non-identity 1-bit permutation is X
        dispatch1<T, N>(op.bits[0], [&](auto B) {
          constexpr int BIT = decltype(B)::value;
#pragma unroll
          for (int g = 0; g < (1 << (N - 1)); ++g) {
            const int i0 = pair_i0<N, BIT>(g), i1 = i0 | (1 << BIT);
            const T r = re[i0], q = im[i0];
            re[i0] = re[i1];
            im[i0] = im[i1];
            re[i1] = r;
            im[i1] = q;
          }
        });
      } else {
        // CX; perm[1] == 1 <=> control is bits[0]
        const bool c0 = (int)P.consts[op.aux + 1] == 1;
        const int cb = c0 ? op.bits[0] : op.bits[1], tb = c0 ? op.bits[1] : op.bits[0];
        dispatch2<T, N>(cb, tb, [&](auto CB, auto TB) {
          reg_cx<T, N, decltype(CB)::value, decltype(TB)::value>(re, im);
        });
      }
      continue;
    }
    if (op.kind == QMLB_OP_DIAG && op.k > 1) {  // all bits, MSB first
      const qmlb_source s = P.src[op.src];
      if (s.kind == QMLB_SRC_DIAGPH) {
        const double th = eval_angle(P, R, rows, s.angle);
#pragma unroll
        for (int i = 0; i < D; ++i) {
          T sn, cs;
          sincos_t((T)(-P.consts[s.a0 + i] * th), &sn, &cs);
          const T r = re[i], q = im[i];
          re[i] = cs * r - sn * q;
          im[i] = cs * q + sn * r;
        }
      } else {
#pragma unroll
        for (int i = 0; i < D; ++i) {
          const cx<T> c = ld_const<T>(P.consts, s.a0 + i);
          const T r = re[i], q = im[i];
          re[i] = c.x * r - c.y * q;
          im[i] = c.x * q + c.y * r;
        }
      }
      continue;
    }
    // remaining kinds take one 2x2 matrix
    cx<T> m[4];
    if (op.kind == QMLB_OP_DIAG) {
      const qmlb_source s = P.src[op.src];
      m[1] = mk<T>(0, 0);
      m[2] = mk<T>(0, 0);
      if (s.kind == QMLB_SRC_DIAGPH) {
        const double th = eval_angle(P, R, rows, s.angle);
        T sn, cs;
        sincos_t((T)(-P.consts[s.a0] * th), &sn, &cs);
        m[0] = mk<T>(cs, sn);
        sincos_t((T)(-P.consts[s.a0 + 1] * th), &sn, &cs);
        m[3] = mk<T>(cs, sn);
      } else {
        m[0] = ld_const<T>(P.consts, s.a0);
        m[3] = ld_const<T>(P.consts, s.a0 + 1);
      }
    } else {
      eval_2x2<T>(P, R, rows, op.src, m);
    }
    if (op.kind == QMLB_OP_CTRL1) {
      dispatch2<T, N>(op.bits[0], op.bits[1], [&](auto CB, auto TB) {
        reg_ctrl1<T, N, decltype(CB)::value, decltype(TB)::value>(re, im, m);
      });
    } else {
      dispatch1<T, N>(op.bits[0], [&](auto B) {
        reg_mat1<T, N, decltype(B)::value>(re, im, m);
      });
    }
  }

  if (mode == 0) {
    cx<T>* o = reinterpret_cast<cx<T>*>(out) + bl * D;
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = mk<T>(re[i], im[i]);
  } else if (mode == 1) {
    T* o = reinterpret_cast<T*>(out) + bl * D;
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = re[i] * re[i] + im[i] * im[i];
  } else {
    T p[D];
#pragma unroll
    for (int i = 0; i < D; ++i) p[i] = re[i] * re[i] + im[i] * im[i];
    T* o = reinterpret_cast<T*>(out) + bl * P.n_obs;
    for (int j = 0; j < P.n_obs; ++j) {
      const int mask = (int)P.obs[j].zmask;
      T acc = (T)0;
#pragma unroll
      for (int i = 0; i < D; ++i) acc += (__popc(i & mask) & 1) ? -p[i] : p[i];
      o[j] = acc;
    }
  }
}

}  // namespace qmlb
