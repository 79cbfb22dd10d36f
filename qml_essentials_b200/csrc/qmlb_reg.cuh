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

#ifndef QMLB_REG_MIN_CTAS
#define QMLB_REG_MIN_CTAS 3
#endif

namespace qmlb {

// Register-resident state of 2^N amplitudes.  double: separate re / im arrays.  float:
// amplitudes 2k and 2k+1 share one float2 (an aligned 64-bit register pair), which is
// what lets the hot 2x2 updates issue packed fma.rn.f32x2 (FFMA2) on Blackwell.
template <typename T, int N>
struct RegState {
  T r[1 << N], i[1 << N];
  __device__ __forceinline__ T& re(int k) { return r[k]; }
  __device__ __forceinline__ T& im(int k) { return i[k]; }
};
template <int N>
struct RegState<float, N> {
  static constexpr int H = (1 << N) >= 2 ? (1 << N) / 2 : 1;
  float2 r2[H], i2[H];
  __device__ __forceinline__ float& re(int k) { return (k & 1) ? r2[k >> 1].y : r2[k >> 1].x; }
  __device__ __forceinline__ float& im(int k) { return (k & 1) ? i2[k >> 1].y : i2[k >> 1].x; }
};

// Exchange two state registers IN PLACE.  Written as opaque asm with read-write operands
// on purpose: a C++ swap is pure renaming for the compiler, and because every op variant
// sits in its own branch of the dispatch, each variant then ends with a different
// value->register map and the merge points pay for it by copying the whole state (ncu:
// 200-300 moves per CX plus 64 per loop iteration, half of all instructions).  With the
// swap done physically every branch leaves the state where it found it.
__device__ __forceinline__ void swap_in_place(double& a, double& b) {
  // three XORs on the bit patterns: real ALU work for ptxas (a mov-based swap is folded
  // back into renaming by its copy propagation)
  long long x = __double_as_longlong(a), y = __double_as_longlong(b);
  asm volatile("xor.b64 %0, %0, %1;" : "+l"(x) : "l"(y));
  asm volatile("xor.b64 %0, %0, %1;" : "+l"(y) : "l"(x));
  asm volatile("xor.b64 %0, %0, %1;" : "+l"(x) : "l"(y));
  a = __longlong_as_double(x);
  b = __longlong_as_double(y);
}
__device__ __forceinline__ void swap_in_place(float& a, float& b) {
  int x = __float_as_int(a), y = __float_as_int(b);
  asm volatile("xor.b32 %0, %0, %1;" : "+r"(x) : "r"(y));
  asm volatile("xor.b32 %0, %0, %1;" : "+r"(y) : "r"(x));
  asm volatile("xor.b32 %0, %0, %1;" : "+r"(x) : "r"(y));
  a = __int_as_float(x);
  b = __int_as_float(y);
}

template <int N, int BIT>
__device__ __forceinline__ constexpr int pair_i0(int g) {
  return ((g >> BIT) << (BIT + 1)) | (g & ((1 << BIT) - 1));
}

// Packed FP32 2x2 update (Blackwell fma.rn.f32x2 / FFMA2: two FMAs per issue slot).  For a
// target bit >= 1 the amplitudes 2k and 2k+1 play the same role, so one float2 of real
// parts and one of imaginary parts go through the update together against lane-broadcast
// matrix entries: 16 packed instructions per four amplitudes instead of 32 scalar ones.
template <int N>
__device__ __forceinline__ void f2_update(RegState<float, N>& S, int k0, int k1,
                                          const cx<float> (&m)[4]) {
  const float2 ar = S.r2[k0], ai = S.i2[k0], br = S.r2[k1], bi = S.i2[k1];
  auto bc = [](float v) { return make_float2(v, v); };
  float2 xr = __fmul2_rn(bc(m[0].x), ar);
  xr = __ffma2_rn(bc(-m[0].y), ai, xr);
  xr = __ffma2_rn(bc(m[1].x), br, xr);
  xr = __ffma2_rn(bc(-m[1].y), bi, xr);
  float2 xi = __fmul2_rn(bc(m[0].x), ai);
  xi = __ffma2_rn(bc(m[0].y), ar, xi);
  xi = __ffma2_rn(bc(m[1].x), bi, xi);
  xi = __ffma2_rn(bc(m[1].y), br, xi);
  float2 yr = __fmul2_rn(bc(m[2].x), ar);
  yr = __ffma2_rn(bc(-m[2].y), ai, yr);
  yr = __ffma2_rn(bc(m[3].x), br, yr);
  yr = __ffma2_rn(bc(-m[3].y), bi, yr);
  float2 yi = __fmul2_rn(bc(m[2].x), ai);
  yi = __ffma2_rn(bc(m[2].y), ar, yi);
  yi = __ffma2_rn(bc(m[3].x), bi, yi);
  yi = __ffma2_rn(bc(m[3].y), br, yi);
  S.r2[k0] = xr;
  S.i2[k0] = xi;
  S.r2[k1] = yr;
  S.i2[k1] = yi;
}

template <typename T, int N, int BIT>
__device__ __forceinline__ void reg_mat1(RegState<T, N>& S, const cx<T> (&m)[4]) {
  if constexpr (std::is_same<T, float>::value && BIT >= 1) {
#pragma unroll
    for (int h = 0; h < (1 << (N - 2)); ++h) {  // pairs of float2 that differ in bit BIT-1
      const int k0 = ((h >> (BIT - 1)) << BIT) | (h & ((1 << (BIT - 1)) - 1));
      f2_update<N>(S, k0, k0 | (1 << (BIT - 1)), m);
    }
  } else {
#pragma unroll
    for (int g = 0; g < (1 << (N - 1)); ++g) {
      const int i0 = pair_i0<N, BIT>(g), i1 = i0 | (1 << BIT);
      const T ar = S.re(i0), ai = S.im(i0), br = S.re(i1), bi = S.im(i1);
      S.re(i0) = m[0].x * ar - m[0].y * ai + m[1].x * br - m[1].y * bi;
      S.im(i0) = m[0].x * ai + m[0].y * ar + m[1].x * bi + m[1].y * br;
      S.re(i1) = m[2].x * ar - m[2].y * ai + m[3].x * br - m[3].y * bi;
      S.im(i1) = m[2].x * ai + m[2].y * ar + m[3].x * bi + m[3].y * br;
    }
  }
}

// 2x2 on TB where CB is set
template <typename T, int N, int CB, int TB>
__device__ __forceinline__ void reg_ctrl1(RegState<T, N>& S, const cx<T> (&m)[4]) {
  if constexpr (std::is_same<T, float>::value && CB >= 1 && TB >= 1) {
#pragma unroll
    for (int h = 0; h < (1 << (N - 2)); ++h) {
      const int k0 = ((h >> (TB - 1)) << TB) | (h & ((1 << (TB - 1)) - 1));
      if (k0 & (1 << (CB - 1))) f2_update<N>(S, k0, k0 | (1 << (TB - 1)), m);
    }
  } else {
#pragma unroll
    for (int g = 0; g < (1 << (N - 1)); ++g) {
      const int i0 = pair_i0<N, TB>(g), i1 = i0 | (1 << TB);
      if (i0 & (1 << CB)) {
        const T ar = S.re(i0), ai = S.im(i0), br = S.re(i1), bi = S.im(i1);
        S.re(i0) = m[0].x * ar - m[0].y * ai + m[1].x * br - m[1].y * bi;
        S.im(i0) = m[0].x * ai + m[0].y * ar + m[1].x * bi + m[1].y * br;
        S.re(i1) = m[2].x * ar - m[2].y * ai + m[3].x * br - m[3].y * bi;
        S.im(i1) = m[2].x * ai + m[2].y * ar + m[3].x * bi + m[3].y * br;
      }
    }
  }
}

// CX: swap the target pair where the control bit is set (pure register moves)
// INPLACE: exchange the registers physically (k_reg: measured 0.313 -> 0.289 ms on config 2,
// the merge-point copies disappear); the streaming kernel keeps the renaming form (its
// passes are short, and there the XOR form measured 17 % slower).
template <typename T>
__device__ __forceinline__ void swap_renamed(T& a, T& b) {
  const T t = a;
  a = b;
  b = t;
}

template <typename T, int N, int CB, int TB, bool INPLACE = false>
__device__ __forceinline__ void reg_cx(RegState<T, N>& S) {
#pragma unroll
  for (int g = 0; g < (1 << (N - 1)); ++g) {
    const int i0 = pair_i0<N, TB>(g), i1 = i0 | (1 << TB);
    if (i0 & (1 << CB)) {
      if constexpr (INPLACE) {
        swap_in_place(S.re(i0), S.re(i1));
        swap_in_place(S.im(i0), S.im(i1));
      } else {
        swap_renamed(S.re(i0), S.re(i1));
        swap_renamed(S.im(i0), S.im(i1));
      }
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

// CTA tile of the (fast axis x slow axis) batch, e.g. 16 parameter sets x 8 grid points:
// every hoisted factor the tile's 128 evaluations read (16 + 8 table rows per op instead of
// 128 + 1) is staged ONCE into shared memory with cp.async, so a gate's matrix costs
// shared-memory latency instead of an L2 round trip, no prefetch registers are needed, and
// the L2 -> SM traffic per evaluation drops ~7x (config 2: 30 KB per CTA).
//   cls[slot]: 0 = slot has no hoisted factors here, 1 = FAST (row = fast index p, div = 1),
//              2 = SLOW (row = slow index i = b / BP), 3 = CONST (mod = 1: row 0)
struct RegTile {
  int32_t tp_bits, ti_bits;  // fast / slow rows per CTA tile (log2); 2^(tp_bits + ti_bits) =
                             // 128 * reps: a thread runs `reps` evaluations (slow rows
                             // ti, ti + 2^ti_bits / reps, ...) against the same staged tables
  int32_t reps, pad;
  int64_t n_ptiles, BP, BI;
  uint8_t cls[QMLB_MAX_ARGS];
};

// The op stream in the KERNEL PARAMETERS (constant bank) for programs of up to REG_PARAM_OPS
// ops: op records are then uniform data, the dispatch on kind / bit becomes uniform
// branches (no BSSY / BSYNC reconvergence code, no shared-memory load in front of every
// gate).  PO = 1 variants of k_reg take the ops from here instead of the staged copy.
constexpr int REG_PARAM_OPS = 72;
struct RegParamOps {
  RegOp ops[REG_PARAM_OPS];
};

template <typename T>
__device__ __forceinline__ void reg_cp_async(cx<T>* smem, const cx<T>* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  if constexpr (sizeof(cx<T>) == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem));
}

// shared-memory stride of one staged 2x2 factor in cx<T> units: 4 entries + 16 bytes of
// padding, which spreads the rows of a quarter warp over all banks (80 / 48 byte stride)
template <typename T>
__host__ __device__ constexpr int reg_tile_stride() {
  return sizeof(T) == 8 ? 5 : 6;
}

// rows cached in shared memory: slot a of this thread at base[a * 128]
struct RowsShared {
  const int32_t* base;
  __device__ __forceinline__ int64_t operator()(int a) const { return base[a * 128]; }
};

// mode: 0 -> write state, 1 -> probs, 2 -> Z-string expectation values
// Resident CTAs per SM the kernel is compiled for: the state takes 2 * 2^N * sizeof(T) / 4
// registers; up to 64 of them (c128 n <= 4, c64 n <= 5) leave room for 3 CTAs (168
// registers / thread), which is what hides the table-lookup latency.
template <typename T, int N>
constexpr int reg_min_ctas() {
  return (2 * (1 << N) * (int)sizeof(T) / 4 <= 64) ? (sizeof(T) == 8 ? 2 : QMLB_REG_MIN_CTAS) : 1;
}

// MINB: resident CTAs per SM the variant is compiled for (0 = reg_min_ctas default).  The
// complex128 n = 4 kernel exists at 2 (210 registers, no spills) and 3 (168 registers, 124
// bytes of spills, 12 instead of 8 warps per SM).
template <typename T, int N, int MINB = 0, bool TILED = false, bool PO = false>
__global__ void __launch_bounds__(128, MINB ? MINB : reg_min_ctas<T, N>())
    k_reg(DevProg P, RunArgs R, int mode, int n_args, void* __restrict__ out, const RegTile Tl,
          const __grid_constant__ RegParamOps Po) {
  constexpr int D = 1 << N;
  // op stream -> shared memory (all threads take part before anyone leaves)
  extern __shared__ __align__(16) unsigned char reg_smem[];
  RegOp* s_ops = reinterpret_cast<RegOp*>(reg_smem);
  const bool staged = P.n_ops <= REG_SMEM_OPS;
  if (staged) {
    const int4* src4 = reinterpret_cast<const int4*>(P.rops);
    int4* dst4 = reinterpret_cast<int4*>(s_ops);
    for (int i = threadIdx.x; i < P.n_ops * 2; i += blockDim.x) dst4[i] = src4[i];
    __syncthreads();
  }
  const RegOp* ops = staged ? s_ops : P.rops;
  int64_t bl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // TILED: [ops | 2 offsets per op | staged factor tables]
  constexpr int ES = reg_tile_stride<T>();
  int32_t* s_off = reinterpret_cast<int32_t*>(s_ops + P.n_ops);
  cx<T>* s_tab = reinterpret_cast<cx<T>*>(
      reg_smem + ((sizeof(RegOp) * P.n_ops + 8 * (size_t)P.n_ops + 15) & ~size_t(15)));
  int row_f = 0, row_s = 0;  // this thread's rows in the staged FAST / SLOW factors
  int rows_f = 0, rows_s = 0, ti_step = 0, n_rep = 1;
  int64_t p0 = 0, i0 = 0;
  if constexpr (TILED) {
    const int TP = 1 << Tl.tp_bits, TI = 1 << Tl.ti_bits;
    const int64_t cp = (int64_t)blockIdx.x % Tl.n_ptiles, ci = (int64_t)blockIdx.x / Tl.n_ptiles;
    p0 = cp * TP;
    i0 = ci * TI;
    rows_f = (int)(Tl.BP - p0 < TP ? Tl.BP - p0 : TP);
    rows_s = (int)(Tl.BI - i0 < TI ? Tl.BI - i0 : TI);
    n_rep = Tl.reps;
    ti_step = TI / Tl.reps;
    row_f = threadIdx.x & (TP - 1);
    if (threadIdx.x == 0) {
      int acc = 0;
      for (int o = 0; o < P.n_ops; ++o) {
        const RegOp d = PO ? Po.ops[o] : s_ops[o];
        const bool use = d.n >= 1 && R.pre_on[d.slot0] && (d.n == 1 || R.pre_on[d.slot1]);
        for (int j = 0; j < 2; ++j) {
          int off = -1;
          if (use && j < d.n) {
            const int c = Tl.cls[j ? d.slot1 : d.slot0];
            off = acc;
            acc += (c == 1 ? rows_f : (c == 2 ? rows_s : 1)) * ES;
          }
          s_off[2 * o + j] = off;
        }
      }
    }
    __syncthreads();
    for (int o = 0; o < P.n_ops; ++o) {
      const RegOp d = PO ? Po.ops[o] : s_ops[o];
      for (int j = 0; j < 2; ++j) {
        const int off = s_off[2 * o + j];
        if (off < 0) continue;
        const int slot = j ? d.slot1 : d.slot0, local = j ? d.local1 : d.local0;
        const int c = Tl.cls[slot];
        const int rows = c == 1 ? rows_f : (c == 2 ? rows_s : 1);
        const int64_t row0 = c == 1 ? p0 : (c == 2 ? i0 : 0);
        const cx<T>* src = static_cast<const cx<T>*>(R.pre_tab[slot]) +
                           ((int64_t)local * R.a[slot].mod + row0) * 4;
        cx<T>* dst = s_tab + off;
        for (int e = threadIdx.x; e < rows * 4; e += 128)
          reg_cp_async<T>(dst + (e >> 2) * ES + (e & 3), src + e);
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();
    if (row_f >= rows_f) return;
  } else {
    if (bl >= R.batch) return;
  }
  // TILED: `reps` evaluations per thread against the same staged tables (the staging
  // prologue - two barriers and an L2 round trip - is paid once for all of them)
#pragma unroll 1
  for (int rep = 0; rep < n_rep; ++rep) {
  if constexpr (TILED) {
    row_s = (int)(threadIdx.x >> Tl.tp_bits) + rep * ti_step;
    if (row_s >= rows_s) break;
    bl = (i0 + row_s) * Tl.BP + p0 + row_f;
  }
  const int64_t b = bl + R.batch_offset;

  // rows of the argument slots this element reads, computed once (64-bit div/mod)
  __shared__ int32_t s_rows[QMLB_MAX_ARGS][128];
#pragma unroll
  for (int a = 0; a < QMLB_MAX_ARGS; ++a)
    if (a < n_args) s_rows[a][threadIdx.x] = (int32_t)((b / R.a[a].div) % R.a[a].mod);
  const RowsShared rows{&s_rows[0][threadIdx.x]};

  RegState<T, N> S;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    S.re(i) = (T)0;
    S.im(i) = (T)0;
  }
  S.re(0) = (T)1;

  // Software pipeline over the op stream: the hoisted-factor table rows of op o+1 are
  // requested (plain loads into registers) before op o is applied, so their L2 latency
  // overlaps the arithmetic instead of stalling every gate (ncu: long-scoreboard was the
  // top stall with 8-12 resident warps per SM).
  cx<T> nf0[4], nf1[4];
  int nfn = 0;
  auto fetch = [&](int o) {
    nfn = 0;
    if constexpr (TILED) return;  // factors come from shared memory, no prefetch registers
    if (o >= P.n_ops) return;
    const RegOp d = ops[o];
    if (d.n < 1 || !R.pre_on[d.slot0] || (d.n == 2 && !R.pre_on[d.slot1])) return;
    nfn = d.n;
    const cx<T>* t0 = static_cast<const cx<T>*>(R.pre_tab[d.slot0]) +
                      ((int64_t)d.local0 * R.a[d.slot0].mod + rows(d.slot0)) * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) nf0[i] = t0[i];
    if (d.n == 2) {
      const cx<T>* t1 = static_cast<const cx<T>*>(R.pre_tab[d.slot1]) +
                        ((int64_t)d.local1 * R.a[d.slot1].mod + rows(d.slot1)) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) nf1[i] = t1[i];
    }
  };
  fetch(0);

  for (int o = 0; o < P.n_ops; ++o) {
    const RegOp op = PO ? Po.ops[o] : ops[o];
    cx<T> m[4];
    bool have_m = nfn > 0;
    if constexpr (TILED) {
      const int off0 = s_off[2 * o], off1 = s_off[2 * o + 1];
      have_m = off0 >= 0;
      if (have_m) {
        const int c0 = Tl.cls[op.slot0];
        const cx<T>* e0 = s_tab + off0 + (c0 == 1 ? row_f : (c0 == 2 ? row_s : 0)) * ES;
#pragma unroll
        for (int i = 0; i < 4; ++i) m[i] = e0[i];
        if (off1 >= 0) {
          const int c1 = Tl.cls[op.slot1];
          const cx<T>* e1 = s_tab + off1 + (c1 == 1 ? row_f : (c1 == 2 ? row_s : 0)) * ES;
          cx<T> f1[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) f1[i] = e1[i];
          mul2_left<T>(f1, m);
        }
      }
    } else {
      if (have_m) {
#pragma unroll
        for (int i = 0; i < 4; ++i) m[i] = nf0[i];
        if (nfn == 2) mul2_left<T>(nf1, m);
      }
      fetch(o + 1);
    }
    if (op.kind == QMLB_OP_PERM) {
      if (op.k == 1) {  // the only non-identity 1-bit permutation is X
        dispatch1<T, N>(op.b0, [&](auto B) {
          constexpr int BIT = decltype(B)::value;
#pragma unroll
          for (int g = 0; g < (1 << (N - 1)); ++g) {
            const int i0 = pair_i0<N, BIT>(g), i1 = i0 | (1 << BIT);
            swap_in_place(S.re(i0), S.re(i1));
            swap_in_place(S.im(i0), S.im(i1));
          }
        });
      } else {
        // CX: b0 = control, b1 = target (orientation resolved at upload)
        dispatch2<T, N>(op.b0, op.b1, [&](auto CB, auto TB) {
          reg_cx<T, N, decltype(CB)::value, decltype(TB)::value, true>(S);
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
          sincos_angle(-P.consts[s.a0 + i] * th, &sn, &cs);
          const T r = S.re(i), q = S.im(i);
          S.re(i) = cs * r - sn * q;
          S.im(i) = cs * q + sn * r;
        }
      } else {
#pragma unroll
        for (int i = 0; i < D; ++i) {
          const cx<T> c = ld_const<T>(P.consts, s.a0 + i);
          const T r = S.re(i), q = S.im(i);
          S.re(i) = c.x * r - c.y * q;
          S.im(i) = c.x * q + c.y * r;
        }
      }
      continue;
    }
    // remaining kinds take one 2x2 matrix
    if (op.kind == QMLB_OP_DIAG) {
      const qmlb_source s = P.src[op.src];
      m[1] = mk<T>(0, 0);
      m[2] = mk<T>(0, 0);
      if (s.kind == QMLB_SRC_DIAGPH) {
        const double th = eval_angle(P, R, rows, s.angle);
        T sn, cs;
        sincos_angle(-P.consts[s.a0] * th, &sn, &cs);
        m[0] = mk<T>(cs, sn);
        sincos_angle(-P.consts[s.a0 + 1] * th, &sn, &cs);
        m[3] = mk<T>(cs, sn);
      } else {
        m[0] = ld_const<T>(P.consts, s.a0);
        m[3] = ld_const<T>(P.consts, s.a0 + 1);
      }
    } else if (!have_m) {
      eval_2x2<T>(P, R, rows, op.src, m);
    }
    if (op.kind == QMLB_OP_CTRL1) {
      dispatch2<T, N>(op.b0, op.b1, [&](auto CB, auto TB) {
        reg_ctrl1<T, N, decltype(CB)::value, decltype(TB)::value>(S, m);
      });
    } else {
      dispatch1<T, N>(op.b0, [&](auto B) {
        reg_mat1<T, N, decltype(B)::value>(S, m);
      });
    }
  }

  if (mode == 0) {
    cx<T>* o = reinterpret_cast<cx<T>*>(out) + bl * D;
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = mk<T>(S.re(i), S.im(i));
  } else if (mode == 1) {
    T* o = reinterpret_cast<T*>(out) + bl * D;
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = S.re(i) * S.re(i) + S.im(i) * S.im(i);
  } else {
    T p[D];
#pragma unroll
    for (int i = 0; i < D; ++i) p[i] = S.re(i) * S.re(i) + S.im(i) * S.im(i);
    T* o = reinterpret_cast<T*>(out) + bl * P.n_obs;
    for (int j = 0; j < P.n_obs; ++j) {
      const int mask = (int)P.obs[j].zmask;
      T acc = (T)0;
#pragma unroll
      for (int i = 0; i < D; ++i) acc += (__popc(i & mask) & 1) ? -p[i] : p[i];
      o[j] = acc;
    }
  }
  }  // rep
}

}  // namespace qmlb
