// Device-side program view and matrix-source evaluation shared by all kernels.
// sm_100a only.  See include/qmlb200.h for the program semantics.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/qmlb200.h"

namespace qmlb {

template <typename T>
struct cx {
  T x, y;
};

template <typename T>
__host__ __device__ __forceinline__ cx<T> mk(T x, T y) {
  cx<T> r;
  r.x = x;
  r.y = y;
  return r;
}
template <typename T>
__device__ __forceinline__ cx<T> cmul(cx<T> a, cx<T> b) {
  return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// acc += a * b
template <typename T>
__device__ __forceinline__ void cfma(cx<T>& acc, cx<T> a, cx<T> b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}
template <typename T>
__device__ __forceinline__ cx<T> cconj(cx<T> a) {
  return mk<T>(a.x, -a.y);
}
template <typename T>
__device__ __forceinline__ cx<T> cadd(cx<T> a, cx<T> b) {
  return mk<T>(a.x + b.x, a.y + b.y);
}

// sin / cos of an angle that is always formed in double (c0 + sum coeff * arg).  complex64:
// the angle is reduced to [-pi, pi] in double BEFORE it is rounded to float - encodings
// scale inputs by 2^q / 3^q (ansaetze.py:933-961), and a float32 angle of several hundred
// radians carries an error of 1e-5 by itself.
__device__ __forceinline__ void sincos_angle(double a, double* s, double* c) { sincos(a, s, c); }
__device__ __forceinline__ void sincos_angle(double a, float* s, float* c) {
  const double two_pi = 6.283185307179586476925286766559;
  a = fma(-two_pi, rint(a * (1.0 / two_pi)), a);
  sincosf((float)a, s, c);
}

// Register kernel: per op, the hoisted factors its 2x2 matrix is made of when the source
// is one SRC_PRE or a chain of two (matrix = factor1 * factor0), else n = 0.
struct RegFast {
  int32_t n, slot0, local0, slot1, local1, pad0, pad1, pad2;
};

// Compact op of the register kernel (32 bytes), staged in shared memory by every CTA so
// that decoding an op costs shared-memory latency instead of two dependent global loads
// (op record, then its table descriptor) in front of every gate.  CX: b0 = control,
// b1 = target (resolved on the host).
struct RegOp {
  uint8_t kind, k, b0, b1;
  int32_t src;
  int32_t n, slot0, local0, slot1, local1;  // RegFast
  int32_t pad;
};
constexpr int REG_SMEM_OPS = 1024;

// Device copy of a program (all pointers are device pointers into one blob).
struct DevProg {
  const qmlb_op* ops;
  const qmlb_source* src;
  const int32_t* items;
  const qmlb_angle* ang;
  const qmlb_term* terms;
  const double* consts;
  const qmlb_obs* obs;
  const double* obs_consts;
  const qmlb_pre* pre;
  const RegOp* rops;  // strategy 0 only, else nullptr
  int32_t n_ops, n_obs, n_bits, n_qubits, density, out_type, n_pre, pad;
};

struct RunArgs {
  qmlb_arg a[QMLB_MAX_ARGS];
  int64_t batch;         // elements in this launch
  int64_t batch_offset;  // global index of element 0
  const void* pre_tab[QMLB_MAX_ARGS];  // per slot: [n_pre_slot][mod][4] complex, or unused
  int32_t pre_on[QMLB_MAX_ARGS];       // 1: read hoisted factors from pre_tab, 0: inline
};

// Row of argument `a` that batch element b reads.  Kernels that evaluate many
// angles per element cache these (the 64-bit div/mod is expensive).
struct RowsDirect {
  const RunArgs& R;
  int64_t b;
  __device__ __forceinline__ int64_t operator()(int a) const {
    return (b / R.a[a].div) % R.a[a].mod;
  }
};

template <typename Rows>
__device__ __forceinline__ const double* arg_row(const RunArgs& R, const Rows& rows, int arg) {
  return R.a[arg].ptr + rows(arg) * R.a[arg].stride;
}

template <typename Rows>
__device__ __forceinline__ double eval_angle(const DevProg& P, const RunArgs& R,
                                             const Rows& rows, int aid) {
  qmlb_angle a = P.ang[aid];
  double th = a.c0;
  for (int t = 0; t < a.n; ++t) {
    qmlb_term tm = P.terms[a.first + t];
    th = fma(tm.coeff, arg_row(R, rows, tm.arg)[tm.offset], th);
  }
  return th;
}

template <typename T>
__device__ __forceinline__ cx<T> ld_const(const double* consts, int off) {
  const double2 v = reinterpret_cast<const double2*>(consts)[off];
  return mk<T>((T)v.x, (T)v.y);
}

// 2x2 elementary source (CONST / TRIG / TABLE, k = 1) into registers.
template <typename T, typename Rows>
__device__ __forceinline__ void eval_elem2(const DevProg& P, const RunArgs& R,
                                           const Rows& rows, const qmlb_source& s,
                                           cx<T> m[4]) {
  if (s.kind == QMLB_SRC_TRIG) {
    T sn, cs;
    sincos_angle(eval_angle(P, R, rows, s.angle) * s.kappa, &sn, &cs);
    const int axis = (s.flags >> QMLB_FLAG_ROT_SHIFT) & 3;
    if (axis == 1) {  // RX
      m[0] = mk<T>(cs, 0); m[1] = mk<T>(0, -sn); m[2] = mk<T>(0, -sn); m[3] = mk<T>(cs, 0);
    } else if (axis == 2) {  // RY
      m[0] = mk<T>(cs, 0); m[1] = mk<T>(-sn, 0); m[2] = mk<T>(sn, 0); m[3] = mk<T>(cs, 0);
    } else if (axis == 3) {  // RZ
      m[0] = mk<T>(cs, -sn); m[1] = mk<T>(0, 0); m[2] = mk<T>(0, 0); m[3] = mk<T>(cs, sn);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        cx<T> c0 = ld_const<T>(P.consts, s.a0 + i);
        cx<T> a = ld_const<T>(P.consts, s.a1 + i);
        cx<T> bb = ld_const<T>(P.consts, s.a2 + i);
        m[i] = mk<T>(c0.x + cs * a.x + sn * bb.x, c0.y + cs * a.y + sn * bb.y);
      }
    }
  } else if (s.kind == QMLB_SRC_CONST) {
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = ld_const<T>(P.consts, s.a0 + i);
  } else {  // QMLB_SRC_TABLE
    const double2* row = reinterpret_cast<const double2*>(arg_row(R, rows, s.a0)) + s.a1;
    T sg = (s.flags & QMLB_FLAG_CONJ) ? (T)-1 : (T)1;
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = mk<T>((T)row[i].x, sg * (T)row[i].y);
  }
}

// m = f * m for 2x2 (row-major)
template <typename T>
__device__ __forceinline__ void mul2_left(const cx<T> f[4], cx<T> m[4]) {
  cx<T> r[4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      cx<T> acc = cmul(f[i * 2 + 0], m[0 * 2 + j]);
      cfma(acc, f[i * 2 + 1], m[1 * 2 + j]);
      r[i * 2 + j] = acc;
    }
#pragma unroll
  for (int i = 0; i < 4; ++i) m[i] = r[i];
}

// k = 1 source made of elementary items only (elementary itself or a CHAIN of them)
template <typename T, typename Rows>
__device__ __forceinline__ void eval_2x2_base(const DevProg& P, const RunArgs& R,
                                              const Rows& rows, int sid, cx<T> m[4]) {
  qmlb_source s = P.src[sid];
  if (s.kind != QMLB_SRC_CHAIN) {
    eval_elem2<T>(P, R, rows, s, m);
    return;
  }
  eval_elem2<T>(P, R, rows, P.src[P.items[s.a0]], m);
  for (int i = 1; i < s.a1; ++i) {
    cx<T> f[4];
    eval_elem2<T>(P, R, rows, P.src[P.items[s.a0 + i]], f);
    mul2_left<T>(f, m);
  }
}

// one chain item: elementary, or a hoisted factor (table lookup / inline)
template <typename T, typename Rows>
__device__ __forceinline__ void eval_item2(const DevProg& P, const RunArgs& R,
                                           const Rows& rows, const qmlb_source& s,
                                           cx<T> m[4]) {
  if (s.kind == QMLB_SRC_PRE) {
    if (R.pre_on[s.a1]) {
      const cx<T>* t = static_cast<const cx<T>*>(R.pre_tab[s.a1]) +
                       ((int64_t)s.a0 * R.a[s.a1].mod + rows(s.a1)) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) m[i] = t[i];
    } else {
      eval_2x2_base<T>(P, R, rows, P.pre[s.a2].src, m);
    }
  } else {
    eval_elem2<T>(P, R, rows, s, m);
  }
}

// Any k = 1 source into registers.
template <typename T, typename Rows>
__device__ __forceinline__ void eval_2x2(const DevProg& P, const RunArgs& R, const Rows& rows,
                                         int sid, cx<T> m[4]) {
  qmlb_source s = P.src[sid];
  if (s.kind != QMLB_SRC_CHAIN) {
    eval_item2<T>(P, R, rows, s, m);
    return;
  }
  eval_item2<T>(P, R, rows, P.src[P.items[s.a0]], m);
  for (int i = 1; i < s.a1; ++i) {
    cx<T> f[4];
    eval_item2<T>(P, R, rows, P.src[P.items[s.a0 + i]], f);
    mul2_left<T>(f, m);
  }
}

// Generic source into memory `out` (4^k entries row-major, or 2^k for diagonal
// sources).  Executed by ONE thread per source; `out` may be shared or local.
template <typename T, typename Rows>
__device__ void eval_source_mem(const DevProg& P, const RunArgs& R, const Rows& rows, int sid,
                                cx<T>* out) {
  qmlb_source s = P.src[sid];
  const int d = 1 << s.k;
  switch (s.kind) {
    case QMLB_SRC_CONST: {
      int n = (s.flags & QMLB_FLAG_DIAGVEC) ? d : d * d;
      for (int i = 0; i < n; ++i) out[i] = ld_const<T>(P.consts, s.a0 + i);
      break;
    }
    case QMLB_SRC_TRIG: {
      if (s.k == 1) {
        cx<T> m[4];
        eval_elem2<T>(P, R, rows, s, m);
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] = m[i];
        break;
      }
      T sn, cs;
      sincos_angle(eval_angle(P, R, rows, s.angle) * s.kappa, &sn, &cs);
      for (int i = 0; i < d * d; ++i) {
        cx<T> c0 = ld_const<T>(P.consts, s.a0 + i);
        cx<T> a = ld_const<T>(P.consts, s.a1 + i);
        cx<T> bb = ld_const<T>(P.consts, s.a2 + i);
        out[i] = mk<T>(c0.x + cs * a.x + sn * bb.x, c0.y + cs * a.y + sn * bb.y);
      }
      break;
    }
    case QMLB_SRC_PRE:
    case QMLB_SRC_CHAIN: {
      cx<T> m[4];
      eval_2x2<T>(P, R, rows, sid, m);
#pragma unroll
      for (int i = 0; i < 4; ++i) out[i] = m[i];
      break;
    }
    case QMLB_SRC_DIAGPH: {
      double th = eval_angle(P, R, rows, s.angle);
      for (int i = 0; i < d; ++i) {
        T sn, cs;
        sincos_angle(-P.consts[s.a0 + i] * th, &sn, &cs);
        out[i] = mk<T>(cs, sn);
      }
      break;
    }
    case QMLB_SRC_TABLE: {
      const double2* row = reinterpret_cast<const double2*>(arg_row(R, rows, s.a0)) + s.a1;
      T sg = (s.flags & QMLB_FLAG_CONJ) ? (T)-1 : (T)1;
      for (int i = 0; i < d * d; ++i) out[i] = mk<T>((T)row[i].x, sg * (T)row[i].y);
      break;
    }
    case QMLB_SRC_SUPER: {
      // S = prod_i item_i, item = U (x) conj(U) for a 2x2 source, or a constant 4x4
      cx<T> S[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) S[i] = mk<T>((i % 5 == 0) ? (T)1 : (T)0, (T)0);
      for (int it = 0; it < s.a1; ++it) {
        int id = P.items[s.a0 + it];
        cx<T> F[16];
        if (P.src[id].k == 2) {
#pragma unroll
          for (int i = 0; i < 16; ++i) F[i] = ld_const<T>(P.consts, P.src[id].a0 + i);
        } else {
          cx<T> u[4];
          eval_2x2<T>(P, R, rows, id, u);
          // (U (x) conj U)[(a,b),(c,d)] = U[a][c] * conj(U[b][d])
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int bb = 0; bb < 2; ++bb)
#pragma unroll
              for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int dd = 0; dd < 2; ++dd)
                  F[(a * 2 + bb) * 4 + (c * 2 + dd)] =
                      cmul(u[a * 2 + c], cconj(u[bb * 2 + dd]));
        }
        cx<T> N[16];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            cx<T> acc = mk<T>((T)0, (T)0);
#pragma unroll
            for (int l = 0; l < 4; ++l) cfma(acc, F[i * 4 + l], S[l * 4 + j]);
            N[i * 4 + j] = acc;
          }
#pragma unroll
        for (int i = 0; i < 16; ++i) S[i] = N[i];
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) out[i] = S[i];
      break;
    }
  }
}

// Precompute kernel: thread (j, idx) evaluates hoisted factor j of slot `slot` for
// row idx and stores it at tab[(local_j * mod + idx) * 4].
template <typename T>
__global__ void k_pre(DevProg P, RunArgs R, int slot, const int32_t* __restrict__ ids,
                      int n_ids, cx<T>* __restrict__ tab) {
  const int64_t mod = R.a[slot].mod;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= mod * n_ids) return;
  const int j = (int)(t / mod);
  const int64_t idx = t % mod;
  const qmlb_pre pe = P.pre[ids[j]];
  RowsDirect rows{R, idx * R.a[slot].div};
  cx<T> m[4];
  eval_2x2_base<T>(P, R, rows, pe.src, m);
  cx<T>* o = tab + ((int64_t)pe.local * mod + idx) * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = m[i];
}

// number of complex entries a source writes
__host__ __device__ __forceinline__ int source_entries(int kind, int k, int flags) {
  int d = 1 << k;
  if (kind == QMLB_SRC_DIAGPH) return d;
  if (kind == QMLB_SRC_CONST && (flags & QMLB_FLAG_DIAGVEC)) return d;
  if (kind == QMLB_SRC_PRE || kind == QMLB_SRC_CHAIN) return 4;
  return d * d;
}

// insert a zero bit at position `bit` of g
__device__ __forceinline__ uint32_t insert0(uint32_t g, int bit) {
  uint32_t low = g & ((1u << bit) - 1u);
  return ((g >> bit) << (bit + 1)) | low;
}

}  // namespace qmlb
