// Measurement kernels over states resident in global memory
// (simulation.py:204-317 measure_state / measure_density, jaqsi-level purities and
// pair fidelities).  Reductions use a fixed order (warp shuffle tree, then warps in
// index order, then chunk partials in index order) so results are bitwise
// reproducible and independent of how a batch is sharded across GPUs.
#pragma once

#include "qmlb_device.cuh"

namespace qmlb {

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// sum over the CTA; result valid in thread 0.  `red` holds >= 32 T.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  T r = (T)0;
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) r += red[i];
  }
  return r;
}

// probabilities: |psi_i|^2 or Re rho_ii
template <typename T>
__global__ void k_probs(const cx<T>* __restrict__ st, T* __restrict__ out, int64_t batch,
                        int n_qubits, int density) {
  const int64_t dim = (int64_t)1 << n_qubits;
  const int64_t total = batch * dim;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (density) {
      const int64_t b = i / dim, j = i % dim;
      out[i] = st[b * dim * dim + j * (dim + 1)].x;
    } else {
      const cx<T> a = st[i];
      out[i] = a.x * a.x + a.y * a.y;
    }
  }
}

// rho = |psi><psi| for noise-free circuits asked for "density" (simulation.py:188-189)
template <typename T>
__global__ void k_outer(const cx<T>* __restrict__ st, cx<T>* __restrict__ out,
                        int64_t batch, int n_qubits) {
  const int64_t dim = (int64_t)1 << n_qubits;
  const int64_t total = batch * dim * dim;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / (dim * dim), r = (i / dim) % dim, c = i % dim;
    const cx<T> x = st[b * dim + r], y = st[b * dim + c];
    out[i] = mk<T>(x.x * y.x + x.y * y.y, x.y * y.x - x.x * y.y);
  }
}

#define QMLB_OBS_GROUP 8

// Expectation values.  Grid: batch * chunks CTAs; CTA (bl, c) reduces chunk c of
// element bl for every observable and writes partial[(bl * n_obs + j) * chunks + c]
// (directly the result when chunks == 1).
template <typename T>
__global__ void __launch_bounds__(256) k_expval(DevProg P, const cx<T>* __restrict__ st,
                                                T* __restrict__ partial, int64_t batch,
                                                int chunks) {
  __shared__ T red[32];
  const int n = P.n_qubits;
  const int64_t dim = (int64_t)1 << n;
  const int64_t bl = blockIdx.x / chunks;
  const int c = blockIdx.x % chunks;
  const cx<T>* s = st + (size_t)bl * (P.density ? dim * dim : dim);

  auto prob = [&](int64_t i) -> T {
    if (P.density) return s[i * (dim + 1)].x;
    const cx<T> a = s[i];
    return a.x * a.x + a.y * a.y;
  };

  int j = 0;
  while (j < P.n_obs) {
    if (P.obs[j].kind != QMLB_OBS_DENSE) {
      // group of up to 8 consecutive diagonal-type observables: one sweep
      int cnt = 0;
      while (j + cnt < P.n_obs && cnt < QMLB_OBS_GROUP &&
             P.obs[j + cnt].kind != QMLB_OBS_DENSE)
        ++cnt;
      T acc[QMLB_OBS_GROUP];
#pragma unroll
      for (int o = 0; o < QMLB_OBS_GROUP; ++o) acc[o] = (T)0;
      const int64_t per = (dim + chunks - 1) / chunks;
      const int64_t lo = c * per, hi = (lo + per < dim) ? lo + per : dim;
      for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const T p = prob(i);
#pragma unroll
        for (int o = 0; o < QMLB_OBS_GROUP; ++o) {
          if (o < cnt) {
            const qmlb_obs& ob = P.obs[j + o];
            T d;
            if (ob.kind == QMLB_OBS_ZSTRING) {
              d = (__popcll((unsigned long long)(i & ob.zmask)) & 1) ? (T)-1 : (T)1;
            } else {
              int v = 0;
              for (int t = 0; t < ob.k; ++t) v |= (int)((i >> ob.bits[t]) & 1) << (ob.k - 1 - t);
              d = (T)P.obs_consts[2 * (ob.a0 + v)];
            }
            acc[o] = fma(p, d, acc[o]);
          }
        }
      }
#pragma unroll
      for (int o = 0; o < QMLB_OBS_GROUP; ++o) {
        if (o < cnt) {
          const T r = block_sum<T>(acc[o], red);
          if (threadIdx.x == 0) partial[((size_t)bl * P.n_obs + j + o) * chunks + c] = r;
        }
      }
      j += cnt;
    } else {
      // dense k-qubit observable: sum over the n-k untouched bits of
      //   pure : sum_{r,c} conj(psi[r]) O[r][c] psi[c]
      //   mixed: sum_{a,b} O[a][b] rho[idx(b)][idx(a)]          (Tr(O rho))
      const qmlb_obs& ob = P.obs[j];
      const int k = ob.k, d = 1 << k;
      int sorted[QMLB_MAX_OP_BITS];
      for (int t = 0; t < k; ++t) sorted[t] = ob.bits[t];
      for (int a = 0; a < k; ++a)
        for (int t = 0; t + 1 < k - a; ++t)
          if (sorted[t] > sorted[t + 1]) {
            int x = sorted[t];
            sorted[t] = sorted[t + 1];
            sorted[t + 1] = x;
          }
      const int64_t groups = dim >> k;
      const int64_t per = (groups + chunks - 1) / chunks;
      const int64_t lo = c * per, hi = (lo + per < groups) ? lo + per : groups;
      T acc = (T)0;
      for (int64_t g = lo + threadIdx.x; g < hi; g += blockDim.x) {
        int64_t base = g;
        for (int t = 0; t < k; ++t) {
          const int64_t low = base & (((int64_t)1 << sorted[t]) - 1);
          base = ((base >> sorted[t]) << (sorted[t] + 1)) | low;
        }
        for (int r = 0; r < d; ++r) {
          int64_t ir = base;
          for (int t = 0; t < k; ++t) ir |= (int64_t)((r >> (k - 1 - t)) & 1) << ob.bits[t];
          for (int cc = 0; cc < d; ++cc) {
            int64_t ic = base;
            for (int t = 0; t < k; ++t)
              ic |= (int64_t)((cc >> (k - 1 - t)) & 1) << ob.bits[t];
            const T oR = (T)P.obs_consts[2 * (ob.a0 + r * d + cc)];
            const T oI = (T)P.obs_consts[2 * (ob.a0 + r * d + cc) + 1];
            if (oR == (T)0 && oI == (T)0) continue;
            if (P.density) {
              const cx<T> x = s[ic * dim + ir];  // rho[c][r]
              acc += oR * x.x - oI * x.y;
            } else {
              const cx<T> x = s[ir], y = s[ic];
              // Re( conj(x) * O * y )
              const T tr = oR * y.x - oI * y.y, ti = oR * y.y + oI * y.x;
              acc += x.x * tr + x.y * ti;
            }
          }
        }
      }
      const T r = block_sum<T>(acc, red);
      if (threadIdx.x == 0) partial[((size_t)bl * P.n_obs + j) * chunks + c] = r;
      ++j;
    }
  }
}

// <Z_q> for every qubit of pure states in ONE sweep over the state (the generic kernel
// above re-reads the state per group of 8 observables and spends ~20 instructions per
// amplitude on signs).  A warp takes units of 256 consecutive amplitudes: lane l loads
// amplitudes u*256 + j*32 + l (j = 0..7, coalesced).  The bit of each qubit is then either
// a lane bit (0-4: resolved once at the end), an in-thread bit (5-7: three running sums) or
// a bit of the unit index u (>= 8: one predicated add per unit).  Accumulation is in
// double; the reduction order is fixed (thread -> warp tree -> warps in order -> CTAs in
// order), so results do not depend on timing.
// partial[(bl * ctas + c) * 33 + q] = sum of p over amplitudes with bit q set (q < n),
// partial[... + 32] = total probability.   Grid: (ctas, batch).
template <typename T>
__global__ void __launch_bounds__(256) k_expval_z1(const cx<T>* __restrict__ st,
                                                   double* __restrict__ partial, int n) {
  __shared__ double red[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t units = 1ull << (n - 8);
  const cx<T>* s = st + ((size_t)blockIdx.y << n);
  double tot = 0, a5 = 0, a6 = 0, a7 = 0;
  double H[24];
#pragma unroll
  for (int q = 0; q < 24; ++q) H[q] = 0;
  for (uint64_t u = (uint64_t)blockIdx.x * 8 + warp; u < units; u += (uint64_t)gridDim.x * 8) {
    const cx<T>* p = s + (u << 8) + lane;
    T pr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const cx<T> a = p[j * 32];
      pr[j] = a.x * a.x + a.y * a.y;
    }
    const double s5 = (double)(pr[1] + pr[3]) + (double)(pr[5] + pr[7]);
    const double s6 = (double)(pr[2] + pr[3]) + (double)(pr[6] + pr[7]);
    const double s7 = (double)(pr[4] + pr[5]) + (double)(pr[6] + pr[7]);
    const double t = (double)(pr[0] + pr[1]) + (double)(pr[2] + pr[3]) + s7;
    tot += t;
    a5 += s5;
    a6 += s6;
    a7 += s7;
#pragma unroll
    for (int q = 0; q < 24; ++q)
      if (q + 8 < n && ((u >> q) & 1ull)) H[q] += t;
  }
  // per-thread vector -> CTA sums
#pragma unroll
  for (int q = 0; q < 33; ++q) {
    double v;
    if (q < 5) v = ((lane >> q) & 1) ? tot : 0.0;
    else if (q == 5) v = a5;
    else if (q == 6) v = a6;
    else if (q == 7) v = a7;
    else if (q < 32) v = H[q - 8];
    else v = tot;
    v = warp_sum(v);
    if (lane == 0) red[warp][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < 33) {
    double r = 0;
    for (int w = 0; w < 8; ++w) r += red[w][threadIdx.x];
    partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 33 + threadIdx.x] = r;
  }
}

// out[bl][j] = total - 2 * S_{bit(j)} summed over CTAs in index order
// where each logical state bit lives in memory (identity unless a streamed tile program left
// the state in a permuted bit order)
struct BitMap {
  int8_t pos[40];
};

template <typename T>
__global__ void k_expval_z1_final(DevProg P, const double* __restrict__ partial,
                                  T* __restrict__ out, int64_t batch, int ctas,
                                  const BitMap map) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * P.n_obs) return;
  const int64_t bl = i / P.n_obs;
  const int j = (int)(i % P.n_obs);
  const int bit = map.pos[63 - __clzll((unsigned long long)P.obs[j].zmask)];
  double tot = 0, sq = 0;
  for (int c = 0; c < ctas; ++c) {
    const double* row = partial + ((size_t)bl * ctas + c) * 33;
    tot += row[32];
    sq += row[bit];
  }
  out[i] = (T)(tot - 2.0 * sq);
}

// out[bl][q] = sum over CTAs (index order) of partial[(bl * ctas + c) * 33 + q]
__global__ void k_zsums_final(const double* __restrict__ partial, double* __restrict__ out,
                              int64_t batch, int ctas) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * 33) return;
  const int64_t bl = i / 33;
  const int q = (int)(i % 33);
  double r = 0;
  for (int c = 0; c < ctas; ++c) r += partial[((size_t)bl * ctas + c) * 33 + q];
  out[i] = r;
}

// out[i] = sum_c partial[i * chunks + c] in index order
template <typename T>
__global__ void k_sum_chunks(const T* __restrict__ partial, T* __restrict__ out,
                             int64_t n, int chunks) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T r = (T)0;
  for (int c = 0; c < chunks; ++c) r += partial[i * chunks + c];
  out[i] = r;
}

// Shot bookkeeping (simulation.py:352-357): one CTA per element.
template <typename T>
__global__ void __launch_bounds__(256) k_sample(const T* __restrict__ probs,
                                                const double* __restrict__ uniforms,
                                                int n_qubits, int64_t shots,
                                                int32_t* __restrict__ counts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* cum = reinterpret_cast<T*>(smem_raw);
  const int dim = 1 << n_qubits;
  const int64_t b = blockIdx.x;
  const T* p = probs + b * dim;
  if (threadIdx.x == 0) {  // sequential order: matches the oracle's cumsum bit for bit
    T run = (T)0;
    for (int i = 0; i < dim; ++i) {
      run += p[i];
      cum[i] = run;
    }
  }
  __syncthreads();
  const T total = cum[dim - 1];
  int32_t* cnt = counts + b * dim;
  const double* u = uniforms + b * shots;
  for (int64_t sidx = threadIdx.x; sidx < shots; sidx += blockDim.x) {
    const T r = total * ((T)1 - (T)u[sidx]);
    int lo = 0, hi = dim;  // first index with cum[idx] >= r  (searchsorted, side=left)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cum[mid] < r)
        lo = mid + 1;
      else
        hi = mid;
    }
    if (lo > dim - 1) lo = dim - 1;
    atomicAdd(&cnt[lo], 1);
  }
}

// Meyer-Wallach purities (entanglement.py:86-103): out[b][q] = Tr[(Tr_q rho)^2], the
// purity of the state with qubit q traced out.
//  * pure states: equals the purity of the single-qubit reduced state of qubit q
//    (Schmidt decomposition), which is a 2x2 reduction over the statevector;
//  * density matrices: sum over (a, c) in the remaining register of
//    |rho[(a,0),(c,0)] + rho[(a,1),(c,1)]|^2.
// One CTA per element.
template <typename T>
__global__ void __launch_bounds__(256) k_purity(const cx<T>* __restrict__ st, int density,
                                                int n_qubits, T* __restrict__ out) {
  __shared__ T red[32];
  const int64_t dim = (int64_t)1 << n_qubits;
  const int64_t b = blockIdx.x;
  const cx<T>* s = st + (size_t)b * (density ? dim * dim : dim);
  for (int q = 0; q < n_qubits; ++q) {
    const int bit = n_qubits - 1 - q;
    const int64_t lowmask = ((int64_t)1 << bit) - 1;
    if (density) {
      const int64_t half = dim / 2;
      T acc = 0;
      for (int64_t g = threadIdx.x; g < half * half; g += blockDim.x) {
        const int64_t ga = g / half, gc = g % half;
        const int64_t a0 = ((ga >> bit) << (bit + 1)) | (ga & lowmask);
        const int64_t c0 = ((gc >> bit) << (bit + 1)) | (gc & lowmask);
        const int64_t a1 = a0 | ((int64_t)1 << bit), c1 = c0 | ((int64_t)1 << bit);
        const cx<T> x = s[a0 * dim + c0], y = s[a1 * dim + c1];
        const T re = x.x + y.x, im = x.y + y.y;
        acc += re * re + im * im;
      }
      acc = block_sum<T>(acc, red);
      if (threadIdx.x == 0) out[b * n_qubits + q] = acc;
    } else {
      // reduced 2x2 of qubit q: r00, r11 real, r01 complex
      T r00 = 0, r11 = 0, xr = 0, xi = 0;
      for (int64_t g = threadIdx.x; g < dim / 2; g += blockDim.x) {
        const int64_t i0 = ((g >> bit) << (bit + 1)) | (g & lowmask);
        const int64_t i1 = i0 | ((int64_t)1 << bit);
        const cx<T> a = s[i0], c = s[i1];
        r00 += a.x * a.x + a.y * a.y;
        r11 += c.x * c.x + c.y * c.y;
        xr += a.x * c.x + a.y * c.y;  // a * conj(c)
        xi += a.y * c.x - a.x * c.y;
      }
      r00 = block_sum<T>(r00, red);
      r11 = block_sum<T>(r11, red);
      xr = block_sum<T>(xr, red);
      xi = block_sum<T>(xi, red);
      if (threadIdx.x == 0)
        out[b * n_qubits + q] = r00 * r00 + r11 * r11 + (T)2 * (xr * xr + xi * xi);
    }
  }
}

// |<psi_b|psi_{b+half}>|^2 ; one CTA per pair.
template <typename T>
__global__ void __launch_bounds__(256) k_overlap(const cx<T>* __restrict__ st, int64_t half,
                                                 int n_qubits, T* __restrict__ out) {
  __shared__ T red[32];
  const int64_t dim = (int64_t)1 << n_qubits;
  const int64_t b = blockIdx.x;
  const cx<T>* x = st + (size_t)b * dim;
  const cx<T>* y = st + (size_t)(b + half) * dim;
  T re = 0, im = 0;
  for (int64_t i = threadIdx.x; i < dim; i += blockDim.x) {
    const cx<T> a = x[i], c = y[i];
    re += a.x * c.x + a.y * c.y;
    im += a.x * c.y - a.y * c.x;
  }
  re = block_sum<T>(re, red);
  im = block_sum<T>(im, red);
  if (threadIdx.x == 0) out[b] = re * re + im * im;
}

// FMA-throughput probe (roofline denominator of the register regime)
template <typename T>
__global__ void __launch_bounds__(256) k_fma_peak(T* out, int iters) {
  T a0 = (T)threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
    a6 = a0 + 6, a7 = a0 + 7;
  const T m = (T)0.999999, c = (T)1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace qmlb
