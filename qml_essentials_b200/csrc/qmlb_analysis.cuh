// Device post-processing of the Fourier-analysis callers: the
// grid DFT of Coefficients._fourier_transform (coefficients.py:128-150), the additive
// sufficient statistics the FCC correlation estimators need (coefficients.py:1346-1498), and
// a one-shot all-reduce of those statistics over peer-mapped (symmetric) memory - so that a
// Fourier-fingerprint step returns kilobytes and crosses ranks without a library collective.
#pragma once

#include "qmlb_device.cuh"

namespace qmlb {

constexpr int DFT_PCOLS = 8;  // parameter samples per CTA
constexpr int DFT_TC = 4;     // ... of which one thread sums four (one twiddle load feeds 8 FMAs)

// out[row_of[k]][p] = (1 / n_x) * sum_x s[x][p] * exp(-2 pi i k x / n_x),
// s[x][p] = mean over the n_obs observables of ev[x][p][:]   (ev: (n_x, n_p, n_obs) real)
// One CTA = DFT_PCOLS samples x all frequencies; the signal tile and the twiddle table sit
// in shared memory (double precision throughout).  The signal is real: only k <= n_x / 2
// is summed (coefficient n_x - k is written as the conjugate), and grid points x and
// n_x - x are folded first (E = s[x] + s[n-x] meets the cosine, O = s[x] - s[n-x] the sine),
// which halves the inner loop; x is summed in index order.  row_of (optional) places
// frequency k on an output row of the caller's choice (-1 = not wanted): get_spectrum's
// shift / trim cost nothing.
template <typename T>
__global__ void __launch_bounds__(1024) k_grid_dft(const T* __restrict__ ev, int n_x, int64_t n_p,
                                                  int n_obs, const int32_t* __restrict__ row_of,
                                                  cx<T>* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char dsm[];
  double* sig = reinterpret_cast<double*>(dsm);                             // [n_x][DFT_PCOLS]
  double2* tw = reinterpret_cast<double2*>(sig + (size_t)n_x * DFT_PCOLS);  // [n_x]
  const int64_t p0 = (int64_t)blockIdx.x * DFT_PCOLS;
  for (int i = threadIdx.x; i < n_x * DFT_PCOLS; i += blockDim.x) {
    const int x = i / DFT_PCOLS, c = i % DFT_PCOLS;
    double m = 0.0;
    if (p0 + c < n_p) {
      const T* row = ev + ((size_t)x * n_p + (p0 + c)) * n_obs;
      for (int q = 0; q < n_obs; ++q) m += (double)row[q];
      m /= (double)n_obs;
    }
    sig[i] = m;
  }
  for (int j = threadIdx.x; j < n_x; j += blockDim.x) {
    double s, c;
    sincospi(-2.0 * (double)j / (double)n_x, &s, &c);
    tw[j] = make_double2(c, s);
  }
  __syncthreads();
  const int half = (n_x - 1) / 2;  // pairs (x, n_x - x), x = 1 .. half
  for (int i = threadIdx.x; i < half * DFT_PCOLS; i += blockDim.x) {
    const int x = 1 + i / DFT_PCOLS, c = i % DFT_PCOLS;
    const double a = sig[x * DFT_PCOLS + c], b = sig[(n_x - x) * DFT_PCOLS + c];
    sig[x * DFT_PCOLS + c] = a + b;
    sig[(n_x - x) * DFT_PCOLS + c] = a - b;
  }
  __syncthreads();
  constexpr int GROUPS = DFT_PCOLS / DFT_TC;
  const int c0 = (threadIdx.x % GROUPS) * DFT_TC;
  const double inv = 1.0 / (double)n_x;
  for (int k = threadIdx.x / GROUPS; k <= n_x / 2; k += blockDim.x / GROUPS) {
    double re[DFT_TC], im[DFT_TC];
#pragma unroll
    for (int j = 0; j < DFT_TC; ++j) {
      re[j] = sig[c0 + j];
      im[j] = 0.0;
      if ((n_x & 1) == 0) {
        const double ny = sig[(n_x / 2) * DFT_PCOLS + c0 + j];
        re[j] += (k & 1) ? -ny : ny;
      }
    }
    int idx = 0;  // (k * x) mod n_x
    for (int x = 1; x <= half; ++x) {
      idx += k;
      if (idx >= n_x) idx -= n_x;
      const double2 w = tw[idx];
      const double2* e = reinterpret_cast<const double2*>(sig + x * DFT_PCOLS + c0);
      const double2* o = reinterpret_cast<const double2*>(sig + (n_x - x) * DFT_PCOLS + c0);
      const double2 e0 = e[0], e1 = e[1], o0 = o[0], o1 = o[1];
      re[0] = fma(e0.x, w.x, re[0]);
      re[1] = fma(e0.y, w.x, re[1]);
      re[2] = fma(e1.x, w.x, re[2]);
      re[3] = fma(e1.y, w.x, re[3]);
      im[0] = fma(o0.x, w.y, im[0]);
      im[1] = fma(o0.y, w.y, im[1]);
      im[2] = fma(o1.x, w.y, im[2]);
      im[3] = fma(o1.y, w.y, im[3]);
    }
    const int r0 = row_of ? row_of[k] : k;
    const int km = (n_x - k) % n_x;
    const int r1 = km != k ? (row_of ? row_of[km] : km) : -1;
#pragma unroll
    for (int j = 0; j < DFT_TC; ++j) {
      if (p0 + c0 + j >= n_p) continue;
      if (r0 >= 0)
        out[(size_t)r0 * n_p + p0 + c0 + j] = mk<T>((T)(re[j] * inv), (T)(im[j] * inv));
      if (r1 >= 0)
        out[(size_t)r1 * n_p + p0 + c0 + j] = mk<T>((T)(re[j] * inv), (T)(-im[j] * inv));
    }
  }
}

// Moments over the sample axis of selected coefficient rows (coef: (n_k_total, n_p) complex):
//   out[i]             = sum_p c_i(p)                 i < K
//   out[K + i]         = sum_p |c_i(p)|^2 (real)
//   out[2K + i*K + j]  = sum_p conj(c_i(p)) * c_j(p)
// One CTA of 128 threads per output: threads stride over the samples (coalesced), complex128
// accumulators, fixed xor-shuffle tree, warp partials summed in warp order - bitwise
// reproducible.  (One warp per output left 13 CTAs on the GPU: 14 us of latency.)
template <typename T>
__global__ void __launch_bounds__(128) k_coef_moments(const cx<T>* __restrict__ coef,
                                                      const int32_t* __restrict__ rows, int K,
                                                      int64_t n_p, double2* __restrict__ out) {
  __shared__ double2 part[4];
  const int64_t t = blockIdx.x;
  const int lane = threadIdx.x & 31, tid = threadIdx.x;
  if (t >= (int64_t)K * K + 2 * K) return;
  double re = 0.0, im = 0.0;
  if (t < 2 * K) {
    const cx<T>* ci = coef + (size_t)rows[t % K] * n_p;
    if (t < K) {
      for (int64_t p = tid; p < n_p; p += 128) {
        re += (double)ci[p].x;
        im += (double)ci[p].y;
      }
    } else {
      for (int64_t p = tid; p < n_p; p += 128)
        re += (double)ci[p].x * ci[p].x + (double)ci[p].y * ci[p].y;
    }
  } else {
    const int i = (int)((t - 2 * K) / K), j = (int)((t - 2 * K) % K);
    const cx<T>* ci = coef + (size_t)rows[i] * n_p;
    const cx<T>* cj = coef + (size_t)rows[j] * n_p;
    for (int64_t p = tid; p < n_p; p += 128) {
      const double ar = ci[p].x, ai = ci[p].y, br = cj[p].x, bi = cj[p].y;
      re += ar * br + ai * bi;  // conj(a) * b
      im += ar * bi - ai * br;
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    re += __shfl_xor_sync(0xffffffffu, re, off);
    im += __shfl_xor_sync(0xffffffffu, im, off);
  }
  if (lane == 0) part[tid >> 5] = make_double2(re, im);
  __syncthreads();
  if (tid == 0) {
    double2 s = part[0];
    for (int w = 1; w < 4; ++w) {
      s.x += part[w].x;
      s.y += part[w].y;
    }
    out[t] = s;
  }
}

// Sub-register outputs (jaqsi.py:79-146): partial trace of density matrices and marginal
// probabilities, reduced on the device so that only the 2^k x 2^k / 2^k result leaves it.
// keep[j]: bit position (in the n-bit basis index) of output bit j, MSB first; gone[t]: the
// traced positions.  One thread per output entry, traced configurations in index order.
struct BitSel {
  int8_t keep[16], gone[16];
  int32_t k, g;
};

__device__ __forceinline__ uint32_t bits_deposit(uint32_t v, const int8_t* pos, int count,
                                                 bool msb_first) {
  uint32_t o = 0;
  for (int j = 0; j < count; ++j)
    o |= ((v >> (msb_first ? count - 1 - j : j)) & 1u) << pos[j];
  return o;
}

template <typename T>
__global__ void k_partial_trace(const cx<T>* __restrict__ rho, int64_t batch, int n,
                                const BitSel sel, cx<T>* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = (int64_t)1 << (2 * sel.k);
  if (t >= batch * per) return;
  const int64_t b = t / per;
  const uint32_t io = (uint32_t)((t % per) >> sel.k), jo = (uint32_t)((t % per) & ((1u << sel.k) - 1u));
  const uint32_t ik = bits_deposit(io, sel.keep, sel.k, true);
  const uint32_t jk = bits_deposit(jo, sel.keep, sel.k, true);
  const cx<T>* r = rho + ((size_t)b << (2 * n));
  T re = (T)0, im = (T)0;
  for (uint32_t c = 0; c < (1u << sel.g); ++c) {
    const uint32_t tr = bits_deposit(c, sel.gone, sel.g, false);
    const cx<T> v = r[((size_t)(ik | tr) << n) | (jk | tr)];
    re += v.x;
    im += v.y;
  }
  out[t] = mk<T>(re, im);
}

template <typename T>
__global__ void k_marginal_probs(const T* __restrict__ probs, int64_t batch, int n,
                                 const BitSel sel, T* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = (int64_t)1 << sel.k;
  if (t >= batch * per) return;
  const int64_t b = t / per;
  const uint32_t vk = bits_deposit((uint32_t)(t % per), sel.keep, sel.k, true);
  const T* p = probs + ((size_t)b << n);
  T acc = (T)0;
  for (uint32_t c = 0; c < (1u << sel.g); ++c) acc += p[vk | bits_deposit(c, sel.gone, sel.g, false)];
  out[t] = acc;
}

// One-shot all-reduce (sum) of n doubles over peer-mapped buffers, ONE CTA per rank, PUSH
// form: remote stores are fire-and-forget, remote loads are round trips - so a rank writes
// its contribution into EVERY peer's buffer and each rank then sums from its own memory.
// Layout of every rank's symmetric buffer: [16 x uint64 arrival flags | 1 x uint64 epoch |
// pad to 256 B | 3 slots (epoch % 3) x 8 source ranks x n doubles].  A call stores its input
// into slot[epoch % 3][rank] of every peer, fences system-wide and publishes the epoch to
// every peer's flag.  mode 0 (synchronous): wait until every peer has published the SAME
// epoch and sum those contributions in rank order (identical bits on every rank).  mode 1
// (pipelined): wait for / sum the PREVIOUS epoch instead - the result of call k is the
// reduction of call k - 1, so a rank never stalls on a peer that is less than one call
// behind (launch skew between GPUs is not serialised into every step).  mode 2 (drain):
// publish nothing, wait for the current epoch and sum it - closes a pipelined sequence.  A
// writer can be at most two epochs ahead of the slowest reader: three slots.
struct PeerReduce {
  void* buf[8];
  int32_t n_peers, rank, mode, pad;
  int64_t n;
};

__global__ void __launch_bounds__(256) k_allreduce_oneshot(PeerReduce R,
                                                           const double* __restrict__ in,
                                                           double* __restrict__ out) {
  __shared__ unsigned long long s_epoch;
  __shared__ int s_ok;
  unsigned long long* mine = static_cast<unsigned long long*>(R.buf[R.rank]);
  if (threadIdx.x == 0) {
    s_epoch = mine[16] + (R.mode == 2 ? 0ull : 1ull);
    mine[16] = s_epoch;
    s_ok = 1;
  }
  __syncthreads();
  const unsigned long long epoch = s_epoch;
  const unsigned long long want = R.mode == 1 ? epoch - 1ull : epoch;
  const int64_t n = R.n;
  auto slot = [&](int owner, unsigned long long e, int src) {
    return reinterpret_cast<double*>(static_cast<unsigned char*>(R.buf[owner]) + 256) +
           ((e % 3ull) * 8 + src) * n;
  };
  if (R.mode != 2) {
    for (int r = 0; r < R.n_peers; ++r) {
      double* dst = slot(r, epoch, R.rank);
      for (int64_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = in[i];
    }
    if (R.pad == 0) __threadfence_system();  // every thread fences its own stores
  }
  __syncthreads();
  if (threadIdx.x < R.n_peers) {
    if (R.mode != 2) {
      // pad = 1: the CTA barrier orders the data stores of all threads before this thread,
      // whose system-scope fence + release store is cumulative over them (the grid-sync
      // pattern) - n_peers fences instead of 256
      if (R.pad != 0) __threadfence_system();
      unsigned long long* flag = static_cast<unsigned long long*>(R.buf[threadIdx.x]) + R.rank;
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(epoch) : "memory");
    }
    const unsigned long long* wait = mine + threadIdx.x;
    unsigned long long seen = 0;
    long long spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(wait) : "memory");
    } while (seen < want && ++spins < (1ll << 28));
    if (seen < want) s_ok = 0;  // a peer never arrived: poison the result instead of hanging
  }
  __syncthreads();
  const bool ok = s_ok != 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    if (want > 0)
      for (int r = 0; r < R.n_peers; ++r) s += __ldcv(slot(R.rank, want, r) + i);
    out[i] = ok ? s : __longlong_as_double(0x7ff8000000000000ll);
  }
}

}  // namespace qmlb
