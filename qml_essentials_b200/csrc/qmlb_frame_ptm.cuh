// Pauli-basis form of the on-chip frame engine (strategy 5): a noisy density-matrix program
// evolved as the REAL vector r_P = Tr(P rho) of 4^n Pauli coefficients (see plan_frame_ptm in
// qmlb_frame_plan.cu).  Same step program, frame algebra and relayouts as qmlb_frame.cuh;
// what changes is the element (one real number instead of a complex one: config 4 fits a
// cluster of 4 CTAs instead of 8), the gate (a real 4x4 transfer matrix on four reals: 16
// FMA instead of 64; diagonal for depolarizing / flip / phase-damping channels) and the
// Clifford sign op.  <Z_S> is ONE coefficient (r at index S on the z bits), probabilities
// are a Walsh-Hadamard transform of the 2^n coefficients with x = 0.
#pragma once

#include "qmlb_frame.cuh"

namespace qmlb {

// real 4x4 on register bits JA > JB of a group of 16 reals; m = 16 reals, row-major
template <typename T, int JA, int JB, bool DIAG>
__device__ __forceinline__ void ptm_mat2(T (&S)[FRAME_D], const T* __restrict__ m) {
  static_assert(JA > JB, "canonical order");
  if constexpr (DIAG) {
    const T d0 = m[0], d1 = m[5], d2 = m[10], d3 = m[15];
#pragma unroll
    for (int g = 0; g < (1 << (FRAME_R - 2)); ++g) {
      const int t = ((g >> JB) << (JB + 1)) | (g & ((1 << JB) - 1));
      const int i00 = ((t >> JA) << (JA + 1)) | (t & ((1 << JA) - 1));
      S[i00] *= d0;
      S[i00 | (1 << JB)] *= d1;
      S[i00 | (1 << JA)] *= d2;
      S[i00 | (1 << JA) | (1 << JB)] *= d3;
    }
  } else {
    T mm[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) mm[i] = m[i];
#pragma unroll
    for (int g = 0; g < (1 << (FRAME_R - 2)); ++g) {
      const int t = ((g >> JB) << (JB + 1)) | (g & ((1 << JB) - 1));
      const int i00 = ((t >> JA) << (JA + 1)) | (t & ((1 << JA) - 1));
      const int idx[4] = {i00, i00 | (1 << JB), i00 | (1 << JA), i00 | (1 << JA) | (1 << JB)};
      const T a0 = S[idx[0]], a1 = S[idx[1]], a2 = S[idx[2]], a3 = S[idx[3]];
#pragma unroll
      for (int v = 0; v < 4; ++v)
        S[idx[v]] = fma(mm[v * 4 + 3], a3, fma(mm[v * 4 + 2], a2, fma(mm[v * 4 + 1], a1, mm[v * 4] * a0)));
    }
  }
}

// Swizzled tile: element i of a tile lives at i ^ (XOR of the higher NB-bit digits of i, taken
// on its lowest digit) - a GF(2)-linear involution, so swz(a ^ b) = swz(a) ^ swz(b) and the
// bank of an address is the XOR of ALL its digits.  NB = log2(elements per 128-byte
// wavefront): 4 for doubles, 5 for floats.  Aligned runs of 2^NB elements stay runs.
template <int NB>
__device__ __forceinline__ uint32_t ptm_swz(uint32_t i) {
  return i ^ (((i >> NB) ^ (i >> (2 * NB)) ^ (i >> (3 * NB))) & ((1u << NB) - 1u));
}

// negate the slots whose bit is set in w (a flip of the IEEE sign bit, no FP pipe)
template <typename T>
__device__ __forceinline__ void ptm_sign(T (&S)[FRAME_D], unsigned w) {
#pragma unroll
  for (int v = 0; v < FRAME_D; ++v) {
    const unsigned sb = (w << (31 - v)) & 0x80000000u;
    if constexpr (sizeof(T) == 8)
      S[v] = __hiloint2double(__double2hiint(S[v]) ^ (int)sb, __double2loint(S[v]));
    else
      S[v] = __int_as_float(__float_as_int(S[v]) ^ (int)sb);
  }
}

// sign word of the SIGN op in slot o for the item that starts at tile index `base`: the
// local value of the op's four parity rows at slot 0 selects one of 16 planner-made words
__device__ __forceinline__ unsigned ptm_sign_word(const FrameStep& st, int o, uint32_t base,
                                                  unsigned rank) {
  const uint4 rl4 = *reinterpret_cast<const uint4*>(&st.ops[o + 1]);
  const unsigned ro = (unsigned)st.ops[o].smem_off;
  const int lb = ((__popc(base & rl4.x) ^ __popc(rank & (ro & 255u))) & 1) << 3 |
                 ((__popc(base & rl4.y) ^ __popc(rank & ((ro >> 8) & 255u))) & 1) << 2 |
                 ((__popc(base & rl4.z) ^ __popc(rank & ((ro >> 16) & 255u))) & 1) << 1 |
                 ((__popc(base & rl4.w) ^ __popc(rank & (ro >> 24))) & 1);
  return reinterpret_cast<const uint16_t*>(&st.ops[o + 2])[lb];
}

template <typename T, int NB>
__device__ __forceinline__ void ptm_item_addresses(uint32_t base, const uint32_t (&seb)[FRAME_R],
                                                   uint32_t (&ad)[FRAME_D]) {
  ad[0] = ptm_swz<NB>(base);
#pragma unroll
  for (int v = 1; v < FRAME_D; ++v) {
    const int j = 31 - __builtin_clz(v);
    ad[v] = ad[v ^ (1 << j)] ^ seb[j];
  }
}

// Straight-line item body of the common step shape (FrameStep::fast >= 128):
// [signs] A on register pair (1,0) [signs] B on pair (3,2); SA, SB = 0 absent, 1 full, 2 diagonal
template <typename T, int NB, int SA, int SB>
__device__ __forceinline__ void ptm_items_fast(T* tile, const T* mats, const FrameStep& st,
                                               const FrameSubX* sx, uint32_t base0,
                                               const uint32_t (&seb)[FRAME_R], int per_thread,
                                               unsigned rank) {
  const T* ma = mats + 2 * st.foff[0];
  const T* mb = mats + 2 * st.foff[1];
  const int s0 = sx->sg[0], s1 = sx->sg[1], s2 = sx->sg[2], s3 = sx->sg[3];
#pragma unroll 1
  for (int k = 0; k < per_thread; ++k) {
    uint32_t base = base0;
    for (int b = 0; (k >> b) != 0; ++b)
      if (k >> b & 1) base ^= sx->kd[b];
    uint32_t ad[FRAME_D];
    ptm_item_addresses<T, NB>(base, seb, ad);
    T S[FRAME_D];
#pragma unroll
    for (int v = 0; v < FRAME_D; ++v) S[v] = tile[ad[v]];
    if (s0 != 0xff) {
      unsigned w = ptm_sign_word(st, s0, base, rank);
      if (s1 != 0xff) w ^= ptm_sign_word(st, s1, base, rank);
      ptm_sign<T>(S, w);
    }
    if constexpr (SA != 0) ptm_mat2<T, 1, 0, SA == 2>(S, ma);
    if (s2 != 0xff) {
      unsigned w = ptm_sign_word(st, s2, base, rank);
      if (s3 != 0xff) w ^= ptm_sign_word(st, s3, base, rank);
      ptm_sign<T>(S, w);
    }
    if constexpr (SB != 0) ptm_mat2<T, 3, 2, SB == 2>(S, mb);
#pragma unroll
    for (int v = 0; v < FRAME_D; ++v) tile[ad[v]] = S[v];
  }
}

// Density matrix from the Pauli coefficients, one warp per row x (= ket ^ bra):
//   rho[k' ^ x][k'] = 2^-n sum_z i^|x & z| (-1)^(z . k') r[x << n | z]
// (<k|P|k'> of a Pauli string: X / Y move k' to k' ^ x, Z / Y give (-1)^k', Y an extra i).
// The phase goes onto the loaded coefficient, then a Walsh-Hadamard transform over the n
// bits of z: bits >= 5 inside a lane (2^NVL values each), bits < 5 with shuffles.
template <typename T, int NB, int NVL>
__device__ __forceinline__ void ptm_rho_rows(const T* tile, cx<T>* rho, int nq, int Tb,
                                             unsigned rank, int tlane, int tsize, bool valid) {
  constexpr int NV = 1 << NVL;
  const int lane = tlane & 31, warp = tlane >> 5, nwarps = tsize >> 5;
  const int xl_bits = Tb - nq, lb = nq - NVL;  // local x bits; z bits spread over the lanes
  const uint32_t dim = 1u << nq;
  const T inv = (T)1 / (T)dim;
  for (uint32_t xl = warp; xl < (1u << xl_bits); xl += nwarps) {
    const uint32_t x = (rank << xl_bits) | xl;
    T re[NV], im[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const uint32_t z = ((uint32_t)j << 5) | (uint32_t)lane;
      const T v = z < dim ? tile[ptm_swz<NB>((xl << nq) | z)] : (T)0;
      const int m = __popc(x & z) & 3;
      re[j] = m == 0 ? v : (m == 2 ? -v : (T)0);
      im[j] = m == 1 ? v : (m == 3 ? -v : (T)0);
    }
#pragma unroll
    for (int b = 0; b < NVL; ++b)
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (!(j >> b & 1)) {
          const T ar = re[j], ai = im[j], br = re[j | (1 << b)], bi = im[j | (1 << b)];
          re[j] = ar + br, im[j] = ai + bi;
          re[j | (1 << b)] = ar - br, im[j | (1 << b)] = ai - bi;
        }
    for (int b = 0; b < lb; ++b) {
      const bool up = lane >> b & 1;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const T pr = __shfl_xor_sync(0xffffffffu, re[j], 1 << b);
        const T pi = __shfl_xor_sync(0xffffffffu, im[j], 1 << b);
        re[j] = up ? pr - re[j] : re[j] + pr;
        im[j] = up ? pi - im[j] : im[j] + pi;
      }
    }
    if (valid)
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const uint32_t kp = ((uint32_t)j << 5) | (uint32_t)lane;
        if (kp < dim) rho[((size_t)(kp ^ x) << nq) | kp] = mk<T>(re[j] * inv, im[j] * inv);
      }
  }
}

// MINB = 1: one CTA per SM (the tile fills its shared memory anyway), up to 255 registers
template <typename T, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
    k_frame_ptm(DevProg P, RunArgs A, const FrameProg F, const uint64_t xmask,
                const cx<T>* __restrict__ premats, void* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char fsm[];
  const int Tb = F.tile_bits;
  const uint32_t tile_n = 1u << Tb;
  const int teams = F.teams;
  const int tsize = 1 << F.team_bits;
  const int team = threadIdx.x >> F.team_bits;
  const int tlane = threadIdx.x & (tsize - 1);
  // [tiles (real) | matrices | 2 step records | relayout tables | scratch]
  T* tiles = reinterpret_cast<T*>(fsm);
  cx<T>* mats_all = reinterpret_cast<cx<T>*>(tiles + (size_t)teams * tile_n);
  FrameStep* sstep = reinterpret_cast<FrameStep*>(mats_all + (size_t)teams * F.mat_cap);
  uint32_t* tab_lo = reinterpret_cast<uint32_t*>(sstep + 2);
  uint32_t* tab_hi = tab_lo + 256;
  double* scratch_all = reinterpret_cast<double*>(tab_hi + 128);
  T* tile = tiles + (size_t)team * tile_n;
  const T* mats = reinterpret_cast<const T*>(mats_all + (size_t)team * F.mat_cap);  // reals

  const bool clustered = F.outer_bits > 0;
  cg::cluster_group cluster = cg::this_cluster();
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
  const int per_thread = 1 << (Tb - FRAME_R - F.team_bits);  // items of one thread
  const int nq = F.n_qubits;
  constexpr int NB = sizeof(T) == 8 ? 4 : 5;
  const uint32_t tile_mask = tile_n - 1u;

  for (int64_t rd = 0; rd < rounds; ++rd) {
    const int64_t bl = rd * per_round + cluster_id * teams + team;
    const bool valid = bl < A.batch;
    const cx<T>* prow = premats + (size_t)(valid ? bl : 0) * F.premat_row;

    // |0..0><0..0| = 2^-n prod_q (I + Z_q): coefficient 1 wherever every x bit is 0
    for (uint32_t i = tlane; i < tile_n; i += tsize) {
      const uint64_t full = ((uint64_t)rank << Tb) | ptm_swz<NB>(i);
      tile[i] = (full & xmask) ? (T)0 : (T)1;
    }
    {
      cx<T>* mdst = mats_all + (size_t)team * F.mat_cap;
      for (int i = tlane; i < F.premat_row; i += tsize) mdst[i] = prow[i];
    }
    for (int i = threadIdx.x; i < 256; i += THREADS)
      reinterpret_cast<uint32_t*>(&sstep[0])[i] = reinterpret_cast<const uint32_t*>(&F.steps[0])[i];
    sync_all();

    for (int si = 0; si < F.n_steps; ++si) {
      const FrameStep& st = sstep[si & 1];
      if (si + 1 < F.n_steps)
        for (int i = threadIdx.x; i < 256; i += THREADS)
          reinterpret_cast<uint32_t*>(&sstep[(si + 1) & 1])[i] =
              reinterpret_cast<const uint32_t*>(&F.steps[si + 1])[i];

      if (st.kind == QMLB_FSTEP_RELAYOUT) {
        for (int i = threadIdx.x; i < 256 + 128; i += THREADS) {
          uint32_t acc = 0;
          if (i < 256) {
            for (int b = 0; b < 8; ++b)
              if (i >> b & 1) acc ^= (uint32_t)st.qcol[b];
            tab_lo[i] = (acc & ~tile_mask) | ptm_swz<NB>(acc & tile_mask);
          } else {
            const int h = i - 256;
            for (int b = 0; b < 7; ++b)
              if ((h >> b & 1) && 8 + b < Tb) acc ^= (uint32_t)st.qcol[8 + b];
            tab_hi[h] = (acc & ~tile_mask) | ptm_swz<NB>(acc & tile_mask);
          }
        }
        uint32_t cmine = 0;
        for (int g = 0; g < F.outer_bits; ++g)
          if (rank >> g & 1) cmine ^= (uint32_t)st.qcol[Tb + g];
        cmine = (cmine & ~tile_mask) | ptm_swz<NB>(cmine & tile_mask);
        const bool across = clustered && st.mat_entries == 0;
        const uint32_t* htab = nullptr;  // plain relayout form (see frame_relayout_v)
        if (across)
          cluster.sync();
        else
          __syncthreads();
        const int per = (int)(tile_n >> F.team_bits);
        if (per == 16)
          frame_relayout_v<T, 16, NB>(tile, cluster, across, rank, Tb, F.team_bits, tlane,
                                      cmine, tab_lo, tab_hi, htab);
        else if (per == 32)
          frame_relayout_v<T, 32, NB>(tile, cluster, across, rank, Tb, F.team_bits, tlane,
                                      cmine, tab_lo, tab_hi, htab);
        else if (sizeof(T) == 4 && per == 64)  // floats only: 64 doubles do not fit the registers
          frame_relayout_v<T, sizeof(T) == 4 ? 64 : 16, NB>(tile, cluster, across, rank, Tb, F.team_bits, tlane,
                                      cmine, tab_lo, tab_hi, htab);
        __syncthreads();
        continue;
      }

      if (valid) {
        // per-step constants in registers: the parity rows of the register bits, the four
        // basis offsets of the group (swizzled), the thread's first item
        const FrameSubX* sx = reinterpret_cast<const FrameSubX*>(st.qcol);
        uint32_t eb[FRAME_R], seb[FRAME_R];
        uint32_t base0 = 0;
        {
          const uint4 ip4 = *reinterpret_cast<const uint4*>(sx->ipos);
          const uint32_t ipw[4] = {ip4.x, ip4.y, ip4.z, ip4.w};
          constexpr int TBC = THREADS == 1024 ? 10 : THREADS == 512 ? 9 : THREADS == 256 ? 8
                              : THREADS == 128 ? 7 : THREADS == 64 ? 6 : 5;
          if (teams == 1) {  // one tile per CTA: the lane bits are those of the thread index
#pragma unroll
            for (int b = 0; b < TBC; ++b)
              base0 |= (((uint32_t)tlane >> b) & 1u) << ((ipw[b >> 2] >> (8 * (b & 3))) & 31u);
          } else {
#pragma unroll
            for (int b = 0; b < 12; ++b)
              if (b < F.team_bits)
                base0 |= (((uint32_t)tlane >> b) & 1u) << ((ipw[b >> 2] >> (8 * (b & 3))) & 31u);
          }
        }
        {
          int c = 0;
#pragma unroll
          for (int j = 0; j < FRAME_R; ++j) {
            const FramePar pr = st.par[j];
            eb[j] = st.eoff[1 << j];
            seb[j] = ptm_swz<NB>(eb[j]);
            c |= ((__popc(base0 & pr.rloc) ^ __popc(rank & pr.rout)) & 1) << j;
          }
#pragma unroll
          for (int j = 0; j < FRAME_R; ++j)
            if (c >> j & 1) base0 ^= eb[j];  // slot v then holds logical value v
        }
        const int n_ops = st.n_ops;
        bool fast_done = true;
        switch (st.fast) {
#define QMLB_PTM_FAST(a, b)                                                                  \
  case 128 + 3 * a + b:                                                                      \
    ptm_items_fast<T, NB, a, b>(tile, mats, st, sx, base0, seb, per_thread, rank);           \
    break;
          QMLB_PTM_FAST(0, 1) QMLB_PTM_FAST(0, 2) QMLB_PTM_FAST(1, 0) QMLB_PTM_FAST(1, 1)
          QMLB_PTM_FAST(1, 2) QMLB_PTM_FAST(2, 0) QMLB_PTM_FAST(2, 1) QMLB_PTM_FAST(2, 2)
#undef QMLB_PTM_FAST
          default: fast_done = false;
        }
#pragma unroll 1
        for (int k = 0; k < (fast_done ? 0 : per_thread); ++k) {
          uint32_t base = base0;  // item k of this thread: address linear in the item bits
          for (int b = 0; (k >> b) != 0; ++b)
            if (k >> b & 1) base ^= sx->kd[b];
          uint32_t ad[FRAME_D];
          ptm_item_addresses<T, NB>(base, seb, ad);
          T S[FRAME_D];
#pragma unroll
          for (int v = 0; v < FRAME_D; ++v) S[v] = tile[ad[v]];
#pragma unroll 1
          for (int o = 0; o < n_ops; ++o) {
            const FrameOp fo = st.ops[o];
            const T* m = mats + 2 * fo.smem_off;  // smem_off counts complex slots
            int key = 5;
            if (fo.code == QMLB_FOP_SIGN)
              key = 4;
            else if (fo.j0 == 1 && fo.j1 == 0)
              key = fo.shape == QMLB_FSHAPE_PDIAG ? 1 : 0;
            else if (fo.j0 == 3 && fo.j1 == 2)
              key = fo.shape == QMLB_FSHAPE_PDIAG ? 3 : 2;
            switch (key) {
              case 0: ptm_mat2<T, 1, 0, false>(S, m); break;
              case 1: ptm_mat2<T, 1, 0, true>(S, m); break;
              case 2: ptm_mat2<T, 3, 2, false>(S, m); break;
              case 3: ptm_mat2<T, 3, 2, true>(S, m); break;
              case 4: {
                const unsigned w = ptm_sign_word(st, o, base, rank);
                ptm_sign<T>(S, w);
                o += 3;
                break;
              }
              default: {
                const bool dg = fo.shape == QMLB_FSHAPE_PDIAG;
                dispatch2<T, FRAME_R>(fo.j0, fo.j1, [&](auto JA, auto JB) {
                  constexpr int A_ = decltype(JA)::value, B_ = decltype(JB)::value;
                  if constexpr (A_ > B_) {
                    if (dg)
                      ptm_mat2<T, A_, B_, true>(S, m);
                    else
                      ptm_mat2<T, A_, B_, false>(S, m);
                  }
                });
              }
            }
          }
#pragma unroll
          for (int v = 0; v < FRAME_D; ++v) tile[ad[v]] = S[v];
        }
      }
      __syncthreads();
    }

    // ---- result: coefficient of Pauli (x, z) at index x << n | z (wire q = bit n-1-q) -------
    if (F.out_mode == 2) {
      // <Z_S> = r at x = 0, z = S: index zmask < 2^n sits in the tile of cluster rank 0
      if (valid && rank == 0)
        for (int j = tlane; j < F.n_obs; j += tsize)
          reinterpret_cast<T*>(out)[(size_t)bl * F.n_obs + j] =
              tile[ptm_swz<NB>((uint32_t)P.obs[j].zmask)];
    } else if (F.out_mode == 0) {
      cx<T>* rho = reinterpret_cast<cx<T>*>(out) + ((size_t)(valid ? bl : 0) << (2 * nq));
      if (tsize >= 32) {
        switch (nq > 5 ? nq - 5 : 0) {
          case 0: ptm_rho_rows<T, NB, 0>(tile, rho, nq, Tb, rank, tlane, tsize, valid); break;
          case 1: ptm_rho_rows<T, NB, 1>(tile, rho, nq, Tb, rank, tlane, tsize, valid); break;
          case 2: ptm_rho_rows<T, NB, 2>(tile, rho, nq, Tb, rank, tlane, tsize, valid); break;
          case 3: ptm_rho_rows<T, NB, 3>(tile, rho, nq, Tb, rank, tlane, tsize, valid); break;
          default:  // n = 9: complex64 only (a complex128 tile holds at most 2 * 8 bits)
            ptm_rho_rows<T, NB, sizeof(T) == 4 ? 4 : 3>(tile, rho, nq, Tb, rank, tlane, tsize, valid);
        }
      } else if (valid) {
        // a tile smaller than a warp's row scheme (n <= 4): direct sums
        const uint32_t dim = 1u << nq;
        const T inv = (T)1 / (T)dim;
        for (uint32_t e = tlane; e < tile_n; e += tsize) {
          const uint32_t x = e >> nq, kp = e & (dim - 1u);
          T re = 0, im = 0;
          for (uint32_t z = 0; z < dim; ++z) {
            T v = tile[ptm_swz<NB>((x << nq) | z)];
            if (__popc(z & kp) & 1) v = -v;
            const int m = __popc(x & z) & 3;
            re += m == 0 ? v : (m == 2 ? -v : (T)0);
            im += m == 1 ? v : (m == 3 ? -v : (T)0);
          }
          rho[((size_t)(kp ^ x) << nq) | kp] = mk<T>(re * inv, im * inv);
        }
      }
    } else if (F.out_mode == 1) {
      // p(b) = 2^-n sum_S (-1)^(b.S) r[S]: Walsh-Hadamard transform of the x = 0 coefficients
      // (the first 2^n entries of rank 0's tile), done in that CTA's scratch (teams == 1)
      if (rank == 0) {
        const uint32_t dim = 1u << nq;
        double* scratch = scratch_all + (size_t)team * dim;  // teams * 2^n <= 2048
        for (uint32_t i = tlane; i < dim; i += tsize) scratch[i] = (double)tile[ptm_swz<NB>(i)];
        __syncthreads();
        for (int b = 0; b < nq; ++b) {
          for (uint32_t i = tlane; i < dim / 2; i += tsize) {
            const uint32_t lo = ((i >> b) << (b + 1)) | (i & ((1u << b) - 1u)), hi = lo | (1u << b);
            const double u = scratch[lo], w = scratch[hi];
            scratch[lo] = u + w;
            scratch[hi] = u - w;
          }
          __syncthreads();
        }
        const double inv = 1.0 / (double)dim;
        if (valid)
          for (uint32_t i = tlane; i < dim; i += tsize)
            reinterpret_cast<T*>(out)[((size_t)bl << nq) + i] = (T)(scratch[i] * inv);
      }
    }
    sync_all();
  }
}

}  // namespace qmlb
