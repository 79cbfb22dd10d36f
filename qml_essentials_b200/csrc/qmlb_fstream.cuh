// Streaming form of the frame engine (strategy 4): the state stays in HBM; ONE launch = ONE
// pass = every tile of 2^T amplitudes is read once, taken through ALL steps of the pass in
// shared memory (register-group sub-passes, CX folded into the frame - qmlb_frame.cuh) and
// written once.  A pass of the previous streaming kernel (qmlb_stream.cuh) could only apply
// the ops inside ONE 4-bit register group between its read and its write; a tile pass
// applies everything that fits T bits, so the HBM passes of a circuit drop by 2-3x.
//
// Tile movement is done by the TMA engine: the tile consists of 2^(T-L) runs of 2^L
// contiguous amplitudes (the L lowest HBM bit positions are part of every tile); each run
// is one cp.async.bulk global -> shared copy completing on an mbarrier, and one
// cp.async.bulk shared -> global copy on the way back (SASS: UBLKCP).  Several CTAs per SM
// (one tile buffer each) overlap one CTA's copies with another CTA's arithmetic.
#pragma once

#include "qmlb_frame.cuh"

namespace qmlb {

struct FStreamPass {
  const FrameStep* steps;   // device, the steps of this pass
  int32_t n_steps;
  int32_t n_bits, tile_bits, low_bits, outer_bits;
  int32_t init;             // 1: |0..0> (no read), 2: zero vector, 0: read the state
  int32_t premat_row;
  int32_t mat_cap;          // matrix entries of the largest pass (shared-memory area)
  uint8_t tp[16];           // tile index bit -> HBM bit position (tp[i] = i for i < low_bits)
  uint8_t opos[32];         // tile number bit -> HBM bit position
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA bulk copies (non-tensor form): global -> shared completes on the mbarrier,
// shared -> global is tracked by the bulk async-group
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem)),
      "l"(gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem, const void* smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem),
               "r"(smem_u32(smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 512 ? 2 : (sizeof(T) == 4 ? 3 : 2))
    k_fstream(RunArgs A, const FStreamPass P, cx<T>* __restrict__ gstate,
              const cx<T>* __restrict__ premats) {
  extern __shared__ __align__(128) unsigned char fsm[];
  const int Tb = P.tile_bits, Lb = P.low_bits;
  const uint32_t tile_n = 1u << Tb;
  // [tile | matrices of this pass | two step records | relayout tables | mbarrier]
  cx<T>* tile = reinterpret_cast<cx<T>*>(fsm);
  cx<T>* mats = tile + tile_n;
  FrameStep* sstep = reinterpret_cast<FrameStep*>(mats + P.mat_cap);
  uint32_t* tab_lo = reinterpret_cast<uint32_t*>(sstep + 2);
  uint32_t* tab_hi = tab_lo + 256;
  uint64_t* bar = reinterpret_cast<uint64_t*>(tab_hi + 64);

  const uint32_t n_items = 1u << (Tb - FRAME_R);
  const uint32_t n_runs = 1u << (Tb - Lb);
  const unsigned run_bytes = (unsigned)(sizeof(cx<T>) << Lb);
  const int64_t tiles_per_elem = (int64_t)1 << P.outer_bits;
  const int64_t total = A.batch * tiles_per_elem;

  // step records are double-buffered: record si + 1 is fetched while step si runs (the
  // barrier that ends a step publishes it), so the shared memory holds two of them
  auto fetch_step = [&](int si) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
      reinterpret_cast<uint32_t*>(&sstep[si & 1])[i] =
          reinterpret_cast<const uint32_t*>(&P.steps[si])[i];
  };
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  unsigned phase = 0;
  int64_t cur_elem = -1;
  for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
    const int64_t bl = w / tiles_per_elem;
    const uint32_t tnum = (uint32_t)(w % tiles_per_elem);
    // HBM offset of the tile: the bits of the tile number go to the outer positions
    uint64_t obase = 0;
    for (int g = 0; g < P.outer_bits; ++g) obase |= (uint64_t)((tnum >> g) & 1u) << P.opos[g];
    cx<T>* gs = gstate + ((size_t)bl << P.n_bits) + obase;
    auto run_offset = [&](uint32_t r) -> uint64_t {
      uint64_t o = 0;
      for (int k = 0; k < Tb - Lb; ++k) o |= (uint64_t)((r >> k) & 1u) << P.tp[Lb + k];
      return o;
    };

    if (bl != cur_elem) {  // matrices this pass uses (a CTA rarely changes element)
      const cx<T>* prow = premats + (size_t)bl * P.premat_row;
      for (int si = 0; si < P.n_steps; ++si) {
        const FrameStep& gst = P.steps[si];
        if (gst.kind != QMLB_FSTEP_SUBPASS) continue;
        for (int o = 0; o < gst.n_ops; ++o) {
          const FrameOp fo = gst.ops[o];
          if (fo.code == QMLB_FOP_SIGN) {
            o += 3;
            continue;
          }
          // entries of the op's source: a controlled 2x2 (k = 2) stores only the 2x2
          const int n = fo.code == QMLB_FOP_DIAG    ? (1 << fo.k)
                        : fo.code == QMLB_FOP_CTRL1 ? 4
                                                    : (1 << (2 * fo.k));
          for (int e = threadIdx.x; e < n; e += blockDim.x)
            mats[fo.smem_off + e] = prow[fo.premat_off + e];
          if (fo.code == QMLB_FOP_DIAG) ++o;
        }
      }
      cur_elem = bl;
    }
    if (P.n_steps > 0) fetch_step(0);
    if (P.init) {
      for (uint32_t i = threadIdx.x; i < tile_n; i += blockDim.x)
        tile[i] = mk<T>((i == 0 && tnum == 0 && P.init == 1) ? (T)1 : (T)0, (T)0);
      __syncthreads();
    } else {
      if (threadIdx.x == 0) mbar_expect_tx(bar, (unsigned)(tile_n * sizeof(cx<T>)));
      __syncthreads();  // expect_tx is posted (and the previous tile's stores have been read)
      for (uint32_t r = threadIdx.x; r < n_runs; r += blockDim.x)
        bulk_g2s(tile + ((size_t)r << Lb), gs + run_offset(r), run_bytes, bar);
      mbar_wait(bar, phase);
      phase ^= 1u;
    }

    for (int si = 0; si < P.n_steps; ++si) {
      const FrameStep& st = sstep[si & 1];
      if (si + 1 < P.n_steps) fetch_step(si + 1);
      if (st.kind == QMLB_FSTEP_RELAYOUT) {  // tile-local shuffle (tables depend on the tile)
        for (int i = threadIdx.x; i < 256 + 64; i += blockDim.x) {
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
        for (int g = 0; g < P.outer_bits; ++g)
          if (tnum >> g & 1) cmine ^= (uint32_t)st.qcol[Tb + g];
        cmine &= tile_n - 1u;
        __syncthreads();
        cg::cluster_group cluster = cg::this_cluster();
        constexpr int TB = THREADS == 512 ? 9 : 8;
        const uint32_t* htab = reinterpret_cast<const uint32_t*>(st.ops);
        const int per = (int)(tile_n >> TB);
        if (per == 16)
          frame_relayout<T, 16>(tile, cluster, false, 0u, Tb, TB, (int)threadIdx.x, cmine, tab_lo,
                                tab_hi, htab);
        else if (per == 32 && THREADS == 256)
          frame_relayout<T, 32>(tile, cluster, false, 0u, Tb, TB, (int)threadIdx.x, cmine, tab_lo,
                                tab_hi, htab);
        else if (per == 8)
          frame_relayout<T, 8>(tile, cluster, false, 0u, Tb, TB, (int)threadIdx.x, cmine, tab_lo,
                               tab_hi, htab);
        else if (per == 4)
          frame_relayout<T, 4>(tile, cluster, false, 0u, Tb, TB, (int)threadIdx.x, cmine, tab_lo,
                               tab_hi, htab);
        else if (per == 2)
          frame_relayout<T, 2>(tile, cluster, false, 0u, Tb, TB, (int)threadIdx.x, cmine, tab_lo,
                               tab_hi, htab);
        __syncthreads();
        continue;
      }
      frame_subpass<T, false>(tile, mats, st, tnum, threadIdx.x, blockDim.x, n_items);
      __syncthreads();
    }

    // generic-proxy writes -> visible to the async proxy, then one bulk store per run
    fence_async_smem();
    __syncthreads();
    for (uint32_t r = threadIdx.x; r < n_runs; r += blockDim.x)
      bulk_s2g(gs + run_offset(r), tile + ((size_t)r << Lb), run_bytes);
    bulk_commit();
    bulk_wait_read();  // the tile buffer may be overwritten once the stores have read it
    __syncthreads();
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores landed before exit
}

}  // namespace qmlb
