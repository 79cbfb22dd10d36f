// Instantiates the Pauli-basis frame engine for one precision.
#include "qmlb_internal.h"
#include "qmlb_frame_ptm.cuh"

namespace qmlb {

template <int THREADS, int MINB>
static cudaError_t launch_ptm_t(const qmlb_program* p, const RunArgs& R, const FrameProg& F,
                                const cx<QMLB_T>* premats, void* out, cudaStream_t st) {
  auto kern = k_frame_ptm<QMLB_T, THREADS, MINB>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int per = (1 << F.tile_bits) >> F.team_bits;  // relayout registers per thread
  if (!(per == 16 || per == 32 || (per == 64 && sizeof(QMLB_T) == 4)))
    return cudaErrorInvalidConfiguration;
  const int csize = 1 << F.outer_bits;
  const int64_t units = (R.batch + F.teams - 1) / F.teams;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = p->frame_smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int64_t resident;
  if (csize > 1) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.gridDim = dim3((unsigned)csize * (unsigned)p->sm_count);
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess) return e;
    if (n < 1) return cudaErrorLaunchOutOfResources;
    resident = n;
  } else {
    int per_sm = 0;
    cudaError_t e =
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, p->frame_smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    resident = (int64_t)per_sm * p->sm_count;
  }
  const int64_t clusters = std::max<int64_t>(1, std::min<int64_t>(units, resident));
  cfg.gridDim = dim3((unsigned)(clusters * csize));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaLaunchKernelEx(&cfg, kern, p->dev, R, F, (uint64_t)p->frame_ptm_xmask, premats, out);
}

cudaError_t QMLB_LAUNCH_FRAME_PTM(const qmlb_program* p, const RunArgs& R, const void* premats,
                                  void* out, int out_mode, cudaStream_t st) {
  FrameProg F = p->frame;
  F.steps = p->frame_steps_dev;
  F.out_mode = out_mode;
  const cx<QMLB_T>* pm = static_cast<const cx<QMLB_T>*>(premats);
  if (p->frame_threads == 1024) return launch_ptm_t<1024, 1>(p, R, F, pm, out, st);
  if (p->frame_threads == 512) return launch_ptm_t<512, 1>(p, R, F, pm, out, st);
  if (p->frame_threads == 256 && F.teams == 1) return launch_ptm_t<256, 1>(p, R, F, pm, out, st);
  if (p->frame_threads == 256) return launch_ptm_t<256, 2>(p, R, F, pm, out, st);
  if (p->frame_threads == 128) return launch_ptm_t<128, 2>(p, R, F, pm, out, st);
  if (p->frame_threads == 64) return launch_ptm_t<64, 2>(p, R, F, pm, out, st);
  if (p->frame_threads == 32) return launch_ptm_t<32, 2>(p, R, F, pm, out, st);
  return cudaErrorInvalidConfiguration;
}

}  // namespace qmlb
