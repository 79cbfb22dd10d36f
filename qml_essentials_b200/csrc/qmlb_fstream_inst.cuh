// Instantiates the streaming frame engine for one precision.
#include <cstdlib>

#include "qmlb_internal.h"
#include "qmlb_fstream.cuh"

namespace qmlb {

template <int THREADS>
static cudaError_t launch_fstream_t(const qmlb_program* p, const RunArgs& R, void* state,
                                    const void* premats, int init_mode, cudaStream_t st) {
  auto kern = k_fstream<QMLB_T, THREADS>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, p->frame_smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  const FrameProg& F = p->frame;
  const int64_t total = R.batch << F.outer_bits;
  const unsigned grid = (unsigned)std::max<int64_t>(
      1, std::min<int64_t>(total, (int64_t)per_sm * p->sm_count));
  bool first = true;
  for (const FramePassHost& ps : p->fstream_passes) {
    FStreamPass P{};
    P.steps = p->frame_steps_dev + ps.first_step;
    P.n_steps = ps.n_steps;
    P.n_bits = F.n_bits;
    P.tile_bits = F.tile_bits;
    P.low_bits = p->fstream_low_bits;
    P.outer_bits = F.outer_bits;
    P.init = (first && ps.init) ? init_mode : 0;
    P.premat_row = F.premat_row;
    P.mat_cap = F.mat_cap;
    for (int i = 0; i < F.tile_bits; ++i) P.tp[i] = (uint8_t)ps.tp[i];
    for (int g = 0; g < F.outer_bits; ++g) P.opos[g] = (uint8_t)ps.opos[g];
    g_launches.fetch_add(1, std::memory_order_relaxed);
    kern<<<grid, THREADS, p->frame_smem, st>>>(R, P, static_cast<cx<QMLB_T>*>(state),
                                          static_cast<const cx<QMLB_T>*>(premats));
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    first = false;
  }
  return cudaSuccess;
}

cudaError_t QMLB_LAUNCH_FSTREAM(const qmlb_program* p, const RunArgs& R, void* state,
                                const void* premats, int init_mode, cudaStream_t st) {
  // one item per thread (512 threads, 64 registers, twice the resident warps) measured
  // against two items per thread (256 threads): see DESIGN.md 4.1
  static const int threads = [] {
    const char* v = std::getenv("QMLB_FSTREAM_THREADS");
    return v ? std::atoi(v) : 256;
  }();
  if (threads == 512 && p->frame.tile_bits >= 13)
    return launch_fstream_t<512>(p, R, state, premats, init_mode, st);
  return launch_fstream_t<256>(p, R, state, premats, init_mode, st);
}

}  // namespace qmlb
