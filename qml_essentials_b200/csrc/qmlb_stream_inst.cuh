// Instantiates the streaming kernel for one precision.
#include "qmlb_internal.h"
#include "qmlb_stream.cuh"

namespace qmlb {

template <bool HEAVY, typename IDX>
static void launch_v(const qmlb_program* p, const RunArgs& R, const StreamPass& pass, dim3 grid,
                     cx<QMLB_T>* s, size_t smem, cudaStream_t st) {
  static bool attr_set = false;  // > 48 KB of dynamic shared memory needs the opt-in (once)
  if (!attr_set) {
    cudaFuncSetAttribute(k_stream<QMLB_T, QMLB_STREAM_R, HEAVY, IDX>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr_set = true;
  }
  k_stream<QMLB_T, QMLB_STREAM_R, HEAVY, IDX><<<grid, STREAM_THREADS, smem, st>>>(p->dev, R, pass, s);
}

cudaError_t QMLB_LAUNCH_STREAM(const qmlb_program* p, const RunArgs& R, const StreamPass& pass,
                               dim3 grid, void* state, cudaStream_t st) {
  cx<QMLB_T>* s = static_cast<cx<QMLB_T>*>(state);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const size_t smem = ((size_t)pass.matw + 2 * (size_t(1) << QMLB_STREAM_R) * STREAM_THREADS) *
                      sizeof(cx<QMLB_T>);
  const bool heavy = pass.flags & QMLB_PASS_HEAVY;
  const bool narrow = pass.n_bits <= 32;  // element-relative indices
  if (heavy && narrow) launch_v<true, uint32_t>(p, R, pass, grid, s, smem, st);
  else if (heavy) launch_v<true, uint64_t>(p, R, pass, grid, s, smem, st);
  else if (narrow) launch_v<false, uint32_t>(p, R, pass, grid, s, smem, st);
  else launch_v<false, uint64_t>(p, R, pass, grid, s, smem, st);
  return cudaGetLastError();
}

}  // namespace qmlb
