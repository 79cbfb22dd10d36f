// Instantiates the streaming kernel for one precision and one weight class (lean passes /
// passes with a 3-4 bit dense or permutation op) - one translation unit each so the build
// parallelises.
#include <cstdlib>

#include "qmlb_internal.h"
#include "qmlb_stream.cuh"

namespace qmlb {

template <typename IDX>
static void launch_v(const qmlb_program* p, const RunArgs& R, const StreamPass& pass, dim3 grid,
                     cx<QMLB_T>* s, const cx<QMLB_T>* premats, const StreamPeers& peers,
                     size_t smem, cudaStream_t st) {
  static bool attr_set = false;  // > 48 KB of dynamic shared memory needs the opt-in (once)
  if (!attr_set) {
    cudaFuncSetAttribute(k_stream<QMLB_T, QMLB_STREAM_R, QMLB_STREAM_HEAVY, IDX>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr_set = true;
  }
  k_stream<QMLB_T, QMLB_STREAM_R, QMLB_STREAM_HEAVY, IDX>
      <<<grid, STREAM_THREADS, smem, st>>>(p->dev, R, pass, s, premats, peers);
}

#if !QMLB_STREAM_HEAVY
cudaError_t QMLB_LAUNCH_STREAM_MATS(const qmlb_program* p, const RunArgs& R, void* out,
                                    cudaStream_t st) {
  const int64_t total = R.batch * (int64_t)p->stream_matlist.size();
  if (total == 0) return cudaSuccess;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  bool plain = true;
  for (const StreamMatOp& mo : p->stream_matlist) plain = plain && mo.swap2 == 0;
  const unsigned grid = (unsigned)((total + 127) / 128);
  if (plain)
    k_stream_mats<QMLB_T, true><<<grid, 128, 0, st>>>(
        p->dev, R, p->stream_matlist_dev, (int)p->stream_matlist.size(), p->stream_mat_row,
        static_cast<cx<QMLB_T>*>(out));
  else
    k_stream_mats<QMLB_T, false><<<grid, 128, 0, st>>>(
        p->dev, R, p->stream_matlist_dev, (int)p->stream_matlist.size(), p->stream_mat_row,
        static_cast<cx<QMLB_T>*>(out));
  return cudaGetLastError();
}
#endif

cudaError_t QMLB_LAUNCH_STREAM(const qmlb_program* p, const RunArgs& R, const StreamPass& pass,
                               dim3 grid, void* state, const void* premats_v,
                               const StreamPeers* peers_in, cudaStream_t st) {
  StreamPeers peers{};
  if (peers_in) peers = *peers_in;
  cx<QMLB_T>* s = static_cast<cx<QMLB_T>*>(state);
  const cx<QMLB_T>* premats = static_cast<const cx<QMLB_T>*>(premats_v);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const size_t smem = ((size_t)pass.matw + 2 * (size_t(1) << QMLB_STREAM_R) * STREAM_THREADS) *
                      sizeof(cx<QMLB_T>);
  // element-relative indices fit 32 bits (QMLB_FORCE_IDX64=1: run the 64-bit index
  // instantiation anyway - test hook for the path states beyond 2^32 amplitudes take)
  static const bool force64 = [] {
    const char* v = std::getenv("QMLB_FORCE_IDX64");
    return v && std::atoi(v) != 0;
  }();
  if (pass.n_bits <= 32 && !force64)
    launch_v<uint32_t>(p, R, pass, grid, s, premats, peers, smem, st);
  else
    launch_v<uint64_t>(p, R, pass, grid, s, premats, peers, smem, st);
  return cudaGetLastError();
}

}  // namespace qmlb
