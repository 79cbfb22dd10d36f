// Instantiates the shared-memory tile kernel for one precision.
#include "qmlb_internal.h"
#include "qmlb_tile.cuh"

namespace qmlb {

cudaError_t QMLB_LAUNCH_TILE(const qmlb_program* p, const RunArgs& R, const PassDev& pass,
                             unsigned grid, void* state, cudaStream_t st) {
  cx<QMLB_T>* s = static_cast<cx<QMLB_T>*>(state);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (p->warp_team)
    k_tile<QMLB_T, true><<<grid, TILE_THREADS, p->smem, st>>>(p->dev, R, pass, s);
  else
    k_tile<QMLB_T, false><<<grid, TILE_THREADS, p->smem, st>>>(p->dev, R, pass, s);
  return cudaGetLastError();
}

cudaError_t QMLB_TILE_SET_SMEM(size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(k_tile<QMLB_T, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_tile<QMLB_T, false>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace qmlb
