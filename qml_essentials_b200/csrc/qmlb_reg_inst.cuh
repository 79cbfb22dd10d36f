// Instantiates the register kernel for one precision (QMLB_T / QMLB_SUFFIX).
#include <algorithm>
#include <cstdlib>

#include "qmlb_internal.h"
#include "qmlb_reg.cuh"

namespace qmlb {

template <int N>
static void launch_n(const qmlb_program* p, const RunArgs& R, void* dst, cudaStream_t st) {
  const int threads = 128;
  const unsigned grid = (unsigned)((R.batch + threads - 1) / threads);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const size_t smem = (size_t)std::min<int>(p->dev.n_ops, REG_SMEM_OPS) * sizeof(RegOp);
  // complex128, n = 4: three resident CTAs per SM (168 registers, 12 warps) measured 5.6 %
  // faster than two (210 registers, 8 warps): 0.214 vs 0.227 ms on config 2;
  // QMLB_REG_CTAS=2 selects the latter
  static const int want3 = [] {
    const char* v = std::getenv("QMLB_REG_CTAS");
    return v ? std::atoi(v) : 3;
  }();
  if (sizeof(QMLB_T) == 8 && N == 4 && want3 == 3)
    k_reg<QMLB_T, N, (sizeof(QMLB_T) == 8 && N == 4) ? 3 : 0>
        <<<grid, threads, smem, st>>>(p->dev, R, p->reg_mode, p->max_arg + 1, dst);
  else
    k_reg<QMLB_T, N><<<grid, threads, smem, st>>>(p->dev, R, p->reg_mode, p->max_arg + 1, dst);
}

cudaError_t QMLB_LAUNCH_REG(const qmlb_program* p, const RunArgs& R, void* dst,
                            cudaStream_t st) {
  switch (p->n_bits) {
    case 1: launch_n<1>(p, R, dst, st); break;
    case 2: launch_n<2>(p, R, dst, st); break;
    case 3: launch_n<3>(p, R, dst, st); break;
    case 4: launch_n<4>(p, R, dst, st); break;
    case 5: launch_n<5>(p, R, dst, st); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace qmlb
