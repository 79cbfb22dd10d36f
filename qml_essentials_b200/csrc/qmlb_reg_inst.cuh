// Instantiates the register kernel for one precision (QMLB_T / QMLB_SUFFIX).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "qmlb_internal.h"
#include "qmlb_reg.cuh"

namespace qmlb {

// Can the launch use the CTA-tiled factor staging?  The hoisted factors must hang on a FAST
// axis (div = 1), a SLOW axis (div = extent of the fast one) or constants (mod = 1), and the
// batch must be the full product of the two axes.  Fills Tl and the table bytes per CTA.
static bool reg_tile_plan(const qmlb_program* p, const RunArgs& R, RegTile& Tl, size_t& tab_bytes) {
  static const int enabled = [] {
    const char* v = std::getenv("QMLB_REG_TILED");
    return v ? std::atoi(v) : 1;
  }();
  if (!enabled || R.batch_offset != 0 || p->dev.n_ops > REG_SMEM_OPS || p->dev.n_ops < 1)
    return false;
  std::memset(&Tl, 0, sizeof(Tl));
  int64_t BP = 0, slow_div = 0, slow_mod = 0;
  bool any = false;
  for (const RegOp& d : p->reg_ops_host) {
    if (d.n < 1) continue;
    for (int j = 0; j < d.n; ++j) {
      const int slot = j ? d.slot1 : d.slot0;
      if (!R.pre_on[slot]) return false;
      const int64_t div = R.a[slot].div, mod = R.a[slot].mod;
      any = true;
      if (mod == 1) {
        Tl.cls[slot] = 3;
      } else if (div == 1) {
        if (BP && BP != mod) return false;
        BP = mod;
        Tl.cls[slot] = 1;
      } else {
        if (slow_div && (slow_div != div || slow_mod != mod)) return false;
        slow_div = div, slow_mod = mod;
        Tl.cls[slot] = 2;
      }
    }
  }
  if (!any) return false;
  if (!BP) BP = slow_div ? slow_div : R.batch;
  if (BP < 1 || R.batch % BP != 0) return false;
  const int64_t BI = R.batch / BP;
  if (slow_div && (slow_div != BP || slow_mod != BI)) return false;
  auto pow2ceil_bits = [](int64_t v) {
    int b = 0;
    while ((int64_t(1) << b) < v) ++b;
    return b;
  };
  int ti_bits = std::min(3, pow2ceil_bits(BI));
  int tp_bits = 7 - ti_bits;
  if (tp_bits > pow2ceil_bits(BP)) {
    tp_bits = pow2ceil_bits(BP);
    ti_bits = 7 - tp_bits;
  }
  // two evaluations per thread when the slow axis is long enough (QMLB_REG_REPS=1: one)
  static const int want_reps = [] {
    const char* v = std::getenv("QMLB_REG_REPS");
    return v ? std::atoi(v) : 2;
  }();
  Tl.reps = 1;
  if (want_reps == 2 && BI >= (int64_t(2) << ti_bits)) {
    Tl.reps = 2;
    ti_bits += 1;
  }
  Tl.tp_bits = tp_bits;
  Tl.ti_bits = ti_bits;
  Tl.BP = BP;
  Tl.BI = BI;
  Tl.n_ptiles = (BP + (int64_t(1) << tp_bits) - 1) >> tp_bits;
  size_t entries = 0;
  for (const RegOp& d : p->reg_ops_host)
    for (int j = 0; j < d.n; ++j) {
      const int c = Tl.cls[j ? d.slot1 : d.slot0];
      entries += c == 1 ? (size_t(1) << tp_bits) : (c == 2 ? (size_t(1) << ti_bits) : 1);
    }
  tab_bytes = entries * reg_tile_stride<QMLB_T>() * sizeof(cx<QMLB_T>);
  return true;
}

template <int N>
static cudaError_t launch_n(const qmlb_program* p, const RunArgs& R, void* dst, cudaStream_t st) {
  const int threads = 128;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const size_t ops_bytes = (size_t)std::min<int>(p->dev.n_ops, REG_SMEM_OPS) * sizeof(RegOp);
  // complex128, n = 4: more resident warps beat fewer spills - two CTAs per SM (210
  // registers) 0.227 ms, three (168) 0.214 ms on config 2 in the untiled form; in the tiled
  // form with parameter ops, four (128 registers, 176 bytes of spills) 0.1855 vs three 0.1895.
  // QMLB_REG_CTAS=2 / 3 select the others.
  static const int want3 = [] {
    const char* v = std::getenv("QMLB_REG_CTAS");
    return v ? std::atoi(v) : 4;
  }();
  constexpr int MB3 = (sizeof(QMLB_T) == 8 && N == 4) ? 3 : 0;
  const bool three = sizeof(QMLB_T) == 8 && N == 4 && want3 >= 3;
  RegTile Tl{};
  size_t tab_bytes = 0;
  if (reg_tile_plan(p, R, Tl, tab_bytes)) {
    const size_t smem =
        ((ops_bytes + 8 * (size_t)p->dev.n_ops + 15) & ~size_t(15)) + tab_bytes;
    if (smem <= 96 * 1024) {
      const int64_t n_itiles = (Tl.BI + (int64_t(1) << Tl.ti_bits) - 1) >> Tl.ti_bits;
      const unsigned grid = (unsigned)(Tl.n_ptiles * n_itiles);
      static const int want_po = [] {
        const char* v = std::getenv("QMLB_REG_PARAM_OPS");
        return v ? std::atoi(v) : 1;
      }();
      RegParamOps po{};
      const bool use_po = want_po && p->dev.n_ops <= REG_PARAM_OPS;
      if (use_po)
        std::memcpy(po.ops, p->reg_ops_host.data(), p->reg_ops_host.size() * sizeof(RegOp));
      // every instantiation has the same pointer type, so the "attribute set" flags are
      // kept per variant explicitly (one flag inside a generic lambda would be shared)
      static bool attr_set[4] = {false, false, false, false};
      cudaError_t attr_err = cudaSuccess;
      auto launch = [&](auto kern, bool& attr) {
        if (!attr) {
          attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          96 * 1024);
          attr = attr_err == cudaSuccess;
        }
        if (attr_err == cudaSuccess)
          kern<<<grid, threads, smem, st>>>(p->dev, R, p->reg_mode, p->max_arg + 1, dst, Tl, po);
      };
      static const int f4 = [] {
        const char* v = std::getenv("QMLB_REG_CTAS_F32");
        return v ? std::atoi(v) : 4;
      }();
      // complex64 n = 4: four CTAs (128 registers) 0.135 ms, three 0.159, five 0.136, six 0.168
      if (N == 4 && use_po && ((sizeof(QMLB_T) == 8 && want3 == 4) || (sizeof(QMLB_T) == 4 && f4 == 4))) {
        static bool a4 = false;
        launch(k_reg<QMLB_T, N, N == 4 ? 4 : 0, true, true>, a4);
      } else if (three && use_po)
        launch(k_reg<QMLB_T, N, MB3, true, true>, attr_set[0]);
      else if (three)
        launch(k_reg<QMLB_T, N, MB3, true, false>, attr_set[1]);
      else if (use_po)
        launch(k_reg<QMLB_T, N, 0, true, true>, attr_set[2]);
      else
        launch(k_reg<QMLB_T, N, 0, true, false>, attr_set[3]);
      if (attr_err != cudaSuccess) return attr_err;
      return cudaGetLastError();
    }
  }
  const unsigned grid = (unsigned)((R.batch + threads - 1) / threads);
  RegParamOps none{};
  if (three)
    k_reg<QMLB_T, N, MB3><<<grid, threads, ops_bytes, st>>>(p->dev, R, p->reg_mode,
                                                             p->max_arg + 1, dst, Tl, none);
  else
    k_reg<QMLB_T, N><<<grid, threads, ops_bytes, st>>>(p->dev, R, p->reg_mode, p->max_arg + 1,
                                                        dst, Tl, none);
  return cudaGetLastError();
}

cudaError_t QMLB_LAUNCH_REG(const qmlb_program* p, const RunArgs& R, void* dst,
                            cudaStream_t st) {
  switch (p->n_bits) {
    case 1: return launch_n<1>(p, R, dst, st);
    case 2: return launch_n<2>(p, R, dst, st);
    case 3: return launch_n<3>(p, R, dst, st);
    case 4: return launch_n<4>(p, R, dst, st);
    case 5: return launch_n<5>(p, R, dst, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace qmlb
