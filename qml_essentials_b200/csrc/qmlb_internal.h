// Internal declarations shared by the translation units of libqmlb200.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <string>
#include <vector>

#include "qmlb_device.cuh"
#include "qmlb_frame_types.h"
#include "qmlb_stream_types.h"
#include "qmlb_tile_types.h"

struct QmlbPassHost {
  std::vector<qmlb_op> ops;      // tile-local bits (DIAG keeps global bits)
  std::vector<int32_t> matoff;
  std::vector<int2> windows;
  std::vector<int> tile_bits;    // ascending global bits
  int flags = 0;
  int matw = 0;
  qmlb::PassDev dev{};           // device view (pointers into the program blob)
};

struct QmlbStreamPassHost {
  std::vector<qmlb_op> ops;      // register-position bits, packed PERM tables
  std::vector<int32_t> matoff;
  std::vector<int32_t> src_index;  // index of each op in the program's op list
  int gb[qmlb::STREAM_MAX_R] = {0, 0, 0, 0, 0};
  int sorted[qmlb::STREAM_MAX_R] = {0, 0, 0, 0, 0};
  int flags = 0;
  int matw = 1;
  qmlb::StreamPass dev{};
};

// one HBM pass of the streaming frame engine (strategy 4)
struct FramePassHost {
  int first_step = 0, n_steps = 0;
  bool init = false;           // the pass starts from |0..0> instead of reading the state
  std::vector<int> tp, opos;   // tile index / tile number bit -> HBM bit position (ascending)
};

struct qmlb_program {
  int n_qubits = 0, n_bits = 0, density = 0, dtype = 0, out_type = 0;
  std::vector<qmlb_op> ops;
  std::vector<qmlb_source> sources;
  std::vector<int32_t> items;
  std::vector<qmlb_angle> angles;
  std::vector<qmlb_term> terms;
  std::vector<double> consts;
  std::vector<qmlb_obs> obs;
  std::vector<double> obs_consts;
  std::vector<qmlb_pre> pre;
  std::vector<qmlb::RegOp> reg_ops_host;                    // strategy 0: the compact op stream
  std::vector<int32_t> pre_ids[QMLB_MAX_ARGS];        // pre entries per argument slot
  const int32_t* pre_ids_dev[QMLB_MAX_ARGS] = {};     // same, in the device blob
  int max_arg = -1;
  bool force_stream = false;

  int strategy = 0;
  bool direct_out = false;   // evolution kernel writes the final result itself
  int reg_mode = 0;
  bool warp_team = false;
  int teams = 1;
  size_t smem = 0;
  std::vector<QmlbPassHost> passes;
  std::vector<QmlbStreamPassHost> stream_passes;
  int stream_r = 4;
  std::vector<qmlb::StreamMatOp> stream_matlist;         // every matrix of every pass
  const qmlb::StreamMatOp* stream_matlist_dev = nullptr;
  int stream_mat_row = 0;                                // entries per element
  int sm_count = 148;

  // strategy 3: on-chip frame engine (qmlb_frame_types.h)
  std::vector<qmlb::FrameStep> frame_steps;
  std::vector<std::vector<int>> frame_step_ops;  // program-op indices per step
  const qmlb::FrameStep* frame_steps_dev = nullptr;
  qmlb::FrameProg frame{};
  int frame_threads = 0;
  int frame_out_mode = 0;
  bool frame_heavy = false;  // a dense op on 3-4 bits
  uint64_t frame_ptm_xmask = 0;  // strategy 5: physical positions of the x bits at the start
  size_t frame_smem = 0;

  // strategy 4: streaming frame engine (tiles of an HBM-resident state)
  std::vector<FramePassHost> fstream_passes;
  std::vector<int> fstream_final_hpos;   // logical bit -> HBM position after the last pass
  int fstream_low_bits = 0;

  void* blob = nullptr;
  qmlb::DevProg dev{};
};

namespace qmlb {

extern std::atomic<unsigned long long> g_launches;  // kernels launched by this library

constexpr int REG_MAX_BITS = 5;
constexpr int TILE_THREADS = 256;

// launchers (one translation unit per precision so the build parallelises)
cudaError_t launch_reg_f32(const qmlb_program* p, const RunArgs& R, void* dst, cudaStream_t st);
cudaError_t launch_reg_f64(const qmlb_program* p, const RunArgs& R, void* dst, cudaStream_t st);
cudaError_t launch_tile_f32(const qmlb_program* p, const RunArgs& R, const PassDev& pass,
                            unsigned grid, void* state, cudaStream_t st);
cudaError_t launch_tile_f64(const qmlb_program* p, const RunArgs& R, const PassDev& pass,
                            unsigned grid, void* state, cudaStream_t st);
#define QMLB_DECL_STREAM(name)                                                              \
  cudaError_t name(const qmlb_program* p, const RunArgs& R, const StreamPass& pass, dim3 grid, \
                   void* state, const void* premats, const StreamPeers* peers, cudaStream_t st)
QMLB_DECL_STREAM(launch_stream_f32_lean);
QMLB_DECL_STREAM(launch_stream_f32_heavy);
QMLB_DECL_STREAM(launch_stream_f64_lean);
QMLB_DECL_STREAM(launch_stream_f64_heavy);
#undef QMLB_DECL_STREAM

inline cudaError_t launch_stream_f32(const qmlb_program* p, const RunArgs& R,
                                     const StreamPass& pass, dim3 grid, void* state,
                                     const void* premats, const StreamPeers* peers,
                                     cudaStream_t st) {
  return (pass.flags & QMLB_PASS_HEAVY)
             ? launch_stream_f32_heavy(p, R, pass, grid, state, premats, peers, st)
             : launch_stream_f32_lean(p, R, pass, grid, state, premats, peers, st);
}
inline cudaError_t launch_stream_f64(const qmlb_program* p, const RunArgs& R,
                                     const StreamPass& pass, dim3 grid, void* state,
                                     const void* premats, const StreamPeers* peers,
                                     cudaStream_t st) {
  return (pass.flags & QMLB_PASS_HEAVY)
             ? launch_stream_f64_heavy(p, R, pass, grid, state, premats, peers, st)
             : launch_stream_f64_lean(p, R, pass, grid, state, premats, peers, st);
}
cudaError_t launch_stream_mats_f32(const qmlb_program* p, const RunArgs& R, void* out,
                                   cudaStream_t st);
cudaError_t launch_stream_mats_f64(const qmlb_program* p, const RunArgs& R, void* out,
                                   cudaStream_t st);
cudaError_t launch_frame_f32(const qmlb_program* p, const RunArgs& R, const void* premats,
                             void* out, int out_mode, cudaStream_t st);
cudaError_t launch_frame_f64(const qmlb_program* p, const RunArgs& R, const void* premats,
                             void* out, int out_mode, cudaStream_t st);
cudaError_t launch_fstream_f32(const qmlb_program* p, const RunArgs& R, void* state,
                               const void* premats, int init_mode, cudaStream_t st);
cudaError_t launch_fstream_f64(const qmlb_program* p, const RunArgs& R, void* state,
                               const void* premats, int init_mode, cudaStream_t st);
int plan_frame(qmlb_program* p);                 // QMLB_OK or QMLB_ERR_UNSUPPORTED (fall back)
int plan_frame_stream(qmlb_program* p);
int plan_frame_ptm(qmlb_program* p);
cudaError_t launch_frame_ptm_f32(const qmlb_program* p, const RunArgs& R, const void* premats,
                                 void* out, int out_mode, cudaStream_t st);
cudaError_t launch_frame_ptm_f64(const qmlb_program* p, const RunArgs& R, const void* premats,
                                 void* out, int out_mode, cudaStream_t st);
std::string describe_frame(const qmlb_program* p);
cudaError_t tile_set_smem_f32(size_t bytes);
cudaError_t tile_set_smem_f64(size_t bytes);

}  // namespace qmlb
