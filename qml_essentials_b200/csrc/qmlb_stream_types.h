// Pass descriptor of the streaming kernel, shared by host scheduling and device code.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/qmlb200.h"
#include "qmlb_tile_types.h"

namespace qmlb {

constexpr int STREAM_THREADS = 128;
constexpr int STREAM_MAX_R = 5;
constexpr int STREAM_MIN_CTAS = 3;  // resident CTAs per SM the lean kernel is compiled for

constexpr int STREAM_MAX_OPS = 40;

// Compact op of a streaming pass; lives in the kernel parameters (constant bank), so
// decoding it costs uniform loads only.
struct StreamOp {
  uint8_t kind, k, b0, b1;  // b0/b1: register positions of bits[0]/bits[1] (k <= 2)
  int32_t src;              // matrix source (prologue only), -1 for PERM
  uint64_t data;            // PERM: table packed k bits/entry; DIAG: global bits, 6 bits each
};

// Fused global<->local qubit swap (qubit-sharded statevector): when enabled, the pass
// READS amplitude idx of the new local layout straight from the peer GPU that holds it
// (NVLink P2P loads through cp.async) and writes its own buffer - the exchange costs no
// HBM round trip of its own and overlaps with the arithmetic.  Chunk s = idx >> cshift
// lives on rank s at its chunk `rank`.
struct StreamPeers {
  const void* ptr[8];
  int32_t enabled;
  int32_t cshift;   // log2(amplitudes per chunk) = local bits - log2(ranks)
  int32_t rank;
  int32_t pad;
};

// one matrix of the per-element table filled by k_stream_mats (batched runs)
struct StreamMatOp {
  int32_t src;    // matrix source
  int32_t off;    // offset (complex entries) in the element's row
  int32_t swap2;  // 4x4 whose two local bits must swap roles (canonical register order)
  int32_t pad;
};

struct StreamPass {
  StreamOp ops[STREAM_MAX_OPS];
  uint16_t matoff[STREAM_MAX_OPS];  // per op: offset (complex entries) into the matrix buffer
  int32_t n_ops;
  int32_t n_bits;          // total state bits
  int32_t flags;           // QMLB_PASS_INIT | QMLB_PASS_HEAVY
  int32_t matw;            // matrix buffer entries
  int32_t mat_base;        // offset of this pass in a per-element row of precomputed matrices
  int32_t mat_row;         // entries per element in that table (sum of matw over passes)
  int32_t gb[STREAM_MAX_R];      // register bit j -> state bit
  int32_t sorted[STREAM_MAX_R];  // the same bits ascending
};

}  // namespace qmlb
