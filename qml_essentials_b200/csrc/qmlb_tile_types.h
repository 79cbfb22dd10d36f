// Pass descriptor shared by host scheduling code and the tile kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/qmlb200.h"

namespace qmlb {

#define QMLB_PASS_INIT 1   // tile starts as |0..0> instead of being loaded
#define QMLB_PASS_STORE 2  // tile is written back to global memory
#define QMLB_PASS_INIT_ZERO 8  // with INIT: start from the zero vector (a shard without index 0)
#define QMLB_PASS_HEAVY 4  // streaming pass with a dense / permutation op on 3-4 bits
#define QMLB_MAX_TILE_BITS 14

struct PassDev {
  const qmlb_op* ops;      // ops of this pass, bits rewritten to tile-local positions
  const int32_t* matoff;   // per op: offset (complex entries) into the team's matbuf
  const int2* windows;     // (first op, n ops) per window
  int32_t n_windows;
  int32_t k_tile;          // log2(tile amplitudes)
  int32_t n_bits;          // total state bits
  int32_t flags;
  int32_t matw;            // matbuf entries per team
  int32_t identity_map;    // tile-local bit j == global bit j for all j
  int32_t tile_bits[QMLB_MAX_TILE_BITS];  // local bit -> global bit (ascending)
};

}  // namespace qmlb
