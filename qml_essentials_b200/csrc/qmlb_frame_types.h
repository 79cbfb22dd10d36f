// Step program of the on-chip "frame" engine (strategy 3), shared by the host planner
// (qmlb_frame_plan.cu) and the kernel (qmlb_frame.cuh).
//
// The whole state of one circuit evaluation - a statevector of up to 2^17 amplitudes or the
// 4^n entries of a density matrix (BASELINE config 4: n = 8, 1 MiB in complex128) - stays
// in shared memory for the entire tape: in ONE CTA when it fits 128 KiB, else distributed
// over the shared memories of a thread-block cluster of 2^G CTAs (DSMEM).  The planner cuts
// the device program into STEPS:
//
//   SUBPASS   every thread takes one item = 2^4 amplitudes (a register group of four
//             logical bits), applies every op of the step that acts inside the group with
//             compile-time register indices, and writes them back in place;
//   RELAYOUT  the physical layout changes (global<->local bit exchange across the cluster
//             through distributed shared memory, or the final return to index order).
//
// CX / SWAP and every other GF(2)-linear permutation cost NOTHING at run time: the planner
// folds them into the FRAME, the invertible bit matrix A with physical = A * logical.  A
// gate on logical bit j then pairs the physical addresses p and p ^ col_j(A) (its MASK) and
// the logical value of bit j at address p is parity(p & row_j(A^-1)).  A register group is
// therefore described by its masks (eoff = all XOR combinations), the pivot positions of
// the span of the masks (items enumerate the indices that are zero there) and the parity
// rows; when a parity row reaches outside the group, slot 0 of an item would hold local
// value c instead of 0 - the item then simply starts at base ^ eoff[c] (eoff is linear in
// the slot number), after which slot v holds logical value v for every item.
#pragma once
#include <stdint.h>

namespace qmlb {

constexpr int FRAME_R = 4;          // register-group bits
constexpr int FRAME_D = 1 << FRAME_R;
constexpr int FRAME_MAX_OPS = 24;   // op slots per step (a DIAG op takes two)
constexpr int FRAME_MAX_PAR = 16;   // parity rows per step (0..3 = the register bits)
constexpr int FRAME_MAX_BITS = 40;

#define QMLB_FSTEP_SUBPASS 0
#define QMLB_FSTEP_RELAYOUT 1

#define QMLB_FOP_MAT1 0   // 2x2 on register bit j0
#define QMLB_FOP_MAT2 1   // 4x4 on register bits j0 > j1 (local value = bit j0 << 1 | bit j1)
#define QMLB_FOP_MATK 2   // dense 2^k x 2^k on register bits k-1..0 (k = 3, 4)
#define QMLB_FOP_CTRL1 3  // 2x2 on register bit j0 where parity row j1 reads 1
#define QMLB_FOP_DIAG 4   // diagonal over k parity rows (their indices sit in the next slot)
#define QMLB_FOP_SIGN 5   // +-1 over 4 parity rows: negate where bit (local value) of the mask in
                          // premat_off is set (Pauli-basis engine: the sign part of a Clifford).
                          // Four slots: the op (j0, j1, shape, has_c = the parity-row indices,
                          // smem_off = their four rout bytes), the four rloc words, and 16 sign
                          // words (entry = local value at slot 0, bit v = sign of slot v)

#define QMLB_OP_SIGN 4    // planner-internal op kind behind QMLB_FOP_SIGN (never in a user program)

#define QMLB_FSHAPE_FULL 0   // general complex matrix
#define QMLB_FSHAPE_REAL 1   // every entry real (RY chains, Pauli channels): half the FMAs
#define QMLB_FSHAPE_XREAL 2  // 4x4, real, only v == u and v == u ^ 3 (depolarizing, flips, damping)
#define QMLB_FSHAPE_PDIAG 3  // Pauli-basis engine: the 4x4 transfer matrix is diagonal

struct FrameOp {           // 16 bytes
  uint8_t code, k, j0, j1;
  uint8_t shape;           // QMLB_FSHAPE_*: structure known at plan time
  uint8_t has_c;           // some register bit of the op can see a flipped local value
  uint16_t flags;          // bit 0: MAT2 whose two logical bits sit in reversed register order
  int32_t premat_off;      // offset of the op's matrix in the element's row of evaluated matrices
  int32_t smem_off;        // offset (complex entries) of the matrix in shared memory
};

// logical bit value at physical index (outer, local) of slot v of an item:
//   parity(local_base & rloc) ^ parity(outer & rout) ^ bit v of smask
struct FramePar {
  uint32_t rloc, rout;
  uint16_t smask, pad;
};

struct FrameStep {  // 1024 bytes, loaded into shared memory by the CTA that runs it
  int32_t kind, n_ops, mat_entries, n_par;  // RELAYOUT: mat_entries = 1 -> tile-local shuffle
  uint32_t pivots[FRAME_R];           // ascending: items have zeros at these tile positions
  uint32_t eoff[FRAME_D];             // slot v lives at tile index base ^ eoff[v]
  FramePar par[FRAME_MAX_PAR];
  FrameOp ops[FRAME_MAX_OPS];
  // RELAYOUT: the amplitude that ends at physical index d comes from physical index
  // XOR of qcol[b] over the set bits b of d (bits >= tile_bits select the CTA of the cluster)
  uint64_t qcol[FRAME_MAX_BITS];
  // SUBPASS fast paths (straight-line item bodies, no op interpreter):
  //   fast = 16 + 4 * (sA + 1) + (sB + 1): at most one 4x4 on register pair (1,0) with shape
  //          sA and one on (3,2) with shape sB (-1 = absent); foff[0] / foff[1] = their
  //          matrix offsets
  //   fast = 64 + 16 * real + mask: only 2x2 ops, at most one per register bit (mask), all
  //          real (1) or treated as full (0); foff[j] = matrix offset of the op on bit j
  //   fast = 128 + 3 * a + b (Pauli-basis engine): [signs] A [signs] B with A the transfer
  //          matrix on pair (1,0), B the one on (3,2); a, b = 0 absent, 1 full, 2 diagonal;
  //          foff[0] / foff[1] = their offsets, FrameSubX::sg = the sign slots
  //   fast = 0: generic interpreter
  int32_t fast;
  int32_t foff[4];
  uint32_t eoffb_unused[3];
};
static_assert(sizeof(FrameStep) == 1024, "FrameStep is loaded as 256 words");

// SUBPASS extras, stored over the qcol area (which only a RELAYOUT reads).  Item number ->
// tile index: item bit b sits at tile position ipos[b].  The planner picks the lowest item
// bits (the lanes of one shared-memory wavefront) so that the lanes land in distinct banks
// of the SWIZZLED tile (qmlb_frame_ptm.cuh); kd[b] = address delta of item bit team_bits + b
// (a thread's further items), the item shift eoff[c] included.
struct FrameSubX {
  uint8_t ipos[16];
  uint32_t kd[4];
  uint32_t lanes_ok;  // 1: the lane bits are conflict-free
  // Pauli-basis fast path (FrameStep::fast >= 128): slots of the sign ops that run before the
  // transfer matrix on register pair (1,0) (sg[0], sg[1]) and between it and the one on pair
  // (3,2) (sg[2], sg[3]); 0xff = none
  uint8_t sg[4];
};
static_assert(sizeof(FrameSubX) <= sizeof(uint64_t) * FRAME_MAX_BITS, "fits the qcol area");

struct FrameProg {
  const FrameStep* steps;
  int32_t n_steps;
  int32_t n_bits;      // logical state bits (n or 2n)
  int32_t tile_bits;   // T: bits held by one CTA
  int32_t outer_bits;  // G: log2(cluster size)
  int32_t team_bits;   // log2(threads working on one tile)
  int32_t teams;       // tiles per CTA (1 when the CTA or cluster holds one state)
  int32_t mat_cap;     // matrix entries per team in shared memory
  int32_t mat_resident;  // 1: the element's whole row of matrices is staged once per element
  int32_t premat_row;  // evaluated-matrix entries per element
  int32_t out_mode;    // 0: complex state in index order, 1: probabilities, 2: Z-string expvals
  int32_t density, n_qubits, n_obs;
  int32_t ptm;         // 1: the state is the REAL Pauli-coefficient vector of a density matrix
  // several tiles per CTA (small states): every step record is loaded once per CTA instead of
  // once per round and step; tiles sit tile_pitch elements apart (2^T + 1: the teams of a warp
  // then hit different banks); mat_resident = 2 reads matrices straight from the element's
  // row in global memory (they are used once - staging them would cost the shared memory
  // that limits the resident warps)
  int32_t steps_resident, tile_pitch;
};

}  // namespace qmlb
