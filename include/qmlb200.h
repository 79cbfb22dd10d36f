/*
 * qmlb200.h - C ABI of the B200 (sm_100a) backend for the qml-essentials
 * circuit-execution hot path.
 *
 * The reference (cirKITers/qml-essentials, pure Python on JAX) has no FFI; its
 * de-facto backend seam is `Script.execute` (qml_essentials/script.py:137-147),
 * which traces the circuit under `jax.vmap` (script.py:302-329) and runs the
 * einsum kernels of qml_essentials/simulation.py.  This library replaces what
 * sits below that seam:
 *
 *   qmlb_program_create   <- script.py:272-329 `_build_plan` (trace + vmap + jit):
 *                            the host compiler hands over a flat program instead
 *   qmlb_run              <- script.py:358-397 `_dispatch` -> compiled(*args), i.e.
 *                            simulation.py:131-201 `simulate_and_measure`:
 *                            simulate_pure (:65-104), simulate_mixed (:107-128,
 *                            operations.py:485-512,1551-1578), measure_state
 *                            (:204-271), measure_density (:274-317)
 *   qmlb_sample           <- simulation.py:320-377 `sample_shots` (the integer
 *                            bookkeeping; the uniform stream is an input)
 *   qmlb_workspace_bytes  <- memory.py:54-139 `estimate_peak_bytes`
 *   qmlb_purity / qmlb_overlap_fidelity
 *                         <- entanglement.py:86-103 (Meyer-Wallach purities) and
 *                            expressibility.py:48-66 (pair fidelities), reduced on
 *                            device so only O(B) numbers leave the GPU
 *
 * Conventions.  One state index has `n_bits` bits.  Statevector programs:
 * n_bits = n_qubits, wire q is bit n_qubits-1-q (wire 0 = MSB, simulation.py:100).
 * Density programs: n_bits = 2*n_qubits, rho[i][j] lives at index i*2^n + j, so
 * ket wire q is bit 2n-1-q and bra wire q is bit n-1-q (operations.py:505-510).
 * The state starts as |0..0> (index 0 = 1).  A k-bit operation lists its bits
 * most-significant first: local value v = sum_j bit(bits[j]) << (k-1-j), and a
 * matrix M acts as new[v] = sum_u M[v][u] old[u] (operations.py:38-50).
 *
 * Ownership.  The caller owns every device buffer (arguments, output,
 * workspace); the library keeps no pointer past a call.  `qmlb_program` is an
 * opaque handle with explicit create/destroy.  All functions return QMLB_OK or
 * a negative code; `qmlb_last_error()` is a thread-local message.  Nothing
 * aborts.  All launches go to the given stream; no function synchronises
 * except qmlb_program_create/destroy and qmlb_fma_peak.
 */
#ifndef QMLB200_H
#define QMLB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QMLB_VERSION 100

#define QMLB_OK 0
#define QMLB_ERR_INVALID (-1)     /* malformed program / arguments            */
#define QMLB_ERR_UNSUPPORTED (-2) /* valid but not implemented by the kernels */
#define QMLB_ERR_CUDA (-3)        /* CUDA runtime error (message has details) */
#define QMLB_ERR_WORKSPACE (-4)   /* workspace too small                      */

/* precision of the evolved state */
#define QMLB_C64 0
#define QMLB_C128 1

/* operation kinds */
#define QMLB_OP_MAT 0   /* dense 2^k x 2^k matrix from `src`                       */
#define QMLB_OP_CTRL1 1 /* bits = {control, target}: 2x2 `src` on target if control */
#define QMLB_OP_PERM 2  /* new[v] = old[perm[v]], perm = consts[aux .. aux+2^k)      */
#define QMLB_OP_DIAG 3  /* new[v] = d[v] * old[v], d from `src` (2^k entries)        */

/* matrix sources */
#define QMLB_SRC_CONST 0  /* consts[a0]: 4^k complex (flags&2: 2^k complex diagonal)  */
#define QMLB_SRC_TRIG 1   /* C0 + cos(kappa*t) A + sin(kappa*t) B; a0,a1,a2; angle    */
#define QMLB_SRC_CHAIN 2  /* 2x2 product of sources items[a0 .. a0+a1), first acts 1st */
#define QMLB_SRC_DIAGPH 3 /* d[v] = exp(-i * marks[v] * t), marks = consts[a0] (real) */
#define QMLB_SRC_TABLE 4  /* per-element matrix: argument a0, complex offset a1,
                             flags&1 = conjugate                                      */
#define QMLB_SRC_SUPER 5  /* 4x4 superoperator on (ket bit, bra bit): product over
                             items[a0 .. a0+a1) of either a 2x2 source U (-> U (x)
                             conj U) or a constant 4x4 (k = 2)                        */

/* result types (script.py:151-159) */
#define QMLB_OUT_STATE 0   /* (B, 2^n) complex                       */
#define QMLB_OUT_PROBS 1   /* (B, 2^n) real                          */
#define QMLB_OUT_EXPVAL 2  /* (B, n_obs) real                        */
#define QMLB_OUT_DENSITY 3 /* (B, 2^n, 2^n) complex                  */

/* observables */
#define QMLB_OBS_ZSTRING 0 /* product of Z on the bits of zmask (n-bit index) */
#define QMLB_OBS_DIAG 1    /* diagonal: obs_consts[a0] holds 2^k complex      */
#define QMLB_OBS_DENSE 2   /* dense:    obs_consts[a0] holds 4^k complex      */

#define QMLB_SRC_PRE 6    /* hoisted 2x2 factor pre[a2]: depends on argument slot a1 only;
                             a0 = its index among the pre entries of that slot.  Read
                             from a per-row table filled by the precompute kernel when
                             the slot has fewer distinct rows than the batch, else
                             evaluated inline                                         */

/* source flags */
#define QMLB_FLAG_CONJ 1     /* TABLE: conjugate                                      */
#define QMLB_FLAG_DIAGVEC 2  /* CONST: 2^k diagonal entries instead of a matrix       */
#define QMLB_FLAG_ROT_SHIFT 2 /* TRIG, k = 1: bits 2-3 = 1/2/3 -> exactly RX/RY/RZ    */

/* qmlb_program_desc.reserved flags */
#define QMLB_DESC_FORCE_STREAM 1 /* always plan HBM-streaming passes (qmlb_evolve) */

#define QMLB_MAX_OP_BITS 8
#define QMLB_MAX_ARGS 8

typedef struct {
  int32_t kind, k, src, aux;
  int32_t bits[QMLB_MAX_OP_BITS];
} qmlb_op;

typedef struct {
  int32_t kind, k, a0, a1, a2, angle, flags, pad;
  double kappa;
} qmlb_source;

/* angle = c0 + sum over terms[first .. first+n) of coeff * arg[arg][row][offset] */
typedef struct {
  int32_t first, n;
  double c0;
} qmlb_angle;

typedef struct {
  int32_t arg, offset;
  double coeff;
} qmlb_term;

/* hoisted factor: 2x2 source `src` (elementary or chain of elementary sources)
 * that depends only on argument slot `arg`; `local` = index within that slot */
typedef struct {
  int32_t src, arg, local, pad;
} qmlb_pre;

typedef struct {
  int32_t kind, k, a0, pad;
  int64_t zmask;
  int32_t bits[QMLB_MAX_OP_BITS]; /* positions in the n-qubit index, MSB first */
} qmlb_obs;

/* Batched argument: float64 device matrix; element b (global batch index) reads
 * row ((b / div) % mod).  ptr may be NULL for unused slots. */
typedef struct {
  const double* ptr;
  int64_t stride; /* doubles per row */
  int64_t div, mod;
} qmlb_arg;

/* Host-side description of a program; everything is copied by create. */
typedef struct {
  int32_t n_qubits, n_bits, density, dtype, out_type, reserved;
  const qmlb_op* ops;
  int32_t n_ops;
  const qmlb_source* sources;
  int32_t n_sources;
  const int32_t* items;
  int32_t n_items;
  const qmlb_angle* angles;
  int32_t n_angles;
  const qmlb_term* terms;
  int32_t n_terms;
  const double* consts; /* complex entries interleaved (re, im) */
  int64_t n_consts;
  const qmlb_obs* obs;
  int32_t n_obs;
  const double* obs_consts;
  int64_t n_obs_consts;
  const qmlb_pre* pre;
  int32_t n_pre;
} qmlb_program_desc;

typedef struct qmlb_program qmlb_program;

int qmlb_version(void);
/* Number of CUDA kernels this library has launched in this process so far. */
unsigned long long qmlb_launch_count(void);
const char* qmlb_last_error(void);

/* Validates, schedules and uploads a program to the current CUDA device. */
int qmlb_program_create(const qmlb_program_desc* desc, qmlb_program** out);
int qmlb_program_destroy(qmlb_program* prog);

/* Host-only planning (no CUDA call, works without a GPU): validates `desc`, plans it and
 * writes a text description into buf - "strategy S", then for streamed programs one
 * line per fused gate pass: "pass flags F group b0 b1 .. ops i:kind:bits ..." where i is the
 * index of the op in desc->ops and bits are register positions (state bits for diagonal
 * ops); for the frame engines (strategies 3 - 5) a "frame ..." geometry line and one line
 * per step ("subpass ..." / "relayout ...", format in qmlb_frame_plan.cu:describe_frame).
 * Used by the CPU test-suite to check the schedulers (tests/_frame_emulator.py replays the
 * frame-engine steps in NumPy). */
int qmlb_plan_describe(const qmlb_program_desc* desc, char* buf, size_t buflen);

/* strategy: 0 = register-resident (one thread per circuit), 1 = shared-memory
 * resident (one warp / CTA per circuit), 2 = streamed fused gate passes over HBM
 * (register groups of 4 state bits, one launch per pass), 3 = on-chip frame engine (state
 * in shared memory / cluster DSMEM for the whole tape), 4 = streaming frame engine (tile
 * passes over an HBM-resident state, TMA bulk copies), 5 = Pauli-basis frame engine (noisy
 * density programs as real Pauli-coefficient vectors).
 * n_passes: state passes per run (strategies 2, 4) or steps of the on-chip step program
 * (strategies 3, 5); n_device_ops: ops after fusion. */
int qmlb_program_info(const qmlb_program* prog, int32_t* strategy, int32_t* n_passes,
                      int32_t* n_device_ops);

/* Bytes of scratch `qmlb_run` needs for `batch` elements with these arguments
 * (hoisted-factor tables + state + reduction partials; may be 0). */
size_t qmlb_workspace_bytes(const qmlb_program* prog, const qmlb_arg* args, int32_t n_args,
                            int64_t batch);

/* Evolves `batch` elements with global indices batch_offset .. batch_offset+batch-1
 * (the offset only enters the argument row computation, so a chunk or a rank's
 * shard can run against the full, unsliced arguments) and writes the result of
 * the program's out_type to `out` (row-major (batch, ...), real or complex of the
 * program precision).  `stream` is a cudaStream_t. */
int qmlb_run(const qmlb_program* prog, const qmlb_arg* args, int32_t n_args, int64_t batch,
             int64_t batch_offset, void* out, void* workspace, size_t workspace_bytes,
             void* stream);

/* Building blocks of the qubit-sharded statevector (nothing in the
 * reference corresponds - it cannot hold a state beyond host RAM).  A rank keeps one
 * shard of 2^n_bits amplitudes; the program (created with QMLB_DESC_FORCE_STREAM) holds
 * the ops of one epoch between two global<->local exchanges, on local bits.
 *
 * qmlb_evolve applies the program's fused gate passes IN PLACE to `state`
 * ((batch, 2^n_bits) complex of the program precision).  init_mode: 0 = continue from
 * the amplitudes in `state`, 1 = start from |0..0>, 2 = start from the zero vector (a
 * shard that does not contain index 0).  workspace: qmlb_workspace_bytes(...) bytes. */
int qmlb_evolve(const qmlb_program* prog, const qmlb_arg* args, int32_t n_args, int64_t batch,
                int64_t batch_offset, void* state, int32_t init_mode, void* workspace,
                size_t workspace_bytes, void* stream);

/* The fused form of "exchange, then evolve": the FIRST gate pass of the epoch reads every
 * amplitude of the post-swap layout directly from the peer GPU that holds it (NVLink
 * peer loads issued by the gate-pass kernel itself) and writes `dst_state`; the
 * remaining passes run in place on `dst_state`.  peer_src[s] is the address, valid in
 * this process, of rank s's pre-swap shard (peer_src[rank] is this rank's own);
 * n_peers = 2, 4 or 8.  The swap exchanges the top log2(n_peers) local bits with the
 * rank-index bits.  The caller orders the ranks (every rank must have finished writing
 * its pre-swap shard, and `dst_state` must not be a buffer peers still read). */
int qmlb_evolve_peer(const qmlb_program* prog, const qmlb_arg* args, int32_t n_args,
                     void* dst_state, const void* const* peer_src, int32_t n_peers,
                     int32_t rank, void* workspace, size_t workspace_bytes, void* stream);

/* One sweep over pure states: out[b][q] = sum of |amp|^2 over indices with bit q set
 * (q < n_bits), out[b][32] = total probability; double, (batch, 33).  <Z> of the qubit
 * on bit q is out[32] - 2 out[q]; a sharded state adds the per-rank tables (and `total`
 * for the rank-index bits that are set).  8 <= n_bits <= 32. */
size_t qmlb_zsums_workspace_bytes(int64_t batch, int32_t n_bits);
int qmlb_zsums(const void* state, int dtype, int64_t batch, int32_t n_bits, double* out,
               void* workspace, size_t workspace_bytes, void* stream);

/* Shot bookkeeping of simulation.py:352-357 for `batch` probability vectors of
 * length 2^n_qubits (float64 when dtype == QMLB_C128, else float32):
 *   p_cuml = cumsum(p) (sequential order);  r = p_cuml[last] * (1 - u);
 *   index  = number of entries of p_cuml strictly below r, clamped to 2^n - 1;
 *   counts[b][index] += 1          (int32, zeroed by the call)
 * uniforms: (batch, shots) float64 in [0, 1). */
int qmlb_sample(const void* probs, int dtype, const double* uniforms, int64_t batch,
                int32_t n_qubits, int64_t shots, int32_t* counts, void* stream);

/* Meyer-Wallach purities out[b][q] = Tr[(Tr_q rho_b)^2] - the purity of the state
 * with qubit q traced out (entanglement.py:86-103) - of `batch` pure states
 * (is_density = 0, (batch, 2^n); computed from the single-qubit reduced state, which
 * has the same purity) or density matrices (is_density = 1, (batch, 2^n, 2^n)).
 * out: (batch, n_qubits) real of the given precision. */
int qmlb_purity(const void* states, int dtype, int is_density, int64_t batch,
                int32_t n_qubits, void* out, void* stream);

/* |<psi_b | psi_{b+half}>|^2 for b < half over (2*half, 2^n) pure states - the pair
 * fidelity of expressibility.py:48-66 for pure states.  out: (half,) real. */
int qmlb_overlap_fidelity(const void* states, int dtype, int64_t half, int32_t n_qubits,
                          void* out, void* stream);

/* Device post-processing of the Fourier-analysis callers.
 *
 * qmlb_grid_dft: the transform of Coefficients._fourier_transform (coefficients.py:128-150)
 * for ONE input feature.  ev: (n_x, n_p, n_obs) real expectation values (real of the given
 * precision) as `Model.__call__` returns them for an n_x-point input grid and n_p parameter
 * samples; coefficient k = (1 / n_x) sum_x mean_obs(ev[x][p][:]) exp(-2 pi i k x / n_x) for
 * k = 0 .. n_x - 1 (numpy.fft order) is written to out[row_of[k]][p] (row_of == NULL: row k;
 * row_of[k] < 0: dropped) - a device array of n_x entries that lets the caller fold
 * get_spectrum's fftshift / trim (coefficients.py:72-84) into the store.  out: (rows, n_p)
 * complex of the given precision. */
int qmlb_grid_dft(const void* ev, int dtype, int32_t n_x, int64_t n_p, int32_t n_obs,
                  const int32_t* row_of, void* out, void* stream);

/* Additive sufficient statistics of the FCC correlation estimators
 * (coefficients.py:1346-1498) over the n_p samples of K selected coefficient rows
 * (rows[i] indexes the first axis of coef, a (*, n_p) complex array of the given
 * precision): out (complex128, 2K + K*K entries) = [sum_p c_i | sum_p |c_i|^2 |
 * sum_p conj(c_i) c_j].  These are what crosses ranks in a batch-sharded run. */
int qmlb_coef_moments(const void* coef, int dtype, const int32_t* rows, int32_t K, int64_t n_p,
                      void* out, void* stream);

/* One-shot all-reduce (sum, float64) over peer-mapped symmetric buffers: peer_buf[r] is the
 * address, valid in this process, of rank r's buffer of qmlb_allreduce_buffer_bytes(n) bytes
 * (zero-initialised once); `in` and `out` are local, n doubles each.  One launch per rank
 * on n_peers (2..8) GPUs; every rank obtains the same bits (rank-order summation).
 * mode 0: synchronous - out = sum over ranks of this call's `in`.  mode 1: pipelined - out =
 * the reduction of the PREVIOUS call (zeros for the first), so launch skew between GPUs is
 * not serialised into every step.  mode 2: drain - contributes nothing and returns the
 * reduction of the last call (closes a pipelined sequence).  If a peer does not arrive
 * within ~2^28 polls the result is NaN instead of a hang. */
size_t qmlb_allreduce_buffer_bytes(int64_t n);
int qmlb_allreduce_peer(const void* const* peer_buf, int32_t n_peers, int32_t rank, int64_t n,
                        const double* in, double* out, int32_t mode, void* stream);

/* Sub-register outputs (jaqsi.py:79-146, applied by Model._forward, model.py:1711-1724):
 * qmlb_partial_trace keeps the qubits `keep[0..k)` (wire numbers, ascending) of `batch`
 * density matrices (batch, 2^n, 2^n) -> (batch, 2^k, 2^k); qmlb_marginal_probs the same for
 * probability vectors (batch, 2^n) -> (batch, 2^k).  Output index bits follow the ascending
 * wire order, wire keep[0] most significant, like the reference.  n <= 16. */
int qmlb_partial_trace(const void* rho, int dtype, int64_t batch, int32_t n_qubits,
                       const int32_t* keep, int32_t k, void* out, void* stream);
int qmlb_marginal_probs(const void* probs, int dtype, int64_t batch, int32_t n_qubits,
                        const int32_t* keep, int32_t k, void* out, void* stream);

/* Measurement aid: sustained FMA throughput of this GPU in the given real
 * precision (TFLOP/s), used as the roofline denominator of the register-resident
 * regime.  Blocks until done. */
int qmlb_fma_peak(int dtype, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* QMLB200_H */
