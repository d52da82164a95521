/*
 * qsim_b200.h -- C ABI of libqsim_b200.so, the B200-native (sm_100a) state-vector
 * engine that replaces qsim's CUDA Simulator / StateSpace / VectorSpace
 * backends (reference: quantumlib/qsim @ 502b4a37, paths relative to its root).
 *
 * Plain pointers and sizes only.  Every entry point names the reference
 * interface it replaces.  The header-only C++ mirror of qsim's duck-typed
 * backend API (include/qsim_b200/simulator_b200.h, statespace_b200.h,
 * vectorspace_b200.h) and the Python mirror (qsim_b200/) are both thin shims
 * over this file.
 *
 * STATE LAYOUT.  A state of n qubits is a device array of 2*2^n scalars
 * (float or double) in qsim's *normal order*: amplitude i = (s[2i], s[2i+1]),
 * qubit q <-> bit q of i (lib/simulator.h:40-66).  This is the layout
 * StateSpaceBasic uses on the CPU (lib/statespace_basic.h:84-105) and the one
 * qsim's pybind layer hands to Python, so InternalToNormalOrder /
 * NormalToInternalOrder (lib/statespace_cuda.h:85-107) are no-ops here.
 *
 * GATE MATRICES are caller-owned HOST pointers, row-major 2^G x 2^G,
 * interleaved (re,im), same scalar type as the state, bit k of the row/column
 * index <-> qs[k] (lib/matrix.h:26-33); they are consumed before the call
 * returns.  qs must be sorted ascending like the reference requires
 * (lib/simulator_cuda.h:72).
 *
 * ERRORS.  Every function returns a qb200_status.  Nothing prints or exits:
 * the C++ shim maps QB200_ERR_CUDA to the reference's "print + exit"
 * convention (lib/util_cuda.h:31-39) and QB200_ERR_OOM to Null()
 * (lib/vectorspace_cuda.h:90-95).
 *
 * STREAMS.  All work of a context is enqueued on that context's stream
 * (default: the legacy default stream, like the reference).  Gate application
 * and state initialisation are asynchronous; calls that return a value to the
 * host synchronise the stream.
 */
#ifndef QSIM_B200_H_
#define QSIM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  QB200_OK = 0,
  QB200_ERR_CUDA = 1,        /* a CUDA runtime call failed; see qb200_last_cuda_error */
  QB200_ERR_OOM = 2,         /* device allocation failed */
  QB200_ERR_INVALID = 3,     /* bad argument (qubit out of range, mismatch ...) */
  QB200_ERR_UNSUPPORTED = 4  /* gate size outside the reference's limits; state untouched */
} qb200_status;

typedef enum { QB200_F32 = 0, QB200_F64 = 1 } qb200_dtype;

typedef struct qb200_ctx qb200_ctx;

/* ---- library / context ------------------------------------------------- */
/* ABI version of this header (bumped on any signature change). */
int qb200_abi_version(void);
int qb200_device_count(int* count);
/* Per-object resources: stream handle, reduction scratch, pinned result slot.
 * Replaces the members SimulatorCUDA / StateSpaceCUDA own
 * (lib/simulator_cuda.h:52-62,899-915; lib/statespace_cuda.h:378-390).
 * device < 0 means "the current device".  Cheap: allocations are lazy. */
int qb200_ctx_create(int device, qb200_ctx** ctx);
int qb200_ctx_destroy(qb200_ctx* ctx);
/* stream is a cudaStream_t; NULL selects the legacy default stream. */
int qb200_ctx_set_stream(qb200_ctx* ctx, void* stream);
/* cudaError_t of the last failed runtime call on this context, and its text. */
int qb200_last_cuda_error(const qb200_ctx* ctx);
const char* qb200_last_cuda_error_string(const qb200_ctx* ctx);
/* Number of kernels this context has launched (bench.py's gpu_launches). */
uint64_t qb200_launch_count(const qb200_ctx* ctx);
/* Name of the gate / expectation kernel the dispatcher chose for the LAST pass of this context ("k_gate_tca<4>",
 * "k_gate_tile<4>", "k_gate_reg<2>", ...): bench.py labels its per-kernel roofline with it instead of
 * re-stating the dispatch rule.  Static storage; "" before the first pass. */
const char* qb200_last_kernel_name(const qb200_ctx* ctx);
/* Process-wide counter that advances with every library call that may have written a state (gate passes, the Set... calls, Add,
 * Multiply, Collapse, copies into device memory, exchanges of sharded states, frees and allocations).  A cache of values
 * derived from a state (the operator groups of include/qsim_b200/simulator_b200.h) is valid while it stands still.
 * Writes to the raw device pointer that bypass this library are not seen. */
uint64_t qb200_mutation_epoch(void);
/* Persistent grids of this context are sized for `sms` SMs instead of all 148 (0 = all): used while an
 * exchange kernel of a sharded state owns the remaining SMs (csrc/sharded.cu). */
int qb200_ctx_set_sm_limit(qb200_ctx* ctx, int sms);
/* Persistent grids of this context count on `ctas_per_sm` fewer resident CTAs per SM (0 = all): set while an exchange
 * kernel runs beside the gates and holds registers, threads and shared memory on every SM. */
int qb200_ctx_set_occupancy_reduction(qb200_ctx* ctx, int ctas_per_sm);
/* Kernel-selection overrides for experiments and for the cross-check tests (defaults = -1 = auto):
 *   gate_mode 0      one amplitude per access in the register kernels
 *   force_generic 1  runtime-generic kernel for everything
 *   tile 0/1/2       fp32 G=4 FFMA2 path: register kernels only / cp.async ring only / warp tile wherever legal
 *   prefetch 0       no software-pipelined persistent loop in the register kernels
 *   big 0/1          fp32 G=5,6 FFMA2 path: register+generic kernels / alternative launch shape
 *   tc 0..6          fp32 G=4,5 gates on the tensor cores: 0 off; 1,2 operands staged in shared memory;
 *                    3 = default data path (A operand in TMEM) forced for every layout; 4 without the
 *                    accumulation-bias compensation term; 5,6 with a cp.async staging ring
 *   tc_low k         replace the per-layout G=4 tensor-core rule by: lowest non-zero target >= k
 *   tcx 0            keep G=6 gates and G=4..6 expectation values on the FFMA2 kernels
 *   mono 0           dense kernels also for XOR-monomial operators (Pauli strings) in qb200_expectation_value
 *   expect_ug 2/3    several groups per thread per iteration in the G<=2 expectation kernel (slower; measured)
 *   tc_comp6 v       compensation constant of the G=6 tensor-core gate in units of 1e-9 (default 276)
 * Unknown keys return QB200_ERR_INVALID.  See DESIGN.md section 3. */
int qb200_ctx_set_tuning(qb200_ctx* ctx, const char* key, int value);
/* Event timing on the context's stream (CUDA events). */
int qb200_timer_start(qb200_ctx* ctx);
int qb200_timer_stop_ms(qb200_ctx* ctx, float* ms);

/* ---- VectorSpace (lib/vectorspace_cuda.h:43-167) ------------------------ */
/* StateSpaceCUDA::MinSize (lib/statespace_cuda.h:81-83): scalars per state. */
uint64_t qb200_min_size(unsigned num_qubits);
/* VectorSpaceCUDA::Create(n) (:87-96).  QB200_ERR_OOM on failure.  The reference's Create is static and
 * allocates on the CURRENT device; qb200_state_alloc keeps that meaning, qb200_state_alloc_on allocates on the
 * context's device (what a caller driving several GPUs from one process wants).  Kernels of a context refuse a
 * state that lives on another GPU (QB200_ERR_INVALID). */
int qb200_state_alloc(unsigned num_qubits, int dtype, void** state);
int qb200_state_alloc_on(qb200_ctx* ctx, unsigned num_qubits, int dtype, void** state);
/* detail::free (:31-33) */
int qb200_state_free(void* state);
/* Copy state->state / state->host / host->state (:112-160).  `count` is in
 * scalars; host buffers may be pageable.  Blocking, like the reference. */
int qb200_copy_d2d(qb200_ctx* ctx, int dtype, const void* src, void* dst, uint64_t count);
/* the same device-to-device copy enqueued on the context's stream without waiting for it */
int qb200_copy_d2d_async(qb200_ctx* ctx, int dtype, const void* src, void* dst, uint64_t count);
int qb200_copy_d2h(qb200_ctx* ctx, int dtype, const void* src, void* host_dst, uint64_t count);
int qb200_copy_h2d(qb200_ctx* ctx, int dtype, const void* host_src, void* dst, uint64_t count);
/* VectorSpaceCUDA::DeviceSync (:162-164): waits for the context's stream. */
int qb200_sync(qb200_ctx* ctx);
/* The reference's static DeviceSync: cudaDeviceSynchronize on the current device. */
int qb200_device_sync(void);
/* cudaDeviceSynchronize on a given device (a multi-device state's DeviceSync loops over the devices). */
int qb200_device_sync_on(int device);

/* ---- Simulator (lib/simulator_cuda.h) ---------------------------------- */
/* SimulatorCUDA::ApplyGate (:70-125): in place, num_targets in [0,6].
 * More than 6 targets -> QB200_ERR_UNSUPPORTED (the reference ignores them). */
int qb200_apply_gate(qb200_ctx* ctx, int dtype, void* state, unsigned num_qubits,
                     const unsigned* qs, unsigned num_targets, const void* matrix);
/* SimulatorCUDA::ApplyControlledGate (:135-207): bit i of cvals is the required
 * value of the i-th lowest control qubit (lib/simulator.h:364-375).
 * num_targets in [0,4] like the reference (:162-164); num_controls == 0
 * forwards to qb200_apply_gate. */
int qb200_apply_controlled_gate(qb200_ctx* ctx, int dtype, void* state, unsigned num_qubits,
                                const unsigned* qs, unsigned num_targets,
                                const unsigned* cqs, unsigned num_controls, uint64_t cvals,
                                const void* matrix);
/* PRECISION CONTRACT of fp32 gates.  Gates of up to 3 qubits and every fp64 gate: products and sums in the state's
 * precision (FMA), like the reference's CPU path.  fp32 gates of 4, 5 and 6 qubits run on the tensor cores as a
 * 3xTF32 split (x = hi + lo, hi*hi + hi*lo + lo*hi; lo*lo, ~2^-22 relative, dropped) with fp32 accumulation; the
 * tensor core's truncating accumulation loses ~1.7e-7 (G=4) .. 5.5e-7 (G=6) of the norm per pass on dense
 * unitaries, which a fitted compensation term cancels to < 1e-9 per pass (csrc/gate_tc.cuh, tools/tc_check.py).
 * Matrices with at most one non-zero per row and column (permutations, Pauli strings, diagonal phases) get no
 * compensation: permutations with entries in {0, +-1, +-i} reproduce every amplitude to the 22 significant bits of
 * the hi + lo split (exact on basis states and on amplitudes with <= 22 significant bits), diagonal phases lose
 * < 1.5e-7 of the norm per pass (measured 6e-8).  Measured against the reference CPU path:
 * per-amplitude error 2e-8 on normalised states, 6e-11 on the amplitudes of circuit_q30; tuning "tc" = 0 keeps
 * everything on the FFMA2 kernels.
 *
 * SimulatorCUDA::ExpectationValue (:216-260): <psi|M|psi>, num_targets in [1,6],
 * products in the state's precision, accumulation in double.  Synchronises.
 * Matrices with one non-zero per row in column r ^ const (Pauli strings, products of phase gates) on 3..6
 * targets are evaluated as a single read pass without the mat-vec (csrc/expect_monomial.cu). */
int qb200_expectation_value(qb200_ctx* ctx, int dtype, const void* state, unsigned num_qubits,
                            const unsigned* qs, unsigned num_targets, const void* matrix,
                            double out_re_im[2]);
/* Batched reductions (no reference counterpart: lib/expect.h:106-151 and
 * lib/qtrajectory.h read one expectation value at a time, one stream synchronisation each).
 * Between begin and end every reduction of this context -- qb200_expectation_value,
 * qb200_norm, qb200_inner_product, qb200_real_inner_product -- is enqueued on the stream and
 * returns at once with its `out` set to NaN; its result goes to the next slot of a mapped
 * pinned host array.  `end` synchronises once and copies the `*count` results, (re, im) per
 * slot in call order, to `out` (capacity in slots; QB200_ERR_INVALID if it is too small).
 * `expected` only pre-sizes the slot array; it grows on demand.  An operator on more than 6 qubits takes a
 * slot holding 0 (the reference's answer for it), so results pair with calls by position.  qb200_collapse
 * needs its norm on the host at once and returns QB200_ERR_INVALID inside a batch. */
/* Reduced density matrices of ALL qubits in a handful of read-only passes (3 at 26 qubits, 4 at 30) instead
 * of one pass per single-qubit operator (no reference counterpart; csrc/moments.cu).  out[4q .. 4q+3] =
 * S00, S11, Re S01, Im S01 of qubit q with S00 / S11 = sum |a_i|^2 over bit_q(i) = 0 / 1 and
 * S01 = sum conj(a_i0) a_i1 over the pairs i1 = i0 | 1 << q, so that for any 2x2 operator M on q
 * <psi|M|psi> = m00 S00 + m11 S11 + m01 S01 + m10 conj(S01).  Products in the state's precision,
 * accumulation in double.  Synchronises.  QB200_ERR_UNSUPPORTED for a state that is not 16-byte aligned. */
int qb200_one_qubit_moments(qb200_ctx* ctx, int dtype, const void* state, unsigned num_qubits, double* out);
int qb200_reduce_batch_begin(qb200_ctx* ctx, uint32_t expected);
int qb200_reduce_batch_end(qb200_ctx* ctx, double* out, uint32_t capacity, uint32_t* count);

/* ---- StateSpace (lib/statespace_cuda.h, lib/statespace.h) -------------- */
int qb200_set_all_zeros(qb200_ctx* ctx, int dtype, void* state, unsigned num_qubits);      /* :109-112 */
int qb200_set_state_zero(qb200_ctx* ctx, int dtype, void* state, unsigned num_qubits);     /* :130-135 */
int qb200_set_state_uniform(qb200_ctx* ctx, int dtype, void* state, unsigned num_qubits);  /* :115-127 */
int qb200_get_ampl(qb200_ctx* ctx, int dtype, const void* state, uint64_t i, double out_re_im[2]); /* :138-145 */
int qb200_set_ampl(qb200_ctx* ctx, int dtype, void* state, uint64_t i, double re, double im);      /* :148-164 */
/* BulkSetAmpl (:166-187): state[i] = (re,im) where ((i & mask) == bits) ^ exclude. */
int qb200_bulk_set_ampl(qb200_ctx* ctx, int dtype, void* state, unsigned num_qubits,
                        uint64_t mask, uint64_t bits, double re, double im, int exclude);
int qb200_add(qb200_ctx* ctx, int dtype, const void* src, void* dest, unsigned num_qubits); /* :189-204 */
int qb200_multiply(qb200_ctx* ctx, int dtype, double a, void* state, unsigned num_qubits);  /* :206-217 */
/* InnerProduct = sum conj(s1) s2 (:219-229, lib/util_cuda.h:108-114), RealInnerProduct
 * (:231-237), Norm (:239-241).  Double accumulation, run-to-run deterministic. */
int qb200_inner_product(qb200_ctx* ctx, int dtype, const void* s1, const void* s2,
                        unsigned num_qubits, double out_re_im[2]);
int qb200_real_inner_product(qb200_ctx* ctx, int dtype, const void* s1, const void* s2,
                             unsigned num_qubits, double* out);
int qb200_norm(qb200_ctx* ctx, int dtype, const void* state, unsigned num_qubits, double* out);
/* <psi|M_i|psi> for `count` (<= 8) operators on the SAME one or two qubits in ONE read pass (csrc/expect_multi.cu):
 * matrices = count row-major 2^G x 2^G interleaved (re, im) matrices back to back, out_re_im[2i], [2i+1].  What the
 * Kraus-operator sampling of a non-unitary channel needs (lib/qtrajectory.h:344-352: one ExpectationValue per
 * operator, state and qubits unchanged in between).  Arithmetic per operator as in qb200_expectation_value.
 * QB200_ERR_UNSUPPORTED for more than 2 target qubits or more than 8 operators.  Inside a reduce batch: `count` slots. */
int qb200_expectation_values_multi(qb200_ctx* ctx, int dtype, const void* state, unsigned num_qubits,
                                   const unsigned* qs, unsigned num_targets, const void* matrices, unsigned count,
                                   double* out_re_im);
/* Sample (:243-312) minus the host RNG: sorted_rs are the sorted uniform
 * [0,norm) values the caller drew (lib/util.h:67-85); out[m] = first index k
 * whose cumulative probability exceeds sorted_rs[m]; 2^n - 1 when none does
 * (lib/statespace_basic.h:227-229). */
int qb200_sample(qb200_ctx* ctx, int dtype, const void* state, unsigned num_qubits,
                 const double* sorted_rs, uint64_t num_samples, uint64_t* out);
/* Host helper = GenerateRandomValues<double> (lib/util.h:67-85): std::mt19937(seed),
 * uniform_real_distribution(0,max_value), sorted ascending. */
int qb200_generate_random_values(uint64_t num_samples, unsigned seed, double max_value, double* out);
/* The same values drawn and sorted ON THE DEVICE (the reference's TODO at lib/statespace_cuda.h:292): a device
 * MT19937 + libstdc++'s uniform_real_distribution<double> arithmetic + radix sort, bit-identical to the host helper
 * (csrc/sample_rng.cu).  qb200_sample_seeded = Sample(state, num_samples, seed) with `norm` = the caller's Norm(state):
 * no host random numbers, no host sort, no host->device copy; same indices as qb200_sample on the host values.
 * norm < 0: the values are drawn in [0, total of the sampler's own chunk sums) -- Norm(state) up to the summation order,
 * without the separate Norm pass (two reads of the state per call instead of three).
 * qb200_generate_random_values_device copies the sorted values to host memory `out` (tests). */
int qb200_sample_seeded(qb200_ctx* ctx, int dtype, const void* state, unsigned num_qubits, uint64_t num_samples,
                        unsigned seed, double norm, uint64_t* out);
int qb200_generate_random_values_device(qb200_ctx* ctx, uint64_t num_samples, unsigned seed, double max_value, double* out);
/* PartialNorms (:331-355): the state is cut into qb200_partial_norms_count(n)
 * equal contiguous chunks; out[m] = sum |amp|^2 over chunk m. */
uint64_t qb200_partial_norms_count(unsigned num_qubits);
int qb200_partial_norms(qb200_ctx* ctx, int dtype, const void* state, unsigned num_qubits, double* out);
/* FindMeasuredBits (:357-373): inside chunk m, first index whose running sum
 * exceeds r; returns index & mask. */
int qb200_find_measured_bits(qb200_ctx* ctx, int dtype, const void* state, unsigned num_qubits,
                             uint64_t m, double r, uint64_t mask, uint64_t* out_bits);
/* Collapse (:316-329): zero amplitudes with (i & mask) != bits, renormalise
 * the rest by 1/sqrt(masked norm).  out_norm (may be NULL) receives that norm. */
int qb200_collapse(qb200_ctx* ctx, int dtype, void* state, unsigned num_qubits,
                   uint64_t mask, uint64_t bits, double* out_norm);
/* The two halves of Collapse, for callers that combine several shards (csrc/sharded.cu): sum |amp|^2 over
 * (i & mask) == bits, and "zero where (i & mask) != bits, scale the rest by renorm". */
int qb200_masked_norm(qb200_ctx* ctx, int dtype, const void* state, unsigned num_qubits,
                      uint64_t mask, uint64_t bits, double* out);
int qb200_collapse_scaled(qb200_ctx* ctx, int dtype, void* state, unsigned num_qubits,
                          uint64_t mask, uint64_t bits, double renorm);
/* InternalToNormalOrder / NormalToInternalOrder (:85-107): identity here. */
int qb200_internal_to_normal_order(qb200_ctx* ctx, int dtype, void* state, unsigned num_qubits);
int qb200_normal_to_internal_order(qb200_ctx* ctx, int dtype, void* state, unsigned num_qubits);

/* ---- sharded states: local<->global qubit swap over NVLink peer memory ------ */
/* A state of n qubits sharded over 2^g GPUs keeps 2^(n-g) amplitudes per GPU; the rank
 * is the top g index bits.  Exchanging k rank bits with k local bits is the one
 * collective step of the path.  The reference reaches it only through the closed
 * cuStateVecEx library (custatevecExStateVectorPermuteIndexBits in
 * lib/simulator_custatevecex.h:147-196, swaps inside custatevecExSVUpdaterApply,
 * lib/run_custatevecex.h:243-305); here it is ONE kernel per GPU that reads and writes
 * the peers' shards directly over NVLink (CUDA IPC mappings), in place, no staging:
 * each GPU swaps half of every pairwise slice, so both link directions carry
 * shard*(1-2^-k)/... bytes concurrently.
 *
 * qb200_ipc_export / _import / _close: cudaIpcGetMemHandle / OpenMemHandle / CloseMemHandle
 * on a shard allocated by qb200_state_alloc (64-byte opaque handle). */
int qb200_ipc_export(const void* state, unsigned char handle[64]);
int qb200_ipc_import(const unsigned char handle[64], void** peer_state);
int qb200_ipc_close(void* peer_state);
/* local_bits[j] (ascending, < num_local_qubits) is exchanged with the j-th selected rank
 * bit; my_value = this rank's value of those k rank bits; peer_states[b] = mapped shard of
 * the rank whose selected bits equal b (entry my_value ignored).  Every rank of the group
 * must call it between two stream-ordered barriers. */
int qb200_swap_global_local(qb200_ctx* ctx, int dtype, void* state, unsigned num_local_qubits,
                            void* const* peer_states, unsigned k, const unsigned* local_bits,
                            unsigned my_value);


/* ---- sharded states behind the same operations (csrc/sharded.cu) ----------------------------------------
 * qb200_sv = one state of n qubits over 2^g shards + its qubit map + one context/stream per local shard: the
 * counterpart of the multi-device State of lib/vectorspace_custatevecex.h:189-287,385-470 with its wire ordering
 * (:163-177).  Qubit arguments are LOGICAL qubits; the library tracks where each one lives.  Gates may touch any
 * qubits as long as the targets fit one shard (lib/simulator_custatevecex.h:67-71): a target on a global qubit
 * triggers a local<->global exchange (online: least-recently-used victims; qb200_sv_run: planned over the whole
 * gate list with reordering of commuting gates, csrc/sv_plan.h).  The exchange is one kernel per GPU that pushes
 * amplitudes into the peers' memory over NVLink, bracketed by stream-ordered barriers; nothing on this path
 * synchronises the host.  Reductions add per-shard results; Sample / PartialNorms / FindMeasuredBits / host
 * copies first restore the canonical map (pos[q] = q) so they see the same amplitude order as an unsharded state.
 *
 * Ownership modes: qb200_sv_create -- this process drives every shard (devices[r] hosts shard r; a device may be
 * named several times, which is how a one-GPU box exercises the exchange); qb200_sv_create_mp -- one process per
 * shard, `comm` supplies the three host-side collectives (the role of custatevecExCommunicator,
 * lib/multiprocess_custatevecex.h:82-87).  In mp mode every rank must make the same calls in the same order;
 * values returned to the host are identical on all ranks. */
typedef struct qb200_sv qb200_sv;

typedef struct {
  void* user;
  /* recv = concatenation over ranks of each rank's `bytes` bytes at `send` (host memory); 0 on success */
  int (*allgather)(void* user, const void* send, void* recv, uint64_t bytes);
  /* in-place sum over ranks of `count` doubles (host memory); 0 on success */
  int (*allreduce_sum_f64)(void* user, double* inout, uint64_t count);
  int (*barrier)(void* user);
} qb200_comm;

/* One fused gate of a circuit: lib/gate.h Gate / FusedGate / ControlledGate as plain pointers.  qs sorted
 * ascending, matrix as in qb200_apply_gate, bit i of cvals <-> i-th lowest control qubit. */
typedef struct {
  unsigned num_targets;
  const unsigned* qs;
  unsigned num_controls;
  const unsigned* cqs;
  uint64_t cvals;
  const void* matrix;
} qb200_gate;

typedef struct {
  uint64_t swaps;                /* exchanges so far */
  uint64_t local_swap_passes;    /* 2-qubit SWAP passes (in-place exchange of low bits, canonicalisation) */
  uint64_t gate_passes;
  double bytes_sent_per_shard;   /* sum over exchanges of shard_bytes * (1 - 2^-k) */
  double exchange_ms;            /* device time of the exchange kernels (CUDA events on the first local shard) */
  double barrier_wait_ms;        /* device time that shard spent in the barriers around them (skew between GPUs) */
  uint64_t overlapped_swaps;     /* exchanges that ran chunk by chunk beside the last gate passes of their epoch ... */
  uint64_t overlapped_gate_passes;   /* ... how many gate passes those were ... */
  double overlap_ms;             /* ... and the device time of those pipelines (gates + exchange; not in exchange_ms) */
  uint64_t copy_engine_swaps;    /* exchanges moved by the copy engines (pitched 3-D copies) instead of a push kernel */
} qb200_sv_stats;

int qb200_sv_create(const int* devices, unsigned num_shards, unsigned num_qubits, int dtype, qb200_sv** sv);
int qb200_sv_create_mp(int device, unsigned rank, unsigned world, const qb200_comm* comm, unsigned num_qubits,
                       int dtype, qb200_sv** sv);
int qb200_sv_destroy(qb200_sv* sv);
unsigned qb200_sv_num_qubits(const qb200_sv* sv);
unsigned qb200_sv_num_shards(const qb200_sv* sv);
unsigned qb200_sv_num_local_qubits(const qb200_sv* sv);
unsigned qb200_sv_num_local_shards(const qb200_sv* sv);
int qb200_sv_last_cuda_error(const qb200_sv* sv);
/* pos[q] = physical index bit of logical qubit q (bits >= num_local_qubits are the shard number) */
int qb200_sv_qubit_map(const qb200_sv* sv, unsigned* pos);
/* the i-th shard this process owns: its number, device, current device buffer and context (any may be NULL) */
int qb200_sv_shard(const qb200_sv* sv, unsigned local_index, unsigned* rank, int* device, void** state, qb200_ctx** ctx);
/* keys: "swap_mode" (-1 auto: out of place when a second buffer fits, 0 in place, 1 out of place), "push_kernel"
 * (out-of-place exchange: 1 = tiles moved by the bulk-copy engine, default; 0 = 16-byte loads / stores), "reorder"
 * (1: qb200_sv_run may reorder commuting gates, 0: program order), "barrier_flags" (1: flag words in peer memory,
 * 0: CUDA events -- single-process only); any other key is forwarded to qb200_ctx_set_tuning of every shard. */
int qb200_sv_set_option(qb200_sv* sv, const char* key, int value);
int qb200_sv_sync(qb200_sv* sv);
uint64_t qb200_sv_launch_count(const qb200_sv* sv);
int qb200_sv_get_stats(qb200_sv* sv, qb200_sv_stats* out);   /* synchronises */
int qb200_sv_reset_stats(qb200_sv* sv);
/* StateSpace members on a sharded state (lib/statespace_custatevecex.h:79-147, lib/statespace_cuda.h) */
int qb200_sv_set_all_zeros(qb200_sv* sv);
int qb200_sv_set_state_zero(qb200_sv* sv);
int qb200_sv_set_state_uniform(qb200_sv* sv);
int qb200_sv_reset_map(qb200_sv* sv);  /* identity map without moving data: only before (re)initialising the state */
int qb200_sv_get_ampl(qb200_sv* sv, uint64_t i, double out_re_im[2]);
/* several amplitudes with ONE host-side collective (multi-process states): out[2j], out[2j+1] = amplitude indices[j] */
int qb200_sv_get_ampls(qb200_sv* sv, const uint64_t* indices, uint64_t count, double* out_re_im);
int qb200_sv_set_ampl(qb200_sv* sv, uint64_t i, double re, double im);
int qb200_sv_bulk_set_ampl(qb200_sv* sv, uint64_t mask, uint64_t bits, double re, double im, int exclude);
int qb200_sv_norm(qb200_sv* sv, double* out);
int qb200_sv_inner_product(qb200_sv* a, qb200_sv* b, double out_re_im[2]);  /* may re-map both states */
int qb200_sv_add(qb200_sv* src, qb200_sv* dest);
int qb200_sv_copy(qb200_sv* src, qb200_sv* dest);
int qb200_sv_multiply(qb200_sv* sv, double a);
int qb200_sv_sample(qb200_sv* sv, const double* sorted_rs, uint64_t num_samples, uint64_t* out);
uint64_t qb200_sv_partial_norms_count(const qb200_sv* sv);
int qb200_sv_partial_norms(qb200_sv* sv, double* out);
int qb200_sv_find_measured_bits(qb200_sv* sv, uint64_t m, double r, uint64_t mask, uint64_t* out_bits);
int qb200_sv_collapse(qb200_sv* sv, uint64_t mask, uint64_t bits, double* out_norm);
/* whole state, normal order, host memory; a multi-process state moves only this process's shard (at its offset) */
int qb200_sv_copy_to_host(qb200_sv* sv, void* host_dst);
int qb200_sv_copy_from_host(qb200_sv* sv, const void* host_src);
/* Simulator members (lib/simulator_custatevecex.h:60-196) */
int qb200_sv_apply_gate(qb200_sv* sv, const unsigned* qs, unsigned num_targets, const void* matrix);
int qb200_sv_apply_controlled_gate(qb200_sv* sv, const unsigned* qs, unsigned num_targets, const unsigned* cqs,
                                   unsigned num_controls, uint64_t cvals, const void* matrix);
int qb200_sv_expectation_value(qb200_sv* sv, const unsigned* qs, unsigned num_targets, const void* matrix,
                               double out_re_im[2]);
/* A whole fused circuit: plans the exchanges over the list (csrc/sv_plan.h) and applies it -- what
 * custatevecExSVUpdaterEnqueueMatrix + Apply do for the reference (lib/run_custatevecex.h:243-305). */
int qb200_sv_run(qb200_sv* sv, const qb200_gate* gates, uint64_t count);
/* The planner alone (host only, no GPU needed): global_qubits = the num_global qubits that are global at the
 * start (NULL: the top ones).  steps: a value >= 0 is the index of the gate to apply next; -k starts an exchange
 * and is followed by its k victims (local -> global) and k incoming qubits (global -> local).  Call with
 * steps = NULL to size the buffer (*num_steps). */
int qb200_sv_plan(unsigned num_qubits, unsigned num_global, const qb200_gate* gates, uint64_t count,
                  const unsigned* global_qubits, int reorder, int64_t* steps, uint64_t capacity, uint64_t* num_steps);
/* explicit exchange of k local (victims) with k global (incoming) logical qubits; restoring pos[q] = q */
int qb200_sv_swap(qb200_sv* sv, const unsigned* victims, const unsigned* incoming, unsigned k);
int qb200_sv_canonicalize(qb200_sv* sv);
/* The initial global set (num_global qubits, ascending) whose schedule for this gate list exchanges the fewest shards
 * (csrc/sv_plan.h BestInitial): what qb200_sv_run relabels the qubit map to when the state is fresh -- |0...0>, all
 * zeros or uniform since the last qb200_sv_set_* call, states that look the same under every map, so no data moves
 * (option "free_initial_map" = 0 keeps the map as it is).  Host-only. */
int qb200_sv_plan_initial(unsigned num_qubits, unsigned num_global, const qb200_gate* gates, uint64_t count, int reorder,
                          unsigned* global_qubits_out);

#ifdef __cplusplus
}
#endif

#endif  /* QSIM_B200_H_ */
