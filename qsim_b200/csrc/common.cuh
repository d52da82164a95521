// common.cuh -- shared plumbing for libqsim_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstddef>
#include <cstdint>

#include "../../include/qsim_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libqsim_b200 is written for sm_100a (B200) only"
#endif

namespace qb200 {

constexpr int kNumSMs = 148;           // B200
constexpr unsigned kMaxTargets = 6;    // lib/simulator_cuda.h:96-98
constexpr unsigned kMaxCtrlTargets = 4;  // lib/simulator_cuda.h:162-164
constexpr unsigned kMaxQubits = 40;
// Chunk of the state one partial norm covers (PartialNorms / Sample /
// FindMeasuredBits share it): 2^13 amplitudes, the same granularity as the
// reference's default 512 threads x 16 dblocks (lib/statespace_cuda.h:59-70).
constexpr unsigned kChunkBits = 13;

template <typename FP> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

struct Tuning {
  int gate_mode = -1;      // -1 auto; 0 = one amplitude per thread; 1 = two (128-bit)
  int block = 0;           // threads per block override for gate kernels (0 = auto)
  int force_generic = 0;   // 1 = always use the runtime-generic gate kernel
  int tile = -1;           // -1 auto; 0 = register kernels only; 1 = per-thread cp.async ring only;
                           // 2 = warp-cooperative tile kernel whenever legal
  int prefetch = -1;       // -1 auto (on); 0 = no software-pipelined persistent loop
  int tc = -1;             // fp32 G=4,5 on the tensor cores (tcgen05 3xTF32): -1 auto, 0 off, 1..5 variants
  int tcx = -1;            // fp32 G=6 gates and G=4..6 expectation values on the tensor cores: 0 off
  int tc_comp6 = 276;      // accumulation-bias compensation of the G=6 tensor-core gate, in units of 1e-9
  int tc_low = -1;         // -1: per-layout rule (gate_launch.cuh); k >= 0: G=4 on the tensor cores iff lowest non-zero target >= k
  int expect_ug = -1;      // fp32 G<=2 expectation values (k_expect_stream): -1/1 = one group per thread per iteration;
                           // 2 = several groups, grid-strided; 3 = several groups, contiguous per block (both slower)
  int mono = -1;           // -1 auto: G >= 3 expectation values of XOR-monomial matrices (Pauli strings) as a read pass
                           // without a mat-vec (expect_monomial.cu); 0 = always the dense kernels
  int big = -1;            // -1 auto (on); 0 = fp32 G>=5 through the register/generic kernels;
                           // 1/2/3 = alternative launch shapes of k_gate_big (tools/microbench.py)
};

}  // namespace qb200

struct qb200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  void* scratch = nullptr;        // device scratch (reduction partials, chunk sums, samples)
  size_t scratch_bytes = 0;
  void* pinned = nullptr;         // pinned host result slot
  size_t pinned_bytes = 0;
  void* d_mat = nullptr;          // device copy of matrices too big for kernel parameters
  double* res = nullptr;          // mapped pinned result slots (re, im): the final reduction kernel writes here
  uint32_t res_cap = 0;           // slots allocated
  bool batching = false;          // reductions are enqueued without synchronising (qb200_reduce_batch_begin/end)
  uint32_t batch_count = 0;       // slots filled by the open batch
  int last_error = 0;             // cudaError_t
  uint64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int sm_limit = 0;               // > 0: persistent grids are sized for this many SMs (a concurrent exchange kernel owns the rest)
  int occ_reduce = 0;             // persistent grids assume this many fewer resident CTAs per SM (an exchange kernel
                                  // running beside the gates holds registers / threads / shared memory on every SM)
  const void* checked_ptr = nullptr;  // last state pointer verified to live on `device` (check_state_device)
  const char* last_kernel = "";   // name of the gate kernel the dispatcher picked last (qb200_last_kernel_name)
  qb200::Tuning tune;
};

namespace qb200 {

// Records a failed runtime call in the context and turns it into a status.
inline int cuda_status(qb200_ctx* ctx, cudaError_t e) {
  if (e == cudaSuccess) return QB200_OK;
  if (ctx) ctx->last_error = (int) e;
  (void) cudaGetLastError();  // clear the sticky flag of non-fatal errors
  return e == cudaErrorMemoryAllocation ? QB200_ERR_OOM : QB200_ERR_CUDA;
}

#define QB_CUDA(ctx, call)                                   \
  do {                                                       \
    cudaError_t e__ = (call);                                \
    if (e__ != cudaSuccess) return ::qb200::cuda_status((ctx), e__); \
  } while (0)

#define QB_LAUNCHED(ctx)                                      \
  do {                                                        \
    ++(ctx)->launches;                                        \
    cudaError_t e__ = cudaPeekAtLastError();                  \
    if (e__ != cudaSuccess) return ::qb200::cuda_status((ctx), e__); \
  } while (0)

// QB200_ERR_INVALID when `p` is plain device memory of ANOTHER GPU than the context's (kernels of this context would
// reach it over the interconnect, or fault); one runtime query per new pointer, remembered in the context.
int check_state_device(qb200_ctx* ctx, const void* p);

// Process-wide count of operations that may have WRITTEN a state (gate passes, element-wise maps, copies into device
// memory, exchanges, frees): what a cache of results derived from a state is validated against
// (qb200_mutation_epoch, include/qsim_b200/simulator_b200.h operator groups).
void note_state_written();

// Lazily grown device scratch / pinned host slot.
int ensure_scratch(qb200_ctx* ctx, size_t bytes);
int ensure_pinned(qb200_ctx* ctx, size_t bytes);
int ensure_dmat(qb200_ctx* ctx);

// resident CTAs per SM a persistent grid may count on
inline int grid_occ(const qb200_ctx* ctx, int occ) { return occ - ctx->occ_reduce >= 1 ? occ - ctx->occ_reduce : 1; }

// SMs a persistent grid may assume it owns.
inline int grid_sms(const qb200_ctx* ctx) { return ctx->sm_limit > 0 && ctx->sm_limit < kNumSMs ? ctx->sm_limit : kNumSMs; }

// Function attributes (dynamic shared memory opt-in) and occupancy answers are PER DEVICE: a launcher keeps one
// of these as a function-local static and fills the slot of ctx->device on first use there (the caller holds a
// DeviceGuard, so the current device is ctx->device).  Slots are idempotent, a race only repeats the query.
constexpr int kMaxDevices = 64;
struct PerDevice {
  std::atomic<int> v[kMaxDevices];
  PerDevice() { for (auto& x : v) x.store(0, std::memory_order_relaxed); }
  template <typename F> int get(const qb200_ctx* ctx, F&& init) {
    const int d = ctx->device >= 0 && ctx->device < kMaxDevices ? ctx->device : 0;
    int r = v[d].load(std::memory_order_acquire);
    if (r == 0) { r = init(); v[d].store(r, std::memory_order_release); }
    return r;
  }
};

struct DeviceGuard {
  explicit DeviceGuard(const qb200_ctx* ctx) {
    cudaGetDevice(&prev_);
    if (prev_ != ctx->device) { cudaSetDevice(ctx->device); switched_ = true; }
  }
  ~DeviceGuard() { if (switched_) cudaSetDevice(prev_); }
  int prev_ = 0;
  bool switched_ = false;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of up to two doubles; result valid in thread 0.  Fixed
// summation tree -> run-to-run deterministic.
template <int NT>
__device__ __forceinline__ void block_sum2(double& a, double& b) {
  __shared__ double sa[NT / 32], sb[NT / 32];
  a = warp_sum(a);
  b = warp_sum(b);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sa[w] = a; sb[w] = b; }
  __syncthreads();
  if (w == 0) {
    a = lane < NT / 32 ? sa[lane] : 0.0;
    b = lane < NT / 32 ? sb[lane] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
  }
  __syncthreads();
}

}  // namespace qb200
