// context.cu -- context, memory and copy entry points of the C ABI
// (replaces lib/vectorspace_cuda.h and the resource members of
// lib/simulator_cuda.h:52-62 / lib/statespace_cuda.h:378-390).
#include <cstdio>
#include <cstring>
#include <limits>
#include <new>

#include "gate_launch.cuh"

namespace qb200 {

constexpr int kMatSlots = 8;
constexpr size_t kMatSlotBytes = 64 * 1024;  // 6-qubit fp64 matrix

struct MatRing {
  void* h = nullptr;  // pinned, kMatSlots * kMatSlotBytes
  void* d = nullptr;
  cudaEvent_t ev[kMatSlots] = {};
  bool busy[kMatSlots] = {};
  int next = 0;
};

int ensure_scratch(qb200_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->scratch_bytes) return QB200_OK;
  // round up so a growing sequence of requests does not realloc every time
  size_t want = 1 << 16;
  while (want < bytes) want <<= 1;
  if (ctx->scratch) {
    QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    QB_CUDA(ctx, cudaFree(ctx->scratch));
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
  }
  QB_CUDA(ctx, cudaMalloc(&ctx->scratch, want));
  ctx->scratch_bytes = want;
  return QB200_OK;
}

int ensure_pinned(qb200_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->pinned_bytes) return QB200_OK;
  size_t want = 4096;
  while (want < bytes) want <<= 1;
  if (ctx->pinned) {
    QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    QB_CUDA(ctx, cudaFreeHost(ctx->pinned));
    ctx->pinned = nullptr;
    ctx->pinned_bytes = 0;
  }
  QB_CUDA(ctx, cudaHostAlloc(&ctx->pinned, want, cudaHostAllocDefault));
  ctx->pinned_bytes = want;
  return QB200_OK;
}

int ensure_dmat(qb200_ctx* ctx) {
  if (ctx->d_mat) return QB200_OK;
  MatRing* r = new (std::nothrow) MatRing();
  if (!r) return QB200_ERR_OOM;
  cudaError_t e = cudaHostAlloc(&r->h, kMatSlots * kMatSlotBytes, cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaMalloc(&r->d, kMatSlots * kMatSlotBytes);
  for (int i = 0; i < kMatSlots && e == cudaSuccess; ++i)
    e = cudaEventCreateWithFlags(&r->ev[i], cudaEventDisableTiming);
  if (e != cudaSuccess) {
    if (r->h) cudaFreeHost(r->h);
    if (r->d) cudaFree(r->d);
    delete r;
    return cuda_status(ctx, e);
  }
  ctx->d_mat = r;
  return QB200_OK;
}

int check_state_device(qb200_ctx* ctx, const void* p) {
  if (p == ctx->checked_ptr) return QB200_OK;
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    (void) cudaGetLastError();
    return QB200_OK;  // not a CUDA pointer the runtime knows: the launch itself will report it
  }
  if (a.type == cudaMemoryTypeDevice && a.device != ctx->device) return QB200_ERR_INVALID;
  ctx->checked_ptr = p;
  return QB200_OK;
}

// Copies a host matrix into the next ring slot (pinned -> device, async) and
// returns the device address.  The slot is recycled only after the kernel that
// consumed it has finished (event recorded by stage_matrix_done).
int stage_matrix(qb200_ctx* ctx, const void* host, size_t bytes, const void** dev) {
  if (bytes > kMatSlotBytes) return QB200_ERR_INVALID;
  int rc = ensure_dmat(ctx);
  if (rc) return rc;
  MatRing* r = (MatRing*) ctx->d_mat;
  const int s = r->next;
  if (r->busy[s]) {
    QB_CUDA(ctx, cudaEventSynchronize(r->ev[s]));
    r->busy[s] = false;
  }
  char* h = (char*) r->h + s * kMatSlotBytes;
  char* d = (char*) r->d + s * kMatSlotBytes;
  std::memcpy(h, host, bytes);
  QB_CUDA(ctx, cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
  *dev = d;
  return QB200_OK;
}

void stage_matrix_done(qb200_ctx* ctx) {
  MatRing* r = (MatRing*) ctx->d_mat;
  const int s = r->next;
  if (cudaEventRecord(r->ev[s], ctx->stream) == cudaSuccess) r->busy[s] = true;
  r->next = (s + 1) % kMatSlots;
}

__global__ void __launch_bounds__(256)
k_sum_partials2(const double* __restrict__ partials, uint32_t count, double* __restrict__ out) {
  double a = 0, b = 0;
  for (uint32_t i = threadIdx.x; i < count; i += 256) {
    a += partials[2 * i];
    b += partials[2 * i + 1];
  }
  block_sum2<256>(a, b);
  if (threadIdx.x == 0) {
    out[0] = a;
    out[1] = b;
  }
}

// Result slots: mapped pinned host memory the final reduction kernel writes straight into
// (no device->host copy on the stream).
int ensure_results(qb200_ctx* ctx, uint32_t slots) {
  if (slots <= ctx->res_cap) return QB200_OK;
  uint32_t want = 64;
  while (want < slots) want <<= 1;
  double* fresh = nullptr;
  QB_CUDA(ctx, cudaHostAlloc((void**) &fresh, size_t{want} * 2 * sizeof(double),
                             cudaHostAllocMapped | cudaHostAllocPortable));
  if (ctx->res) {
    // results of an open batch stay valid: finish what is in flight, carry the filled slots over
    QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::memcpy(fresh, ctx->res, size_t{ctx->batch_count} * 2 * sizeof(double));
    QB_CUDA(ctx, cudaFreeHost(ctx->res));
  }
  ctx->res = fresh;
  ctx->res_cap = want;
  return QB200_OK;
}

static std::atomic<uint64_t> g_mutation_epoch{1};
void note_state_written() { g_mutation_epoch.fetch_add(1, std::memory_order_relaxed); }

// partials[0 .. 2*blocks) -> out[2] on the host (synchronises the stream); inside a batch
// (qb200_reduce_batch_begin) -> the next result slot, no synchronisation, `out` = NaN.
int finish_expectation(qb200_ctx* ctx, double* partials, uint32_t blocks, double out[2]) {
  const uint32_t slot = ctx->batching ? ctx->batch_count : 0;
  int rc = ensure_results(ctx, slot + 1);
  if (rc) return rc;
  k_sum_partials2<<<1, 256, 0, ctx->stream>>>(partials, blocks, ctx->res + 2 * size_t{slot});
  QB_LAUNCHED(ctx);
  if (ctx->batching) {
    ++ctx->batch_count;
    if (out) out[0] = out[1] = std::numeric_limits<double>::quiet_NaN();
    return QB200_OK;
  }
  QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  out[0] = ctx->res[0];
  out[1] = ctx->res[1];
  return QB200_OK;
}

// `count` results at once: partials[(o * blocks + b) * 2 + {0,1}] -> out[2 * o + {0,1}]; inside a batch the next
// `count` slots.
int finish_expectations(qb200_ctx* ctx, double* partials, uint32_t blocks, uint32_t count, double* out) {
  const uint32_t slot = ctx->batching ? ctx->batch_count : 0;
  int rc = ensure_results(ctx, slot + count);
  if (rc) return rc;
  for (uint32_t o = 0; o < count; ++o) {
    k_sum_partials2<<<1, 256, 0, ctx->stream>>>(partials + size_t{o} * blocks * 2, blocks, ctx->res + 2 * size_t{slot + o});
    QB_LAUNCHED(ctx);
  }
  if (ctx->batching) {
    ctx->batch_count += count;
    if (out)
      for (uint32_t i = 0; i < 2 * count; ++i) out[i] = std::numeric_limits<double>::quiet_NaN();
    return QB200_OK;
  }
  QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (uint32_t i = 0; i < 2 * count; ++i) out[i] = ctx->res[i];
  return QB200_OK;
}

// An unsupported operator inside a batch still takes a slot (value 0, what the reference's ExpectationValue
// returns for it, lib/simulator_cuda.h:259), so that callers can pair results with calls by position.
int batch_zero_slot(qb200_ctx* ctx) {
  if (!ctx->batching) return QB200_OK;
  DeviceGuard guard(ctx);
  int rc = ensure_results(ctx, ctx->batch_count + 1);
  if (rc) return rc;
  ctx->res[2 * size_t{ctx->batch_count}] = 0;      // host write: no kernel owns this slot
  ctx->res[2 * size_t{ctx->batch_count} + 1] = 0;
  ++ctx->batch_count;
  return QB200_OK;
}

}  // namespace qb200

using namespace qb200;

extern "C" {

int qb200_abi_version(void) { return 3; }   // 3: qb200_sv_stats grew (overlap / copy-engine counters), seeded sampler, operator groups

uint64_t qb200_mutation_epoch(void) { return g_mutation_epoch.load(std::memory_order_relaxed); }

int qb200_device_count(int* count) {
  if (!count) return QB200_ERR_INVALID;
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) { *count = 0; (void) cudaGetLastError(); return QB200_ERR_CUDA; }
  return QB200_OK;
}

int qb200_ctx_create(int device, qb200_ctx** out) {
  if (!out) return QB200_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    (void) cudaGetLastError();
    return QB200_ERR_CUDA;  // no CPU fallback: the product path needs a GPU
  }
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) return QB200_ERR_CUDA;
  }
  if (device >= count) return QB200_ERR_INVALID;
  qb200_ctx* ctx = new (std::nothrow) qb200_ctx();
  if (!ctx) return QB200_ERR_OOM;
  ctx->device = device;
  *out = ctx;
  return QB200_OK;
}

int qb200_ctx_destroy(qb200_ctx* ctx) {
  if (!ctx) return QB200_OK;
  {
    DeviceGuard guard(ctx);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->res) cudaFreeHost(ctx->res);
    if (ctx->d_mat) {
      MatRing* r = (MatRing*) ctx->d_mat;
      for (int i = 0; i < kMatSlots; ++i)
        if (r->ev[i]) cudaEventDestroy(r->ev[i]);
      cudaFreeHost(r->h);
      cudaFree(r->d);
      delete r;
    }
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  }
  delete ctx;
  return QB200_OK;
}

int qb200_ctx_set_stream(qb200_ctx* ctx, void* stream) {
  if (!ctx) return QB200_ERR_INVALID;
  ctx->stream = (cudaStream_t) stream;
  return QB200_OK;
}

int qb200_last_cuda_error(const qb200_ctx* ctx) { return ctx ? ctx->last_error : 0; }

const char* qb200_last_cuda_error_string(const qb200_ctx* ctx) {
  return cudaGetErrorString((cudaError_t) (ctx ? ctx->last_error : 0));
}

uint64_t qb200_launch_count(const qb200_ctx* ctx) { return ctx ? ctx->launches : 0; }

int qb200_ctx_set_tuning(qb200_ctx* ctx, const char* key, int value) {
  if (!ctx || !key) return QB200_ERR_INVALID;
  if (!std::strcmp(key, "gate_mode")) ctx->tune.gate_mode = value;
  else if (!std::strcmp(key, "block")) ctx->tune.block = value;
  else if (!std::strcmp(key, "force_generic")) ctx->tune.force_generic = value;
  else if (!std::strcmp(key, "tile")) ctx->tune.tile = value;
  else if (!std::strcmp(key, "prefetch")) ctx->tune.prefetch = value;
  else if (!std::strcmp(key, "big")) ctx->tune.big = value;
  else if (!std::strcmp(key, "expect_ug")) ctx->tune.expect_ug = value;
  else if (!std::strcmp(key, "mono")) ctx->tune.mono = value;
  else if (!std::strcmp(key, "tc")) ctx->tune.tc = value;
  else if (!std::strcmp(key, "tc_low")) ctx->tune.tc_low = value;
  else if (!std::strcmp(key, "tcx")) ctx->tune.tcx = value;
  else if (!std::strcmp(key, "tc_comp6")) ctx->tune.tc_comp6 = value;
  else return QB200_ERR_INVALID;
  return QB200_OK;
}

int qb200_timer_start(qb200_ctx* ctx) {
  if (!ctx) return QB200_ERR_INVALID;
  DeviceGuard guard(ctx);
  if (!ctx->ev0) {
    QB_CUDA(ctx, cudaEventCreate(&ctx->ev0));
    QB_CUDA(ctx, cudaEventCreate(&ctx->ev1));
  }
  QB_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  return QB200_OK;
}

int qb200_timer_stop_ms(qb200_ctx* ctx, float* ms) {
  if (!ctx || !ms || !ctx->ev0) return QB200_ERR_INVALID;
  DeviceGuard guard(ctx);
  QB_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  QB_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
  QB_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  return QB200_OK;
}

uint64_t qb200_min_size(unsigned num_qubits) { return uint64_t{2} << num_qubits; }

static int state_alloc(unsigned num_qubits, int dtype, void** state) {
  if (!state || num_qubits > kMaxQubits || (dtype != QB200_F32 && dtype != QB200_F64))
    return QB200_ERR_INVALID;
  *state = nullptr;
  size_t bytes = qb200_min_size(num_qubits) * (dtype == QB200_F32 ? 4 : 8);
  if (bytes < 256) bytes = 256;
  note_state_written();   // a fresh allocation may reuse the address of a freed state
  cudaError_t e = cudaMalloc(state, bytes);
  if (e != cudaSuccess) {
    (void) cudaGetLastError();
    *state = nullptr;
    return e == cudaErrorMemoryAllocation ? QB200_ERR_OOM : QB200_ERR_CUDA;
  }
  return QB200_OK;
}

int qb200_state_alloc(unsigned num_qubits, int dtype, void** state) { return state_alloc(num_qubits, dtype, state); }

int qb200_state_alloc_on(qb200_ctx* ctx, unsigned num_qubits, int dtype, void** state) {
  if (!ctx) return QB200_ERR_INVALID;
  DeviceGuard guard(ctx);
  return state_alloc(num_qubits, dtype, state);
}

const char* qb200_last_kernel_name(const qb200_ctx* ctx) { return ctx ? ctx->last_kernel : ""; }

int qb200_ctx_set_occupancy_reduction(qb200_ctx* ctx, int ctas_per_sm) {
  if (!ctx || ctas_per_sm < 0) return QB200_ERR_INVALID;
  ctx->occ_reduce = ctas_per_sm;
  return QB200_OK;
}

int qb200_ctx_set_sm_limit(qb200_ctx* ctx, int sms) {
  if (!ctx || sms < 0) return QB200_ERR_INVALID;
  ctx->sm_limit = sms;
  return QB200_OK;
}

int qb200_state_free(void* state) {
  note_state_written();
  if (!state) return QB200_OK;
  return cudaFree(state) == cudaSuccess ? QB200_OK : QB200_ERR_CUDA;
}

static size_t scalar_bytes(int dtype) { return dtype == QB200_F32 ? 4 : 8; }

int qb200_copy_d2d(qb200_ctx* ctx, int dtype, const void* src, void* dst, uint64_t count) {
  note_state_written();
  if (!ctx || !src || !dst) return QB200_ERR_INVALID;
  DeviceGuard guard(ctx);
  QB_CUDA(ctx, cudaMemcpyAsync(dst, src, count * scalar_bytes(dtype), cudaMemcpyDeviceToDevice, ctx->stream));
  QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return QB200_OK;
}

int qb200_copy_d2d_async(qb200_ctx* ctx, int dtype, const void* src, void* dst, uint64_t count) {
  note_state_written();
  if (!ctx || !src || !dst) return QB200_ERR_INVALID;
  DeviceGuard guard(ctx);
  QB_CUDA(ctx, cudaMemcpyAsync(dst, src, count * scalar_bytes(dtype), cudaMemcpyDeviceToDevice, ctx->stream));
  return QB200_OK;
}

int qb200_copy_d2h(qb200_ctx* ctx, int dtype, const void* src, void* host_dst, uint64_t count) {
  if (!ctx || !src || !host_dst) return QB200_ERR_INVALID;
  DeviceGuard guard(ctx);
  QB_CUDA(ctx, cudaMemcpyAsync(host_dst, src, count * scalar_bytes(dtype), cudaMemcpyDeviceToHost, ctx->stream));
  QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return QB200_OK;
}

int qb200_copy_h2d(qb200_ctx* ctx, int dtype, const void* host_src, void* dst, uint64_t count) {
  note_state_written();
  if (!ctx || !host_src || !dst) return QB200_ERR_INVALID;
  DeviceGuard guard(ctx);
  QB_CUDA(ctx, cudaMemcpyAsync(dst, host_src, count * scalar_bytes(dtype), cudaMemcpyHostToDevice, ctx->stream));
  QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return QB200_OK;
}

int qb200_device_sync(void) {
  if (cudaDeviceSynchronize() != cudaSuccess) {
    (void) cudaGetLastError();
    return QB200_ERR_CUDA;
  }
  return QB200_OK;
}

int qb200_device_sync_on(int device) {
  int prev = 0;
  if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) {
    (void) cudaGetLastError();
    return QB200_ERR_CUDA;
  }
  const cudaError_t e = cudaDeviceSynchronize();
  cudaSetDevice(prev);
  if (e != cudaSuccess) {
    (void) cudaGetLastError();
    return QB200_ERR_CUDA;
  }
  return QB200_OK;
}

int qb200_reduce_batch_begin(qb200_ctx* ctx, uint32_t expected) {
  if (!ctx || ctx->batching) return QB200_ERR_INVALID;
  DeviceGuard guard(ctx);
  int rc = ensure_results(ctx, expected ? expected : 1);
  if (rc) return rc;
  ctx->batching = true;
  ctx->batch_count = 0;
  return QB200_OK;
}

int qb200_reduce_batch_end(qb200_ctx* ctx, double* out, uint32_t capacity, uint32_t* count) {
  if (!ctx || !ctx->batching || (capacity && !out)) return QB200_ERR_INVALID;
  DeviceGuard guard(ctx);
  ctx->batching = false;
  const uint32_t n = ctx->batch_count;
  ctx->batch_count = 0;
  if (count) *count = n;
  QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (n > capacity) return QB200_ERR_INVALID;
  std::memcpy(out, ctx->res, size_t{n} * 2 * sizeof(double));
  return QB200_OK;
}

int qb200_sync(qb200_ctx* ctx) {
  if (!ctx) return QB200_ERR_INVALID;
  DeviceGuard guard(ctx);
  QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return QB200_OK;
}

}  // extern "C"
