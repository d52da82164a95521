// sharded.cu -- a state of n qubits sharded over 2^g GPUs by global qubits, behind the C ABI (qb200_sv_*).
//
// What the reference only reaches through the closed cuStateVecEx library -- the multi-device State of
// lib/vectorspace_custatevecex.h:189-287,385-470, its wire ordering (:163-177), the index-bit swaps of
// lib/simulator_custatevecex.h:147-196 and the scheduler behind custatevecExSVUpdaterApply
// (lib/run_custatevecex.h:243-305) -- built here from the single-GPU kernels of this library plus
//   * a qubit map (logical qubit -> physical index bit; the top g physical bits are the shard number),
//   * ONE exchange kernel per GPU that pushes a shard's amplitudes straight into the peers' memory over NVLink
//     while re-packing the local index bits (k_remap_push; in place through k_p2p_swap when a second buffer does
//     not fit), bracketed by stream-ordered barriers (events inside one process, flag words in peer memory
//     between processes) -- no host synchronisation, no NCCL, no PyTorch,
//   * the look-ahead, reordering swap planner of sv_plan.h.
// Two ways to own the shards: one process drives all devices (qb200_sv_create; a device may hold several shards,
// which is how a single-GPU box tests this file), or one process per GPU (qb200_sv_create_mp; the caller supplies
// allgather / allreduce / barrier callbacks, the analogue of custatevecExCommunicator,
// lib/multiprocess_custatevecex.h:82-87).
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <vector>

#include "common.cuh"
#include "sv_plan.h"

namespace qb200 {

// gates_f{32,64}_apply.cu: the dispatcher behind qb200_apply_controlled_gate (no limit on targets under control)
int gate_apply_f32(qb200_ctx* ctx, float* st, unsigned n, const unsigned* qs, unsigned nq, const unsigned* cqs,
                   unsigned nc, uint64_t cvals, const float* m, double* out);
int gate_apply_f64(qb200_ctx* ctx, double* st, unsigned n, const unsigned* qs, unsigned nq, const unsigned* cqs,
                   unsigned nc, uint64_t cvals, const double* m, double* out);

constexpr unsigned kMaxGlobal = 5;        // up to 32 shards
constexpr unsigned kMaxShards = 1u << kMaxGlobal;
constexpr unsigned kMinInplaceBit = 4;    // in-place exchange: victims below this bit are lifted first (short runs)

// ---------------------------------------------------------------------------------------------------------
// exchange kernel, out of place: every amplitude of the local shard goes to its place in the NEW layout, which
// is a buffer of this or of a peer GPU.  The k victim bits of the local index select the destination shard; the
// remaining local bits are packed (order kept) into the low n_local - k bits of the new local index and this
// shard's own value of the k exchanged rank bits becomes the top k bits.  Push only: nothing is read over NVLink.
//
// A CTA moves tiles of 2^T consecutive source amplitudes (16 KB) through shared memory: coalesced 16-byte loads,
// a scatter inside shared memory that sorts the tile by the victim bits it contains, then coalesced 16-byte
// stores -- every destination receives runs of >= 2 KB whatever the victim bits are (victims on bits 0..2 would
// otherwise reach the link as 8..64-byte pieces: 200-400 GB/s instead of 690), so no local SWAP pass is needed in
// front of the exchange.  Victim bits above the tile pick the destination per tile; tiles are walked with those
// bits fastest and XOR-ed with this shard's own value, so at any moment the shards of a group send to DIFFERENT
// destinations (walking a shard in address order makes all of them hit the same destination at once: 290 GB/s).
// ---------------------------------------------------------------------------------------------------------
constexpr unsigned kTileBits = 11;  // 2^11 amplitudes per tile
constexpr int kRemapThreads = 256;

struct RemapGeom {
  void* dst[kMaxShards];   // new buffer of the shard whose exchanged rank bits equal v
  uint64_t tiles;          // 2^(nl - T)
  uint64_t walk;           // tile counters this launch walks: tiles >> nchunk
  uint32_t k, kl, my, nl, T;
  // a launch may cover one CHUNK of the shard only (exchange overlapped with gates, run_overlapped): nchunk bits of
  // the tile counter are pinned to cval, at positions cpos[] (ascending, in the counter with those bits present)
  uint32_t nchunk, cval, cpos[3];
  uint32_t lbits[kMaxGlobal];  // ascending; the first kl are below T
};

// dense counter inside a chunk -> tile counter (identity when the launch covers the whole shard)
__device__ __forceinline__ uint64_t chunk_counter(uint64_t cc, const RemapGeom& g) {
  for (uint32_t j = 0; j < g.nchunk; ++j) {
    const uint64_t lo = cc & ((uint64_t{1} << g.cpos[j]) - 1);
    cc = (((cc >> g.cpos[j]) << 1 | ((g.cval >> j) & 1u)) << g.cpos[j]) | lo;
  }
  return cc;
}

// FP = float: 16-byte items hold two amplitudes; FP = double: one.
template <typename FP>
__global__ void __launch_bounds__(kRemapThreads)
k_remap_push(const FP* __restrict__ src, const __grid_constant__ RemapGeom g) {
  using A = typename Vec2<FP>::type;          // one amplitude
  extern __shared__ __align__(16) unsigned char smem_raw[];
  A* tile = reinterpret_cast<A*>(smem_raw);
  constexpr int APT = sizeof(FP) == 4 ? 2 : 1;               // amplitudes per 16-byte item
  const uint32_t tile_amps = 1u << g.T;
  const uint32_t items = tile_amps / APT;                    // 16-byte items per tile
  const uint32_t kh = g.k - g.kl;
  const uint32_t sub_bits = g.T - g.kl;                      // log2(amplitudes per destination run)
  const uint32_t my_high = g.my >> g.kl;
  for (uint64_t cc = blockIdx.x; cc < g.walk; cc += gridDim.x) {
    const uint64_t c = chunk_counter(cc, g);
    // tile counter -> tile index: the kh victim bits above the tile vary fastest, XOR-ed with this shard's value
    const uint32_t v_high = ((uint32_t) c & ((1u << kh) - 1)) ^ my_high;
    uint64_t tau = c >> kh;                                  // bits of the tile index outside the victims
    for (uint32_t j = g.kl; j < g.k; ++j) {                  // insert the victim bits, lowest first
      const uint32_t b = g.lbits[j] - g.T;
      const uint64_t lo = tau & ((uint64_t{1} << b) - 1);
      tau = (((tau >> b) << 1 | ((v_high >> (j - g.kl)) & 1)) << b) | lo;
    }
    const uint64_t packed_high = c >> kh;                    // the same bits with the victims squeezed out
    // ---- load (coalesced) and scatter into shared memory, sorted by the low victim bits
    const uint4* in = reinterpret_cast<const uint4*>(src) + tau * items;
    __syncthreads();                                         // the previous tile has been read out
    for (uint32_t it = threadIdx.x; it < items; it += kRemapThreads) {
      const uint4 x = in[it];
#pragma unroll
      for (int a = 0; a < APT; ++a) {
        uint32_t i = it * APT + a, v = 0, r = i;
        for (int j = (int) g.kl - 1; j >= 0; --j) {
          const uint32_t b = g.lbits[j];
          v |= ((r >> b) & 1u) << j;
          r = ((r >> (b + 1)) << b) | (r & ((1u << b) - 1));
        }
        const uint32_t p = (v << sub_bits) | r;
        if constexpr (APT == 2) {
          reinterpret_cast<uint2*>(tile)[p] = a == 0 ? make_uint2(x.x, x.y) : make_uint2(x.z, x.w);
        } else {
          reinterpret_cast<uint4*>(tile)[p] = x;
        }
      }
    }
    __syncthreads();
    // ---- read out linearly: run v_low of the tile goes, contiguously, to destination v_low | v_high << kl
    for (uint32_t it = threadIdx.x; it < items; it += kRemapThreads) {
      const uint32_t p = it * APT;
      const uint32_t v = (p >> sub_bits) | (v_high << g.kl);
      const uint64_t d = (uint64_t{g.my} << (g.nl - g.k)) | (packed_high << sub_bits) | (p & ((1u << sub_bits) - 1));
      reinterpret_cast<uint4*>(g.dst[v])[d / APT] = reinterpret_cast<const uint4*>(tile)[it];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// The same exchange with the bulk-copy engine (TMA) on both ends of the tile: one thread issues
// cp.async.bulk.shared::cluster.global for the next 16 KB tile (mbarrier complete_tx), the CTA's threads only do the
// shared->shared scatter that sorts the tile by its victim bits, and one thread issues one
// cp.async.bulk.global.shared::cta per destination run (>= 2 KB, straight into the peer's buffer).  Two input and two
// output stages per CTA: 64 KB in flight per CTA without holding registers or warps, so the kernel needs few threads --
// which is what lets gate kernels run beside it.
// ---------------------------------------------------------------------------------------------------------
namespace bulk {
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(mbar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void store(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
}  // namespace bulk

// NT threads: 256 when the exchange runs alone; 64 when gate kernels run beside it (the threads only do the
// shared->shared scatter; 64 x 32 registers fit next to three resident k_gate_tca<4> CTAs of 256 x 80).
// `sin` input stages (a ring, loads issued sin - 1 or sin - 2 tiles ahead): beside gate kernels that keep HBM
// saturated a bulk load takes several microseconds, so the slim variant runs with a deep ring.  Without victim bits
// inside the tile (kl = 0) a tile needs no sorting: it is stored straight from its input stage, no output stages.
constexpr int kMaxPushStages = 12;

template <typename FP, int NT>
__global__ void __launch_bounds__(NT)
k_remap_push_tma(const FP* __restrict__ src, const __grid_constant__ RemapGeom g, const uint32_t sin) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[kMaxPushStages];
  constexpr int APT = sizeof(FP) == 4 ? 2 : 1;
  const uint32_t tile_amps = 1u << g.T;
  const uint32_t tile_bytes = tile_amps * 2 * sizeof(FP);
  const uint32_t items = tile_amps / APT;
  const uint32_t kh = g.k - g.kl;
  const uint32_t sub_bits = g.T - g.kl;
  const uint32_t my_high = g.my >> g.kl;
  const bool sort = g.kl != 0;
  const uint32_t ahead = sort ? sin - 1 : sin - 2;             // tiles a load is issued ahead of its use
  unsigned char* const in_p = smem_raw;                        // sin input stages
  unsigned char* const out_p = smem_raw + sin * tile_bytes;    // 2 output stages (sort only)
  const uint32_t in_s = bulk::smem_addr(in_p), out_s = bulk::smem_addr(out_p);
  const uint32_t full_s = bulk::smem_addr(&full[0]);
  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < sin; ++i) bulk::mbar_init(full_s + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto tile_index = [&](uint64_t c, uint32_t* v_high) {
    *v_high = ((uint32_t) c & ((1u << kh) - 1)) ^ my_high;
    uint64_t tau = c >> kh;
    for (uint32_t j = g.kl; j < g.k; ++j) {
      const uint32_t b = g.lbits[j] - g.T;
      const uint64_t lo = tau & ((uint64_t{1} << b) - 1);
      tau = (((tau >> b) << 1 | ((*v_high >> (j - g.kl)) & 1)) << b) | lo;
    }
    return tau;
  };
  auto issue_load = [&](uint64_t cc, uint32_t stage) {
    uint32_t vh;
    const uint64_t tau = tile_index(chunk_counter(cc, g), &vh);
    bulk::mbar_expect_tx(full_s + 8 * stage, tile_bytes);
    bulk::load(in_s + stage * tile_bytes, reinterpret_cast<const unsigned char*>(src) + tau * tile_bytes, tile_bytes,
               full_s + 8 * stage);
  };

  uint64_t cc = blockIdx.x;
  if (threadIdx.x == 0) {
    uint64_t cp = cc;
    for (uint32_t i = 0; i < ahead && cp < g.walk; ++i, cp += gridDim.x) issue_load(cp, i);
  }
  uint32_t st_in = 0, ph_in = 0;                              // input stage of this iteration and its mbarrier phase
  uint32_t st_ld = ahead % sin;                               // stage the look-ahead load of this iteration goes to
  for (uint32_t it = 0; cc < g.walk; cc += gridDim.x, ++it) {
    const uint32_t s = it & 1;
    const uint64_t c = chunk_counter(cc, g);
    if (threadIdx.x == 0) {
      // sort: the stores issued two iterations ago have read output stage s; the look-ahead stage was emptied by the
      // scatter of the previous iteration.  No sort: the look-ahead stage was the source of the stores of two
      // iterations ago.
      bulk::wait_read<1>();
      const uint64_t cn = cc + uint64_t{ahead} * gridDim.x;
      if (cn < g.walk) issue_load(cn, st_ld);
    }
    if (sort) __syncthreads();
    if (sort || threadIdx.x == 0) bulk::mbar_wait(full_s + 8 * st_in, ph_in);
    uint32_t v_high;
    (void) tile_index(c, &v_high);
    const uint64_t packed_high = c >> kh;
    const uint64_t d = ((uint64_t{g.my} << (g.nl - g.k)) | (packed_high << sub_bits)) * 2 * sizeof(FP);
    if (sort) {
      const unsigned char* const tin = in_p + st_in * tile_bytes;
      unsigned char* const tout = out_p + s * tile_bytes;
      for (uint32_t i0 = threadIdx.x; i0 < items; i0 += NT) {
        const uint4 x = reinterpret_cast<const uint4*>(tin)[i0];
#pragma unroll
        for (int a = 0; a < APT; ++a) {
          uint32_t i = i0 * APT + a, v = 0, r = i;
          for (int j = (int) g.kl - 1; j >= 0; --j) {
            const uint32_t b = g.lbits[j];
            v |= ((r >> b) & 1u) << j;
            r = ((r >> (b + 1)) << b) | (r & ((1u << b) - 1));
          }
          const uint32_t p = (v << sub_bits) | r;
          if constexpr (APT == 2) {
            reinterpret_cast<uint2*>(tout)[p] = a == 0 ? make_uint2(x.x, x.y) : make_uint2(x.z, x.w);
          } else {
            reinterpret_cast<uint4*>(tout)[p] = x;
          }
        }
      }
      bulk::fence_async();   // generic-proxy writes to shared memory -> visible to the bulk-copy engine
      __syncthreads();
      if (threadIdx.x == 0) {
        const uint32_t run_bytes = (1u << sub_bits) * 2 * sizeof(FP);
        for (uint32_t vl = 0; vl < (1u << g.kl); ++vl) {
          const uint32_t v = vl | (v_high << g.kl);
          bulk::store(reinterpret_cast<unsigned char*>(g.dst[v]) + d, out_s + s * tile_bytes + vl * run_bytes, run_bytes);
        }
        bulk::commit();
      }
    } else if (threadIdx.x == 0) {
      bulk::store(reinterpret_cast<unsigned char*>(g.dst[v_high]) + d, in_s + st_in * tile_bytes, tile_bytes);
      bulk::commit();
    }
    if (++st_in == sin) { st_in = 0; ph_in ^= 1; }
    if (++st_ld == sin) st_ld = 0;
  }
  if (threadIdx.x == 0) bulk::wait_all();
}

// ---------------------------------------------------------------------------------------------------------
// flag barrier between shards that live in different processes: shard `me` writes the epoch into its slot of
// every peer's flag array (release, system scope) and waits until all its own slots carry it.  One warp.
// A peer that never arrives trips the timeout instead of hanging the GPU; the error word is mapped host memory.
// ---------------------------------------------------------------------------------------------------------
struct FlagGeom {
  uint32_t* flags[kMaxShards];  // flag array of shard r (P words), addressable from this device
  uint32_t P, me, epoch;
  uint32_t* err;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(32) k_flag_barrier(const __grid_constant__ FlagGeom g) {
  __threadfence_system();
  for (uint32_t s = threadIdx.x; s < g.P; s += 32) {
    if (s == g.me) continue;
    st_release_sys(g.flags[s] + g.me, g.epoch);
  }
  const uint64_t t0 = global_ns();
  for (uint32_t s = threadIdx.x; s < g.P; s += 32) {
    if (s == g.me) continue;
    while ((int32_t) (ld_acquire_sys(g.flags[g.me] + s) - g.epoch) < 0) {
      __nanosleep(200);
      if (global_ns() - t0 > 20000000000ull) {  // 20 s: a peer died
        *g.err = 1;
        break;
      }
    }
  }
  __threadfence_system();
}

struct Shard {
  int device = 0;
  unsigned rank = 0;
  qb200_ctx* ctx = nullptr;
  cudaStream_t stream = nullptr;
  void* buf[2] = {nullptr, nullptr};
  uint32_t* flags = nullptr;
  cudaEvent_t ev = nullptr;
  cudaStream_t stream2 = nullptr;   // gates that run beside an exchange (overlap)
  cudaEvent_t ev_chunk = nullptr, ev_join = nullptr;
};

struct SvStats {
  uint64_t swaps = 0, local_passes = 0, gate_passes = 0;
  double bytes_sent_per_shard = 0;   // summed over swaps
  double exchange_ms = 0;            // device time of the exchange kernels (events on the first local shard's stream)
  double wait_ms = 0;                // device time the first local shard spent in the barriers around them (rank skew)
  uint64_t overlapped_swaps = 0;     // exchanges that ran chunk by chunk beside the last gates of their epoch
  uint64_t overlapped_gate_passes = 0;   // ... and how many gate passes those were
  double overlap_ms = 0;             // device time of those pipelines (gates + exchange), not part of exchange_ms
  uint64_t ce_swaps = 0;             // exchanges moved by the copy engines (ce_push)
};

}  // namespace qb200

using namespace qb200;

struct qb200_sv {
  unsigned n = 0, g = 0, nl = 0, P = 1;
  int dtype = QB200_F32;
  bool mp = false;
  unsigned rank = 0;                     // mp: this process's shard
  qb200_comm comm{};
  std::vector<Shard> sh;                 // local shards
  int cur = 0;                           // buf[cur] holds the state on every shard
  bool have_alt = false, alt_failed = false;
  std::vector<void*> peer_buf[2];        // [b][rank] addressable from the local devices (own entries = local pointers)
  std::vector<uint32_t*> peer_flags;
  std::vector<void*> ipc_opened;         // mappings to close
  std::vector<unsigned> pos;             // logical qubit -> physical bit
  std::vector<uint64_t> last_use;        // LRU clock for the online path
  uint64_t clock = 0;
  uint32_t epoch = 0;
  uint32_t* err_host = nullptr;          // mapped pinned
  uint32_t* err_dev = nullptr;
  int barrier_flags = 0;                 // 1: flag kernels (always in mp mode); 0: events
  int swap_mode = -1;                    // -1 auto (out of place when the second buffer fits), 0 in place, 1 out of place
  int reorder = 1;
  int push_kernel = 1;                   // 1: bulk-copy engine (k_remap_push_tma, default); 0: st.global from registers;
                                         // 2: the copy engines (pitched cudaMemcpy3DAsync) when the victims sit at bit 12
                                         //    or above, else 1
  int push_ctas_per_sm = 0;              // 0 = default of the chosen kernel
  int overlap_ctas_per_sm = 0;           // CTAs per SM of the slim push kernel that runs beside gates (0 = 1)
  int overlap_smem_kb = 0;               // ... and its shared memory per CTA (0 = 16 KB)
  int overlap_tile_bits = 9;             // ... which moves tiles of 2^this amplitudes (4 KB in fp32)
  bool fresh = false;                    // the state is |0...0> / all zeros / uniform and nothing has touched it since: it looks
                                         // the same under every qubit map, so qb200_sv_run may pick the map for the circuit
  int free_initial_map = 1;              // ... and does, unless this is 0
  std::vector<uint64_t> init_key;        // gate structure the cached initial global set was chosen for
  uint64_t init_glob = 0;
  int overlap = 0;                       // qb200_sv_run: the last gates of an epoch run beside its exchange, chunk by chunk
                                         // (off by default: measured no gain on B200, profiles/r02_overlap_trace.txt)
  int overlap_chunks_log2 = 2;           // 2^this chunks per shard
  int overlap_max_gates = 6;             // gate passes pipelined against one exchange (enough to cover it, see run_overlapped)
  int overlap_ce = 1;                    // overlapped exchanges go through the copy engines when the layout allows
  int overlap_trace = 0;                 // 1: print a device timeline of the next overlapped exchange (first local shard) to stderr
  int overlap_occ_reduce = 0;            // resident gate CTAs per SM given up while the slim push kernel runs beside them
  SvStats stats;
  std::vector<uint64_t> plan_key;        // gate structure + global set the cached schedule was made for
  std::vector<int64_t> plan_steps;
  std::vector<std::array<cudaEvent_t, 4>> ev_quads;  // exchange timing, first local shard (timing_mark)
  std::vector<char> ev_overlapped;       // quad i timed an overlapped exchange (its [1]..[2] stretch includes gates)
  size_t ev_used = 0;
  int last_error = 0;
};

namespace qb200 {

static size_t scalar_size(int dtype) { return dtype == QB200_F32 ? 4 : 8; }
static size_t shard_bytes(const qb200_sv* sv) { return (size_t{2} << sv->nl) * scalar_size(sv->dtype); }

#define SV_CUDA(sv, call)                                 \
  do {                                                    \
    cudaError_t e__ = (call);                             \
    if (e__ != cudaSuccess) {                             \
      (sv)->last_error = (int) e__;                       \
      (void) cudaGetLastError();                          \
      return e__ == cudaErrorMemoryAllocation ? QB200_ERR_OOM : QB200_ERR_CUDA; \
    }                                                     \
  } while (0)

#define SV_TRY(expr)           \
  do {                         \
    int rc__ = (expr);         \
    if (rc__ != QB200_OK) return rc__; \
  } while (0)

struct DevScope {
  explicit DevScope(int d) { cudaGetDevice(&prev); if (prev != d) cudaSetDevice(d); }
  ~DevScope() { int now; cudaGetDevice(&now); if (now != prev) cudaSetDevice(prev); }
  int prev = 0;
};

static inline void* cur_buf(const qb200_sv* sv, const Shard& s) { return s.buf[sv->cur]; }

static unsigned qubit_at(const qb200_sv* sv, unsigned phys) {
  for (unsigned q = 0; q < sv->n; ++q)
    if (sv->pos[q] == phys) return q;
  return sv->n;
}

static bool canonical(const qb200_sv* sv) {
  for (unsigned q = 0; q < sv->n; ++q)
    if (sv->pos[q] != q) return false;
  return true;
}

// logical amplitude index -> physical index
static uint64_t to_physical(const qb200_sv* sv, uint64_t i) {
  uint64_t p = 0;
  for (unsigned q = 0; q < sv->n; ++q) p |= ((i >> q) & 1) << sv->pos[q];
  return p;
}
static uint64_t to_logical(const qb200_sv* sv, uint64_t p) {
  uint64_t i = 0;
  for (unsigned q = 0; q < sv->n; ++q) i |= ((p >> sv->pos[q]) & 1) << q;
  return i;
}

static Shard* local_shard(qb200_sv* sv, unsigned rank) {
  for (auto& s : sv->sh)
    if (s.rank == rank) return &s;
  return nullptr;
}

static int sum_over_ranks(qb200_sv* sv, double* v, uint64_t count) {
  if (!sv->mp) return QB200_OK;
  return sv->comm.allreduce_sum_f64(sv->comm.user, v, count) == 0 ? QB200_OK : QB200_ERR_INVALID;
}

// ---- barriers -----------------------------------------------------------------------------------------
static int barrier(qb200_sv* sv) {
  if (sv->P == 1) return QB200_OK;
  if (!sv->barrier_flags) {
    for (auto& s : sv->sh) {
      DevScope d(s.device);
      SV_CUDA(sv, cudaEventRecord(s.ev, s.stream));
    }
    for (auto& s : sv->sh) {
      DevScope d(s.device);
      for (auto& t : sv->sh)
        if (&t != &s) SV_CUDA(sv, cudaStreamWaitEvent(s.stream, t.ev, 0));
    }
    return QB200_OK;
  }
  ++sv->epoch;
  for (auto& s : sv->sh) {
    DevScope d(s.device);
    FlagGeom fg{};
    for (unsigned r = 0; r < sv->P; ++r) fg.flags[r] = sv->peer_flags[r];
    fg.P = sv->P;
    fg.me = s.rank;
    fg.epoch = sv->epoch;
    fg.err = sv->err_dev;
    k_flag_barrier<<<1, 32, 0, s.stream>>>(fg);
    ++s.ctx->launches;
    SV_CUDA(sv, cudaPeekAtLastError());
  }
  return QB200_OK;
}

static int sync_all(qb200_sv* sv) {
  for (auto& s : sv->sh) {
    DevScope d(s.device);
    SV_CUDA(sv, cudaStreamSynchronize(s.stream));
  }
  if (sv->err_host && *sv->err_host) return QB200_ERR_CUDA;  // a flag barrier timed out
  return QB200_OK;
}

// ---- creation -------------------------------------------------------------------------------------------
static int make_shard(qb200_sv* sv, int device, unsigned rank) {
  Shard s;
  s.device = device;
  s.rank = rank;
  SV_TRY(qb200_ctx_create(device, &s.ctx));
  DevScope d(device);
  SV_CUDA(sv, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
  qb200_ctx_set_stream(s.ctx, s.stream);
  SV_CUDA(sv, cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
  SV_CUDA(sv, cudaStreamCreateWithFlags(&s.stream2, cudaStreamNonBlocking));
  SV_CUDA(sv, cudaEventCreateWithFlags(&s.ev_chunk, cudaEventDisableTiming));
  SV_CUDA(sv, cudaEventCreateWithFlags(&s.ev_join, cudaEventDisableTiming));
  int rc = qb200_state_alloc_on(s.ctx, sv->nl, sv->dtype, &s.buf[0]);
  if (rc) { sv->sh.push_back(s); return rc; }
  SV_CUDA(sv, cudaMalloc((void**) &s.flags, 256 * sizeof(uint32_t)));
  SV_CUDA(sv, cudaMemset(s.flags, 0, 256 * sizeof(uint32_t)));
  sv->sh.push_back(s);
  return QB200_OK;
}

static int init_common(qb200_sv* sv, unsigned num_shards, unsigned num_qubits, int dtype) {
  unsigned g = 0;
  while ((1u << g) < num_shards) ++g;
  if ((1u << g) != num_shards || g > kMaxGlobal) return QB200_ERR_INVALID;
  if (dtype != QB200_F32 && dtype != QB200_F64) return QB200_ERR_INVALID;
  // same guard as lib/multiprocess_custatevecex.h:160-163: at least two local qubits
  if ((g > 0 && num_qubits < g + 2) || num_qubits > kMaxQubits) return QB200_ERR_INVALID;
  sv->n = num_qubits;
  sv->g = g;
  sv->nl = num_qubits - g;
  sv->P = num_shards;
  sv->dtype = dtype;
  sv->pos.resize(num_qubits);
  for (unsigned q = 0; q < num_qubits; ++q) sv->pos[q] = q;
  sv->last_use.assign(num_qubits, 0);
  sv->peer_buf[0].assign(num_shards, nullptr);
  sv->peer_buf[1].assign(num_shards, nullptr);
  sv->peer_flags.assign(num_shards, nullptr);
  if (cudaHostAlloc((void**) &sv->err_host, 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
    (void) cudaGetLastError();
    return QB200_ERR_CUDA;
  }
  *sv->err_host = 0;
  if (cudaHostGetDevicePointer((void**) &sv->err_dev, sv->err_host, 0) != cudaSuccess) {
    (void) cudaGetLastError();
    return QB200_ERR_CUDA;
  }
  return QB200_OK;
}

static int allgather_ipc(qb200_sv* sv, void* mine, std::vector<void*>* table) {
  std::vector<unsigned char> send(64), recv(size_t{64} * sv->P);
  SV_TRY(qb200_ipc_export(mine, send.data()));
  if (sv->comm.allgather(sv->comm.user, send.data(), recv.data(), 64) != 0) return QB200_ERR_INVALID;
  for (unsigned r = 0; r < sv->P; ++r) {
    if (r == sv->rank) { (*table)[r] = mine; continue; }
    void* p = nullptr;
    SV_TRY(qb200_ipc_import(recv.data() + size_t{64} * r, &p));
    sv->ipc_opened.push_back(p);
    (*table)[r] = p;
  }
  return QB200_OK;
}

// second buffer per shard for the out-of-place exchange; collective decision in mp mode
static int ensure_alt(qb200_sv* sv) {
  if (sv->have_alt) return QB200_OK;
  if (sv->alt_failed || sv->cur != 0) return QB200_ERR_OOM;  // cur only leaves slot 0 once the spare exists
  bool ok = true;
  for (auto& s : sv->sh) {
    if (s.buf[1]) continue;
    if (qb200_state_alloc_on(s.ctx, sv->nl, sv->dtype, &s.buf[1]) != QB200_OK) { ok = false; break; }
  }
  if (sv->mp) {
    double v = ok ? 0.0 : 1.0;
    SV_TRY(sum_over_ranks(sv, &v, 1));
    ok = v == 0.0;
  }
  if (!ok) {
    for (auto& s : sv->sh)
      if (s.buf[1]) { DevScope d(s.device); cudaFree(s.buf[1]); s.buf[1] = nullptr; }
    sv->alt_failed = true;
    return QB200_ERR_OOM;
  }
  if (sv->mp) {
    std::vector<void*> table(sv->P, nullptr);
    SV_TRY(allgather_ipc(sv, sv->sh[0].buf[1], &table));
    for (unsigned r = 0; r < sv->P; ++r) sv->peer_buf[1][r] = table[r];
  } else {
    for (auto& s : sv->sh) sv->peer_buf[1][s.rank] = s.buf[1];
  }
  sv->have_alt = true;
  return QB200_OK;
}

// ---- local gate on every shard ----------------------------------------------------------------------------
template <typename FP>
static void permute_matrix(const FP* m, unsigned nq, const unsigned* order, std::vector<FP>* out) {
  // new index bit j <-> old index bit order[j]
  const unsigned dim = 1u << nq;
  std::vector<unsigned> map(dim);
  for (unsigned a = 0; a < dim; ++a) {
    unsigned o = 0;
    for (unsigned j = 0; j < nq; ++j) o |= ((a >> j) & 1u) << order[j];
    map[a] = o;
  }
  out->resize(size_t{2} * dim * dim);
  for (unsigned r = 0; r < dim; ++r)
    for (unsigned c = 0; c < dim; ++c) {
      (*out)[2 * (size_t{r} * dim + c)] = m[2 * (size_t{map[r]} * dim + map[c])];
      (*out)[2 * (size_t{r} * dim + c) + 1] = m[2 * (size_t{map[r]} * dim + map[c]) + 1];
    }
}

// Applies a gate (expect = false) or evaluates an operator (expect = true, result summed over the local shards
// into out[2]) whose TARGETS are all local.  Controls may be global: they select the shards that take part.
// chunk_bits / chunk_val (gates only): the pass is restricted to the part of every shard whose PHYSICAL local bits
// chunk_bits[] carry chunk_val -- extra controls of the kernel, none of them a target or control of the gate.
static int local_gate(qb200_sv* sv, const unsigned* qs, unsigned nq, const unsigned* cqs, unsigned nc,
                      uint64_t cvals, const void* matrix, bool expect, double* out,
                      const unsigned* chunk_bits = nullptr, unsigned nchunk = 0, unsigned chunk_val = 0) {
  if (nq > kMaxTargets) return QB200_ERR_UNSUPPORTED;
  if (nc > 0 && nq > kMaxCtrlTargets) return QB200_ERR_UNSUPPORTED;
  if (!expect) sv->fresh = false;
  unsigned phys[kMaxTargets], order[kMaxTargets], sorted[kMaxTargets];
  for (unsigned j = 0; j < nq; ++j) {
    if (qs[j] >= sv->n) return QB200_ERR_INVALID;
    phys[j] = sv->pos[qs[j]];
    if (phys[j] >= sv->nl) return QB200_ERR_INVALID;  // caller must have made it local
    order[j] = j;
  }
  std::sort(order, order + nq, [&](unsigned a, unsigned b) { return phys[a] < phys[b]; });
  bool identity = true;
  for (unsigned j = 0; j < nq; ++j) {
    sorted[j] = phys[order[j]];
    identity &= order[j] == j;
  }
  std::vector<float> mf;
  std::vector<double> md;
  const void* m = matrix;
  if (!identity) {
    if (sv->dtype == QB200_F32) { permute_matrix((const float*) matrix, nq, order, &mf); m = mf.data(); }
    else { permute_matrix((const double*) matrix, nq, order, &md); m = md.data(); }
  }
  // controls: bit i of cvals <-> i-th lowest control qubit (lib/simulator.h:364-375)
  std::vector<std::pair<unsigned, unsigned>> ctl;  // (logical, value)
  {
    std::vector<unsigned> c(cqs, cqs + nc);
    std::vector<unsigned> idx(nc);
    for (unsigned i = 0; i < nc; ++i) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](unsigned a, unsigned b) { return c[a] < c[b]; });
    for (unsigned i = 0; i < nc; ++i) {
      if (c[idx[i]] >= sv->n) return QB200_ERR_INVALID;
      ctl.emplace_back(c[idx[i]], (unsigned) ((cvals >> i) & 1));
    }
  }
  std::vector<std::pair<unsigned, unsigned>> lctl;  // (physical local bit, value)
  uint64_t gmask = 0, gbits = 0;
  for (auto& cv : ctl) {
    const unsigned p = sv->pos[cv.first];
    if (p >= sv->nl) {
      gmask |= uint64_t{1} << (p - sv->nl);
      gbits |= uint64_t{cv.second} << (p - sv->nl);
    } else {
      lctl.emplace_back(p, cv.second);
    }
  }
  for (unsigned j = 0; j < nchunk; ++j) lctl.emplace_back(chunk_bits[j], (chunk_val >> j) & 1u);
  std::sort(lctl.begin(), lctl.end());
  unsigned lc[64];
  uint64_t lcv = 0;
  for (size_t i = 0; i < lctl.size(); ++i) {
    lc[i] = lctl[i].first;
    lcv |= uint64_t{lctl[i].second} << i;
  }
  if (expect) {
    for (auto& s : sv->sh) SV_TRY(qb200_reduce_batch_begin(s.ctx, 1));
  }
  int rc = QB200_OK;
  for (auto& s : sv->sh) {
    if ((s.rank & gmask) != gbits) continue;
    if (expect) {
      double dummy[2];
      rc = qb200_expectation_value(s.ctx, sv->dtype, cur_buf(sv, s), sv->nl, sorted, nq, m, dummy);
    } else if (nchunk) {
      // the chunk bits are kernel-level controls: the reference's "at most 4 targets under control" does not apply
      rc = sv->dtype == QB200_F32
          ? gate_apply_f32(s.ctx, (float*) cur_buf(sv, s), sv->nl, sorted, nq, lc, (unsigned) lctl.size(), lcv, (const float*) m, nullptr)
          : gate_apply_f64(s.ctx, (double*) cur_buf(sv, s), sv->nl, sorted, nq, lc, (unsigned) lctl.size(), lcv, (const double*) m, nullptr);
    } else {
      rc = qb200_apply_controlled_gate(s.ctx, sv->dtype, cur_buf(sv, s), sv->nl, sorted, nq, lc,
                                       (unsigned) lctl.size(), lcv, m);
    }
    if (rc) break;
  }
  if (expect) {
    out[0] = out[1] = 0;
    for (auto& s : sv->sh) {
      double r[4] = {0, 0, 0, 0};
      uint32_t cnt = 0;
      int rc2 = qb200_reduce_batch_end(s.ctx, r, 2, &cnt);
      if (rc2 && !rc) rc = rc2;
      if (cnt) { out[0] += r[0]; out[1] += r[1]; }
    }
  } else if (!nchunk || chunk_val == 0) {
    ++sv->stats.gate_passes;
  }
  return rc;
}

static const float kSwapF[32] = {1, 0, 0, 0, 0, 0, 0, 0,  0, 0, 0, 0, 1, 0, 0, 0,
                                 0, 0, 1, 0, 0, 0, 0, 0,  0, 0, 0, 0, 0, 0, 1, 0};
static const double kSwapD[32] = {1, 0, 0, 0, 0, 0, 0, 0,  0, 0, 0, 0, 1, 0, 0, 0,
                                  0, 0, 1, 0, 0, 0, 0, 0,  0, 0, 0, 0, 0, 0, 1, 0};

// exchanges two LOCAL physical bits with one 2-qubit SWAP pass (exact: products with 0 and 1 on the FFMA kernels)
static int local_bit_swap(qb200_sv* sv, unsigned pa, unsigned pb) {
  if (pa == pb) return QB200_OK;
  const unsigned qa = qubit_at(sv, pa), qb = qubit_at(sv, pb);
  unsigned bits[2] = {std::min(pa, pb), std::max(pa, pb)};
  for (auto& s : sv->sh)
    SV_TRY(qb200_apply_gate(s.ctx, sv->dtype, cur_buf(sv, s), sv->nl, bits, 2,
                            sv->dtype == QB200_F32 ? (const void*) kSwapF : (const void*) kSwapD));
  sv->pos[qa] = pb;
  sv->pos[qb] = pa;
  ++sv->stats.local_passes;
  return QB200_OK;
}

// ---- the exchange ---------------------------------------------------------------------------------------------
static unsigned with_bits(unsigned rank, const unsigned* gb, unsigned k, unsigned v) {
  for (unsigned j = 0; j < k; ++j) rank = (rank & ~(1u << gb[j])) | (((v >> j) & 1u) << gb[j]);
  return rank;
}
static unsigned pick_bits(unsigned rank, const unsigned* gb, unsigned k) {
  unsigned v = 0;
  for (unsigned j = 0; j < k; ++j) v |= ((rank >> gb[j]) & 1u) << j;
  return v;
}

// Four events per exchange on the first local shard's stream: [0] before the leading barrier, [1] kernel start,
// [2] kernel end, [3] after the trailing barrier.  exchange_ms = [1]..[2]; wait_ms = the two barrier stretches.
static void timing_mark(qb200_sv* sv, int which) {
  DevScope d(sv->sh[0].device);
  if (which == 0 && sv->ev_used == sv->ev_quads.size()) {
    std::array<cudaEvent_t, 4> q{};
    for (auto& e : q) cudaEventCreate(&e);
    sv->ev_quads.push_back(q);
    sv->ev_overlapped.push_back(0);
  }
  if (which == 0) sv->ev_overlapped[sv->ev_used] = 0;
  cudaEventRecord(sv->ev_quads[sv->ev_used][which], sv->sh[0].stream);
  if (which == 3) ++sv->ev_used;
}
static int timing_collect(qb200_sv* sv) {
  SV_TRY(sync_all(sv));
  for (size_t i = 0; i < sv->ev_used; ++i) {
    const auto& q = sv->ev_quads[i];
    float pre = 0, run = 0, post = 0;
    if (cudaEventElapsedTime(&pre, q[0], q[1]) == cudaSuccess && cudaEventElapsedTime(&run, q[1], q[2]) == cudaSuccess &&
        cudaEventElapsedTime(&post, q[2], q[3]) == cudaSuccess) {
      if (sv->ev_overlapped[i]) sv->stats.overlap_ms += run;
      else sv->stats.exchange_ms += run;
      sv->stats.wait_ms += pre + post;
    } else {
      (void) cudaGetLastError();
    }
  }
  sv->ev_used = 0;
  return QB200_OK;
}

// ---- out-of-place exchange in three steps: prepare (geometry + spare buffers), launch per shard, finish (map) ----
struct PushPlan {
  RemapGeom rg{};                        // my / dst are filled per shard at launch
  std::vector<unsigned> victims, incoming;   // logical qubits, pairs ordered by the victims' physical bits
  unsigned gb[kMaxGlobal] = {};          // rank bit victim j goes to
  unsigned k = 0;
  size_t smem = 0;                       // bytes of one tile
};

// QB200_ERR_UNSUPPORTED: this exchange cannot go out of place (tiny shard or no room for the spare buffers)
static int push_prepare(qb200_sv* sv, const std::vector<unsigned>& victims, const std::vector<unsigned>& incoming,
                        PushPlan* pp, unsigned tile_bits = kTileBits) {
  const unsigned k = (unsigned) victims.size();
  if (sv->swap_mode == 0) return QB200_ERR_UNSUPPORTED;
  // the push kernel sorts tiles of 2^T amplitudes by the victim bits they contain: every destination run must
  // hold at least one 16-byte item (a shard of a handful of qubits with all of them victims does not qualify)
  const unsigned T = std::min<unsigned>(tile_bits, sv->nl);
  unsigned kl = 0;
  for (unsigned j = 0; j < k; ++j) kl += sv->pos[victims[j]] < T;
  if (T < kl + (sv->dtype == QB200_F32 ? 1u : 0u)) return sv->swap_mode == 1 ? QB200_ERR_INVALID : QB200_ERR_UNSUPPORTED;
  if (ensure_alt(sv) != QB200_OK) return sv->swap_mode == 1 ? QB200_ERR_OOM : QB200_ERR_UNSUPPORTED;
  pp->victims = victims;
  pp->incoming = incoming;
  pp->k = k;
  RemapGeom& rg = pp->rg;
  rg = RemapGeom{};
  rg.k = k;
  rg.nl = sv->nl;
  rg.T = T;   // tiles of 2^T amplitudes; tiny shards shrink the tile
  rg.kl = kl;
  for (unsigned j = 0; j < k; ++j) {
    rg.lbits[j] = sv->pos[victims[j]];
    pp->gb[j] = sv->pos[incoming[j]] - sv->nl;
  }
  rg.tiles = uint64_t{1} << (sv->nl - rg.T);
  rg.walk = rg.tiles;
  pp->smem = (size_t{1} << rg.T) * 2 * scalar_size(sv->dtype);
  return QB200_OK;
}

// shared memory of a push CTA: `budget` bytes cut into input stages (+ 2 output stages when tiles are sorted)
static uint32_t push_stages(const RemapGeom& rg, size_t tile_bytes, size_t budget) {
  const size_t out = rg.kl ? 2 * tile_bytes : 0;
  size_t sin = budget > out ? (budget - out) / tile_bytes : 0;
  const size_t lo = rg.kl ? 2 : 3;
  if (sin < lo) sin = lo;
  if (sin > (size_t) kMaxPushStages) sin = kMaxPushStages;
  return (uint32_t) sin;
}

template <typename FP, int NT>
static void launch_push_tma(qb200_sv* sv, Shard& s, const RemapGeom& rg, size_t tile_bytes, size_t budget, int ctas_per_sm,
                            cudaStream_t stream) {
  static PerDevice attr;
  attr.get(s.ctx, [&] {
    cudaFuncSetAttribute(k_remap_push_tma<FP, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    return 1;
  });
  const uint32_t sin = push_stages(rg, tile_bytes, budget);
  const size_t smem = (sin + (rg.kl ? 2 : 0)) * tile_bytes;
  const uint64_t b2 = std::min<uint64_t>(rg.walk, uint64_t{kNumSMs} * ctas_per_sm);
  k_remap_push_tma<FP, NT><<<(uint32_t) b2, NT, smem, stream>>>((const FP*) s.buf[sv->cur], rg, sin);
}

// one shard's push (the whole shard, or the chunk pp.rg describes) on `stream`; slim = the 64-thread variant that
// fits beside resident gate kernels
static int push_launch(qb200_sv* sv, Shard& s, PushPlan& pp, cudaStream_t stream, bool slim) {
  DevScope d(s.device);
  note_state_written();
  RemapGeom& rg = pp.rg;
  const int nb = 1 - sv->cur;
  rg.my = pick_bits(s.rank, pp.gb, pp.k);
  for (unsigned v = 0; v < (1u << pp.k); ++v) rg.dst[v] = sv->peer_buf[nb][with_bits(s.rank, pp.gb, pp.k, v)];
  if (sv->push_kernel == 1) {
    const bool f32 = sv->dtype == QB200_F32;
    if (slim) {
      // A small footprint matters more than a deep ring: the SM's L1 / shared split follows the resident CTAs, and the
      // tensor-core gate kernels lose 18 % when a neighbour forces the largest carveout (their 64-byte row loads live
      // on L1 sector merging) -- 16 KB (four 4 KB tiles) stay inside the carveout three k_gate_tca<4> CTAs get anyway.
      const int per_sm = sv->overlap_ctas_per_sm > 0 ? sv->overlap_ctas_per_sm : 1;
      const size_t budget = size_t{1024} * (sv->overlap_smem_kb > 0 ? sv->overlap_smem_kb : 16);
      if (f32) launch_push_tma<float, 64>(sv, s, rg, pp.smem, budget, per_sm, stream);
      else launch_push_tma<double, 64>(sv, s, rg, pp.smem, budget, per_sm, stream);
    } else {
      // alone: three CTAs per SM (fp64: one) with 4 tiles each
      const int per_sm = sv->push_ctas_per_sm > 0 ? sv->push_ctas_per_sm : (f32 ? 3 : 1);
      const size_t budget = 4 * pp.smem;
      if (f32) launch_push_tma<float, kRemapThreads>(sv, s, rg, pp.smem, budget, per_sm, stream);
      else launch_push_tma<double, kRemapThreads>(sv, s, rg, pp.smem, budget, per_sm, stream);
    }
  } else {
    const uint64_t blocks = std::min<uint64_t>(rg.walk, uint64_t{kNumSMs} * 6);
    if (sv->dtype == QB200_F32)
      k_remap_push<float><<<(uint32_t) blocks, kRemapThreads, pp.smem, stream>>>((const float*) s.buf[sv->cur], rg);
    else
      k_remap_push<double><<<(uint32_t) blocks, kRemapThreads, pp.smem, stream>>>((const double*) s.buf[sv->cur], rg);
  }
  ++s.ctx->launches;
  SV_CUDA(sv, cudaPeekAtLastError());
  return QB200_OK;
}

// ---- the same exchange on the COPY ENGINES ----------------------------------------------------------------------
// With every victim bit at or above bit 12 the amplitudes that go to one destination form runs of >= 32 KB, and the
// (source, destination) address pattern of a whole shard -- or of one chunk of it -- is a pitched 3-D box: the free
// index bits fall into runs between the pinned bits (victims: the destination's value; chunk bits: the chunk's),
// run 0 is the contiguous row, runs 1 and 2 are the box's height and depth (cudaMemcpy3DAsync), anything above is a
// short host loop.  The destination has the same runs at the packed positions (victim bits squeezed out).  No SM, no
// shared memory, no L1 carveout: the gate kernels that run beside an overlapped exchange keep the SMs to themselves
// (the SM-driven slim kernel time-shares them, profiles/r02_overlap_trace.txt).
constexpr unsigned kCeMinBit = 12;
constexpr unsigned kCeMaxLoop = 256;

struct CeRun { unsigned start, len; };

// index of the run that becomes the height of the 2-D copies (-1: none qualifies)
static int ce_box_run(const qb200_sv* sv, const std::vector<CeRun>& runs) {
  const size_t ab = 2 * scalar_size(sv->dtype);
  int best = -1;
  for (size_t r = 1; r < runs.size(); ++r) {
    if (((uint64_t{1} << runs[r].start) * ab) >= (uint64_t{1} << 31)) continue;
    if (best < 0 || runs[r].len > runs[best].len) best = (int) r;
  }
  return best;
}

static bool ce_runs(const qb200_sv* sv, const PushPlan& pp, const unsigned* chunk_bits, unsigned nchunk,
                    std::vector<CeRun>* runs) {
  std::vector<unsigned> sp(pp.rg.lbits, pp.rg.lbits + pp.k);
  for (unsigned j = 0; j < pp.k; ++j)
    if (sp[j] < kCeMinBit) return false;
  for (unsigned j = 0; j < nchunk; ++j) sp.push_back(chunk_bits[j]);
  std::sort(sp.begin(), sp.end());
  runs->clear();
  unsigned lo = 0;
  for (unsigned b : sp) {
    if (b > lo) runs->push_back({lo, b - lo});
    lo = b + 1;
  }
  if (sv->nl > lo) runs->push_back({lo, sv->nl - lo});
  if (runs->empty() || (*runs)[0].start != 0) return false;
  // the longest run above the row becomes the height of a pitched 2-D copy (pitch below 2 GiB), the others are looped
  // over on the host (3-D copies of linear memory do not run at copy-engine speed: 165 GB/s measured)
  const int boxed = ce_box_run(sv, *runs);
  uint64_t loops = 1;
  for (size_t r = 1; r < runs->size(); ++r)
    if ((int) r != boxed) loops <<= (*runs)[r].len;
  return loops <= kCeMaxLoop;
}

// one shard's amplitudes (all, or the chunk whose bits chunk_bits[] carry cval) to their places in the new layout
static int ce_push(qb200_sv* sv, Shard& s, const PushPlan& pp, const unsigned* chunk_bits, unsigned nchunk,
                   unsigned cval, cudaStream_t stream) {
  DevScope d(s.device);
  note_state_written();
  std::vector<CeRun> runs;
  if (!ce_runs(sv, pp, chunk_bits, nchunk, &runs)) return QB200_ERR_UNSUPPORTED;
  const size_t ab = 2 * scalar_size(sv->dtype);
  const int nb = 1 - sv->cur;
  const unsigned k = pp.k, my = pick_bits(s.rank, pp.gb, k);
  auto dpos = [&](unsigned bit) {   // position of a non-victim source bit in the destination index
    unsigned below = 0;
    for (unsigned j = 0; j < k; ++j) below += pp.rg.lbits[j] < bit;
    return bit - below;
  };
  uint64_t src_fixed = 0, dst_fixed = uint64_t{my} << (sv->nl - k);
  for (unsigned j = 0; j < nchunk; ++j) {
    src_fixed |= uint64_t{(cval >> j) & 1u} << chunk_bits[j];
    dst_fixed |= uint64_t{(cval >> j) & 1u} << dpos(chunk_bits[j]);
  }
  std::vector<CeRun> box, loop;
  const int boxed = ce_box_run(sv, runs);
  for (size_t r = 1; r < runs.size(); ++r) ((int) r == boxed ? box : loop).push_back(runs[r]);
  uint64_t nloop = 1;
  for (const auto& r : loop) nloop <<= r.len;
  const size_t width = (size_t{1} << runs[0].len) * ab;
  for (unsigned v = 0; v < (1u << k); ++v) {
    uint64_t src_v = src_fixed;
    for (unsigned j = 0; j < k; ++j) src_v |= uint64_t{(v >> j) & 1u} << pp.rg.lbits[j];
    char* const dst_buf = (char*) sv->peer_buf[nb][with_bits(s.rank, pp.gb, k, v)];
    const char* const src_buf = (const char*) s.buf[sv->cur];
    for (uint64_t it = 0; it < nloop; ++it) {
      uint64_t so = src_v, dof = dst_fixed, rest = it;
      for (const auto& r : loop) {
        const uint64_t val = rest & ((uint64_t{1} << r.len) - 1);
        rest >>= r.len;
        so |= val << r.start;
        dof |= val << dpos(r.start);
      }
      cudaError_t e;
      if (box.empty()) {
        e = cudaMemcpyAsync(dst_buf + dof * ab, src_buf + so * ab, width, cudaMemcpyDefault, stream);
      } else {
        const size_t spitch = (size_t{1} << box[0].start) * ab, dpitch = (size_t{1} << dpos(box[0].start)) * ab;
        const size_t height = size_t{1} << box[0].len;
        e = cudaMemcpy2DAsync(dst_buf + dof * ab, dpitch, src_buf + so * ab, spitch, width, height, cudaMemcpyDefault, stream);
      }
      if (e != cudaSuccess) {
        sv->last_error = (int) e;
        (void) cudaGetLastError();
        return QB200_ERR_CUDA;
      }
    }
  }
  return QB200_OK;
}

// after the trailing barrier: the spare buffers hold the state; new qubit map
static void push_finish(qb200_sv* sv, const PushPlan& pp) {
  sv->cur = 1 - sv->cur;
  // the remaining local qubits keep their order in the low bits, incoming j sits at nl - k + j
  const unsigned k = pp.k;
  std::vector<unsigned> vb(pp.rg.lbits, pp.rg.lbits + k);
  for (unsigned q = 0; q < sv->n; ++q) {
    const unsigned p = sv->pos[q];
    if (p >= sv->nl) continue;
    if (std::find(pp.victims.begin(), pp.victims.end(), q) != pp.victims.end()) continue;
    unsigned below = 0;
    for (unsigned b : vb) below += b < p;
    sv->pos[q] = p - below;
  }
  for (unsigned j = 0; j < k; ++j) {
    sv->pos[pp.victims[j]] = sv->nl + pp.gb[j];
    sv->pos[pp.incoming[j]] = sv->nl - k + j;
  }
  ++sv->stats.swaps;
  sv->stats.bytes_sent_per_shard += (double) shard_bytes(sv) * (1.0 - 1.0 / (double) (1u << k));
}

// validates an exchange and orders its (victim, incoming) pairs by the victims' physical bits (the kernels want
// those ascending); victim j goes to the rank bit incoming j leaves
static int order_pairs(qb200_sv* sv, const unsigned* victims_in, const unsigned* incoming_in, unsigned k,
                       std::vector<unsigned>* victims, std::vector<unsigned>* incoming) {
  if (k > sv->g) return QB200_ERR_INVALID;
  for (unsigned j = 0; j < k; ++j) {
    if (victims_in[j] >= sv->n || incoming_in[j] >= sv->n) return QB200_ERR_INVALID;
    if (sv->pos[victims_in[j]] >= sv->nl || sv->pos[incoming_in[j]] < sv->nl) return QB200_ERR_INVALID;
  }
  std::vector<unsigned> perm(k);
  for (unsigned j = 0; j < k; ++j) perm[j] = j;
  std::sort(perm.begin(), perm.end(), [&](unsigned a, unsigned b) { return sv->pos[victims_in[a]] < sv->pos[victims_in[b]]; });
  victims->resize(k);
  incoming->resize(k);
  for (unsigned j = 0; j < k; ++j) {
    (*victims)[j] = victims_in[perm[j]];
    (*incoming)[j] = incoming_in[perm[j]];
  }
  return QB200_OK;
}

// victims (logical, local) <-> incoming (logical, global); k <= g.
static int exchange(qb200_sv* sv, const unsigned* victims_in, const unsigned* incoming_in, unsigned k) {
  if (k == 0) return QB200_OK;
  sv->fresh = false;
  std::vector<unsigned> victims, incoming;
  SV_TRY(order_pairs(sv, victims_in, incoming_in, k, &victims, &incoming));
  unsigned gb[kMaxGlobal];
  for (unsigned j = 0; j < k; ++j) gb[j] = sv->pos[incoming[j]] - sv->nl;
  PushPlan pp;
  const int prc = push_prepare(sv, victims, incoming, &pp);
  if (prc != QB200_OK && prc != QB200_ERR_UNSUPPORTED) return prc;
  const double sent = (double) shard_bytes(sv) * (1.0 - 1.0 / (double) (1u << k));

  if (prc == QB200_OK) {
    timing_mark(sv, 0);
    timing_mark(sv, 1);  // no leading barrier: the spare buffers are free (see ensure_alt)
    std::vector<CeRun> runs;
    const bool ce = sv->push_kernel == 2 && ce_runs(sv, pp, nullptr, 0, &runs);
    for (auto& s : sv->sh) {
      if (ce) SV_TRY(ce_push(sv, s, pp, nullptr, 0, 0, s.stream));
      else SV_TRY(push_launch(sv, s, pp, s.stream, false));
    }
    timing_mark(sv, 2);  // (recorded after the launches of every local shard; the first shard's stream only holds its own)
    SV_TRY(barrier(sv));
    timing_mark(sv, 3);
    push_finish(sv, pp);
    sv->stats.ce_swaps += ce;
    return QB200_OK;
  } else {
    // in place (k_p2p_swap): local bit <-> rank bit; low victims are lifted first, at most 3 bits per pass
    for (unsigned off = 0; off < k; off += 3) {
      const unsigned kk = std::min(3u, k - off);
      std::vector<unsigned> taken;
      for (unsigned j = 0; j < kk; ++j) taken.push_back(sv->pos[victims[off + j]]);
      for (unsigned j = 0; j < kk; ++j) {
        const unsigned p = sv->pos[victims[off + j]];
        if (p >= kMinInplaceBit || sv->nl <= kMinInplaceBit + kk) continue;
        unsigned dst = sv->nl;
        for (unsigned c = sv->nl; c-- > kMinInplaceBit;)
          if (std::find(taken.begin(), taken.end(), c) == taken.end()) { dst = c; break; }
        if (dst == sv->nl) continue;
        SV_TRY(local_bit_swap(sv, p, dst));
        *std::find(taken.begin(), taken.end(), p) = dst;
      }
      // the lifting may have changed the order of the victims' bits: sort the pairs again
      std::vector<unsigned> ord(kk);
      for (unsigned j = 0; j < kk; ++j) ord[j] = off + j;
      std::sort(ord.begin(), ord.end(), [&](unsigned a, unsigned b) { return sv->pos[victims[a]] < sv->pos[victims[b]]; });
      unsigned lb[3], g3[3];
      for (unsigned j = 0; j < kk; ++j) { lb[j] = sv->pos[victims[ord[j]]]; g3[j] = gb[ord[j]]; }
      timing_mark(sv, 0);
      SV_TRY(barrier(sv));
      timing_mark(sv, 1);
      for (auto& s : sv->sh) {
        void* peers[8] = {};
        const unsigned my = pick_bits(s.rank, g3, kk);
        for (unsigned v = 0; v < (1u << kk); ++v)
          if (v != my) peers[v] = sv->peer_buf[sv->cur][with_bits(s.rank, g3, kk, v)];
        SV_TRY(qb200_swap_global_local(s.ctx, sv->dtype, cur_buf(sv, s), sv->nl, peers, kk, lb, my));
      }
      timing_mark(sv, 2);
      SV_TRY(barrier(sv));
      timing_mark(sv, 3);
      for (unsigned j = 0; j < kk; ++j) {
        sv->pos[victims[ord[j]]] = sv->nl + g3[j];
        sv->pos[incoming[ord[j]]] = lb[j];
      }
    }
  }
  ++sv->stats.swaps;
  sv->stats.bytes_sent_per_shard += sent;
  return QB200_OK;
}

// ---- exchange overlapped with the last gates of its epoch ------------------------------------------------------
// The push is out of place, so a part of the shard can leave as soon as the last gate has passed over it.  The
// shard is cut into 2^c CHUNKS by c physical local bits that are neither victims nor touched by the last L gates of
// the epoch; chunk by chunk those L gates run as passes with the chunk bits as extra controls (main stream), and the
// chunk's push follows on a second stream -- while the gates are already on the next chunk.  A gate pass streams
// the chunk from HBM (2 x 16 B per amplitude at ~6 TB/s), the push moves (1 - 2^-k) of it at ~0.7 TB/s over NVLink:
// L >= 4.3 (1 - 2^-k) passes cover the exchange, so L is capped at overlap_max_gates.  Exposed: the gates of the
// first chunk and the push of the last one.  No extra synchronisation between GPUs: the destinations are the spare
// buffers, idle until the trailing barrier.
static void touch(qb200_sv* sv, const unsigned* qs, unsigned nq);

struct OverlapSpec {
  bool valid = false;
  uint64_t start = 0, swap = 0;         // steps[start .. swap) are the pipelined gates, steps[swap] the exchange
  unsigned nchunk = 0, chunk_bits[3] = {};
  std::vector<unsigned> victims, incoming;
};

static void plan_overlap(qb200_sv* sv, const qb200_gate* gates, const std::vector<int64_t>& steps, uint64_t from,
                         OverlapSpec* spec) {
  spec->valid = false;
  if (!sv->overlap || sv->swap_mode == 0 || sv->alt_failed || sv->P == 1) return;
  uint64_t ws = from;
  while (ws < steps.size() && steps[ws] >= 0) ++ws;
  if (ws >= steps.size() || ws == from) return;
  const unsigned k = (unsigned) -steps[ws];
  unsigned vq[kMaxGlobal], iq[kMaxGlobal];
  for (unsigned j = 0; j < k; ++j) vq[j] = (unsigned) steps[ws + 1 + j];
  for (unsigned j = 0; j < k; ++j) iq[j] = (unsigned) steps[ws + 1 + k + j];
  if (order_pairs(sv, vq, iq, k, &spec->victims, &spec->incoming) != QB200_OK) return;
  const unsigned T = std::min<unsigned>(kTileBits, sv->nl);
  const unsigned c = (unsigned) std::min(3, std::max(1, sv->overlap_chunks_log2));
  uint64_t free_bits = 0;
  for (unsigned p = T; p < sv->nl; ++p) free_bits |= uint64_t{1} << p;
  for (unsigned q : spec->victims) free_bits &= ~(uint64_t{1} << sv->pos[q]);
  if ((unsigned) __builtin_popcountll(free_bits) < c) return;
  unsigned L = 0;
  for (uint64_t i = ws; i-- > from && (int) L < sv->overlap_max_gates;) {
    const qb200_gate& gt = gates[steps[i]];
    if (gt.num_targets > 5 || (gt.num_targets == 5 && sv->dtype != QB200_F32)) break;
    uint64_t touched = 0;
    bool ok = true;
    for (unsigned j = 0; j < gt.num_targets; ++j) {
      const unsigned p = sv->pos[gt.qs[j]];
      if (p >= sv->nl) { ok = false; break; }   // (cannot happen in a planned epoch)
      touched |= uint64_t{1} << p;
    }
    for (unsigned j = 0; j < gt.num_controls && ok; ++j) {
      const unsigned p = sv->pos[gt.cqs[j]];
      if (p < sv->nl) touched |= uint64_t{1} << p;
    }
    if (!ok || (unsigned) __builtin_popcountll(free_bits & ~touched) < c) break;
    free_bits &= ~touched;
    ++L;
  }
  if (L == 0) return;
  for (unsigned j = c; j-- > 0;) {               // the highest free bits: the longest contiguous runs
    const unsigned p = 63 - (unsigned) __builtin_clzll(free_bits);
    spec->chunk_bits[j] = p;
    free_bits &= ~(uint64_t{1} << p);
  }
  spec->nchunk = c;
  spec->swap = ws;
  spec->start = ws - L;
  spec->valid = true;
}

static int run_overlapped(qb200_sv* sv, const qb200_gate* gates, const std::vector<int64_t>& steps, const OverlapSpec& spec,
                          PushPlan& pp) {
  RemapGeom& rg = pp.rg;
  const unsigned c = spec.nchunk, kh = rg.k - rg.kl;
  rg.nchunk = c;
  for (unsigned j = 0; j < c; ++j) {
    unsigned below = 0;
    for (unsigned i = rg.kl; i < rg.k; ++i) below += rg.lbits[i] < spec.chunk_bits[j];
    rg.cpos[j] = kh + (spec.chunk_bits[j] - rg.T) - below;   // position in the tile counter: [packed tile bits | kh victim bits]
  }
  rg.walk = rg.tiles >> c;
  timing_mark(sv, 0);
  timing_mark(sv, 1);
  sv->ev_overlapped[sv->ev_used] = 1;
  for (auto& s : sv->sh) s.ctx->occ_reduce = sv->overlap_occ_reduce;
  int rc = QB200_OK;
  std::vector<CeRun> ce_probe;
  const bool ce = sv->overlap_ce != 0 && ce_runs(sv, pp, spec.chunk_bits, c, &ce_probe);
  if (sv->overlap_trace) {
    fprintf(stderr, "overlap trace: victims at bits");
    for (unsigned j = 0; j < rg.k; ++j) fprintf(stderr, " %u", rg.lbits[j]);
    fprintf(stderr, ", chunk bits");
    for (unsigned j = 0; j < c; ++j) fprintf(stderr, " %u", spec.chunk_bits[j]);
    fprintf(stderr, ", copy engines %d, runs", (int) ce);
    for (const auto& r : ce_probe) fprintf(stderr, " [%u,+%u)", r.start, r.len);
    fprintf(stderr, "\n");
  }
  // debugging aid: CUDA events around every chunk's gates (main stream) and push (second stream) of the first shard
  std::vector<cudaEvent_t> tr;
  const bool trace = sv->overlap_trace != 0;
  auto mark = [&](cudaStream_t st) {
    if (!trace) return;
    DevScope d(sv->sh[0].device);
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    tr.push_back(e);
  };
  for (unsigned v = 0; v < (1u << c) && rc == QB200_OK; ++v) {
    mark(sv->sh[0].stream);
    for (uint64_t i = spec.start; i < spec.swap && rc == QB200_OK; ++i) {
      const qb200_gate& gt = gates[steps[i]];
      if (v == 0) touch(sv, gt.qs, gt.num_targets);
      rc = local_gate(sv, gt.qs, gt.num_targets, gt.cqs, gt.num_controls, gt.cvals, gt.matrix, false, nullptr,
                      spec.chunk_bits, c, v);
      if (trace && v == 0) fprintf(stderr, "overlap trace: gate %s\n", sv->sh[0].ctx->last_kernel);
    }
    mark(sv->sh[0].stream);
    rg.cval = v;
    for (auto& s : sv->sh) {
      if (rc != QB200_OK) break;
      DevScope d(s.device);
      if (cudaEventRecord(s.ev_chunk, s.stream) != cudaSuccess || cudaStreamWaitEvent(s.stream2, s.ev_chunk, 0) != cudaSuccess) {
        (void) cudaGetLastError();
        rc = QB200_ERR_CUDA;
        break;
      }
      if (&s == &sv->sh[0]) mark(s.stream2);
      rc = ce ? ce_push(sv, s, pp, spec.chunk_bits, c, v, s.stream2) : push_launch(sv, s, pp, s.stream2, true);
      if (&s == &sv->sh[0]) mark(s.stream2);
    }
  }
  if (trace && rc == QB200_OK) {
    sv->overlap_trace = 0;
    sync_all(sv);
    { DevScope d(sv->sh[0].device); cudaStreamSynchronize(sv->sh[0].stream2); }
    for (unsigned v = 0; v < (1u << c); ++v) {
      float t[4] = {};
      for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], tr[0], tr[4 * v + i]);
      fprintf(stderr, "overlap trace: chunk %u gates %.3f..%.3f ms, push %.3f..%.3f ms\n", v, t[0], t[1], t[2], t[3]);
    }
  }
  for (auto e : tr) cudaEventDestroy(e);
  for (auto& s : sv->sh) {
    s.ctx->occ_reduce = 0;
    DevScope d(s.device);
    cudaEventRecord(s.ev_join, s.stream2);
    cudaStreamWaitEvent(s.stream, s.ev_join, 0);
  }
  if (rc != QB200_OK) return rc;
  timing_mark(sv, 2);
  SV_TRY(barrier(sv));
  timing_mark(sv, 3);
  push_finish(sv, pp);
  ++sv->stats.overlapped_swaps;
  sv->stats.ce_swaps += ce;
  sv->stats.overlapped_gate_passes += spec.swap - spec.start;
  return QB200_OK;
}

// brings every qubit of `qs` into the local part with one exchange; victims: LRU (online) local qubits outside qs
static int make_local(qb200_sv* sv, const unsigned* qs, unsigned nq) {
  std::vector<unsigned> incoming;
  for (unsigned j = 0; j < nq; ++j) {
    if (qs[j] >= sv->n) return QB200_ERR_INVALID;
    if (sv->pos[qs[j]] >= sv->nl) incoming.push_back(qs[j]);
  }
  if (incoming.empty()) return QB200_OK;
  if (nq > sv->nl) return QB200_ERR_INVALID;
  std::vector<unsigned> cand;
  for (unsigned q = 0; q < sv->n; ++q)
    if (sv->pos[q] < sv->nl && std::find(qs, qs + nq, q) == qs + nq) cand.push_back(q);
  std::stable_sort(cand.begin(), cand.end(), [&](unsigned a, unsigned b) {
    if (sv->last_use[a] != sv->last_use[b]) return sv->last_use[a] < sv->last_use[b];
    return sv->pos[a] > sv->pos[b];
  });
  if (cand.size() < incoming.size()) return QB200_ERR_INVALID;
  cand.resize(incoming.size());
  return exchange(sv, cand.data(), incoming.data(), (unsigned) incoming.size());
}

// restores pos[q] = q: at most two exchanges for the rank bits, then 2-qubit SWAP passes for the local bits
static int canonicalize(qb200_sv* sv) {
  if (canonical(sv)) return QB200_OK;
  for (int iter = 0; iter < 4; ++iter) {
    std::vector<unsigned> victims, incoming, stuck;
    for (unsigned t = 0; t < sv->g; ++t) {
      const unsigned want = sv->nl + t, have = qubit_at(sv, sv->nl + t);
      if (have == want) continue;
      if (sv->pos[want] < sv->nl) { victims.push_back(want); incoming.push_back(have); }
      else stuck.push_back(have);
    }
    if (victims.empty() && stuck.empty()) break;
    if (!victims.empty()) {
      SV_TRY(exchange(sv, victims.data(), incoming.data(), (unsigned) victims.size()));
      continue;
    }
    // the wanted qubits sit on other rank bits: bring all of them local first
    std::vector<unsigned> spare;
    for (unsigned p = sv->nl; p-- > 0 && spare.size() < stuck.size();) {
      const unsigned q = qubit_at(sv, p);
      if (q < sv->nl) spare.push_back(q);
    }
    if (spare.size() < stuck.size()) return QB200_ERR_INVALID;
    SV_TRY(exchange(sv, spare.data(), stuck.data(), (unsigned) stuck.size()));
  }
  for (unsigned b = 0; b < sv->nl; ++b)
    if (sv->pos[b] != b) SV_TRY(local_bit_swap(sv, b, sv->pos[b]));
  return canonical(sv) ? QB200_OK : QB200_ERR_INVALID;
}

static void touch(qb200_sv* sv, const unsigned* qs, unsigned nq) {
  ++sv->clock;
  for (unsigned j = 0; j < nq; ++j) sv->last_use[qs[j]] = sv->clock;
}

}  // namespace qb200

// =================================================================================================================
extern "C" {

int qb200_sv_create(const int* devices, unsigned num_shards, unsigned num_qubits, int dtype, qb200_sv** out) {
  if (!out || !devices) return QB200_ERR_INVALID;
  *out = nullptr;
  qb200_sv* sv = new (std::nothrow) qb200_sv();
  if (!sv) return QB200_ERR_OOM;
  int rc = init_common(sv, num_shards, num_qubits, dtype);
  for (unsigned r = 0; r < num_shards && rc == QB200_OK; ++r) rc = make_shard(sv, devices[r], r);
  if (rc == QB200_OK) {
    // peer access between distinct devices (shards on one device need none)
    for (auto& a : sv->sh)
      for (auto& b : sv->sh) {
        if (a.device == b.device) continue;
        DevScope d(a.device);
        int can = 0;
        cudaDeviceCanAccessPeer(&can, a.device, b.device);
        if (!can) { rc = QB200_ERR_UNSUPPORTED; break; }
        cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { rc = QB200_ERR_CUDA; break; }
        (void) cudaGetLastError();
      }
  }
  if (rc != QB200_OK) {
    qb200_sv_destroy(sv);
    return rc;
  }
  for (auto& s : sv->sh) {
    sv->peer_buf[0][s.rank] = s.buf[0];
    sv->peer_flags[s.rank] = s.flags;
  }
  *out = sv;
  return QB200_OK;
}

int qb200_sv_create_mp(int device, unsigned rank, unsigned world, const qb200_comm* comm, unsigned num_qubits,
                       int dtype, qb200_sv** out) {
  if (!out || !comm || !comm->allgather || !comm->allreduce_sum_f64 || !comm->barrier || rank >= world)
    return QB200_ERR_INVALID;
  *out = nullptr;
  qb200_sv* sv = new (std::nothrow) qb200_sv();
  if (!sv) return QB200_ERR_OOM;
  sv->mp = true;
  sv->rank = rank;
  sv->comm = *comm;
  sv->barrier_flags = 1;
  int rc = init_common(sv, world, num_qubits, dtype);
  if (rc == QB200_OK) rc = make_shard(sv, device, rank);
  // every rank takes the same path from here on, whatever happened locally
  double bad = rc == QB200_OK ? 0.0 : 1.0;
  if (comm->allreduce_sum_f64(comm->user, &bad, 1) != 0 || bad != 0.0) {
    qb200_sv_destroy(sv);
    return rc != QB200_OK ? rc : QB200_ERR_OOM;
  }
  if (world > 1) {
    rc = allgather_ipc(sv, sv->sh[0].buf[0], &sv->peer_buf[0]);
    if (rc == QB200_OK) {
      std::vector<void*> fl(world, nullptr);
      rc = allgather_ipc(sv, sv->sh[0].flags, &fl);
      for (unsigned r = 0; r < world; ++r) sv->peer_flags[r] = (uint32_t*) fl[r];
    }
    comm->barrier(comm->user);
  } else {
    sv->peer_buf[0][0] = sv->sh[0].buf[0];
    sv->peer_flags[0] = sv->sh[0].flags;
  }
  if (rc != QB200_OK) {
    qb200_sv_destroy(sv);
    return rc;
  }
  *out = sv;
  return QB200_OK;
}

int qb200_sv_destroy(qb200_sv* sv) {
  if (!sv) return QB200_OK;
  for (auto& s : sv->sh) {
    DevScope d(s.device);
    if (s.stream) cudaStreamSynchronize(s.stream);
  }
  if (sv->mp && sv->P > 1 && sv->comm.barrier) sv->comm.barrier(sv->comm.user);  // nobody unmaps what a peer still uses
  for (void* p : sv->ipc_opened) qb200_ipc_close(p);
  for (auto& s : sv->sh) {
    DevScope d(s.device);
    if (s.buf[0]) cudaFree(s.buf[0]);
    if (s.buf[1]) cudaFree(s.buf[1]);
    if (s.flags) cudaFree(s.flags);
    if (s.ev) cudaEventDestroy(s.ev);
    if (s.ev_chunk) cudaEventDestroy(s.ev_chunk);
    if (s.ev_join) cudaEventDestroy(s.ev_join);
    if (s.stream2) { cudaStreamSynchronize(s.stream2); cudaStreamDestroy(s.stream2); }
    if (s.ctx) qb200_ctx_destroy(s.ctx);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  for (auto& q : sv->ev_quads)
    for (auto e : q) cudaEventDestroy(e);
  if (sv->err_host) cudaFreeHost(sv->err_host);
  (void) cudaGetLastError();
  delete sv;
  return QB200_OK;
}

unsigned qb200_sv_num_qubits(const qb200_sv* sv) { return sv ? sv->n : 0; }
unsigned qb200_sv_num_shards(const qb200_sv* sv) { return sv ? sv->P : 0; }
unsigned qb200_sv_num_local_qubits(const qb200_sv* sv) { return sv ? sv->nl : 0; }
int qb200_sv_last_cuda_error(const qb200_sv* sv) { return sv ? sv->last_error : 0; }

int qb200_sv_qubit_map(const qb200_sv* sv, unsigned* pos) {
  if (!sv || !pos) return QB200_ERR_INVALID;
  for (unsigned q = 0; q < sv->n; ++q) pos[q] = sv->pos[q];
  return QB200_OK;
}

int qb200_sv_shard(const qb200_sv* sv, unsigned local_index, unsigned* rank, int* device, void** state, qb200_ctx** ctx) {
  if (!sv || local_index >= sv->sh.size()) return QB200_ERR_INVALID;
  const Shard& s = sv->sh[local_index];
  if (rank) *rank = s.rank;
  if (device) *device = s.device;
  if (state) *state = s.buf[sv->cur];
  if (ctx) *ctx = s.ctx;
  return QB200_OK;
}

unsigned qb200_sv_num_local_shards(const qb200_sv* sv) { return sv ? (unsigned) sv->sh.size() : 0; }

int qb200_sv_set_option(qb200_sv* sv, const char* key, int value) {
  if (!sv || !key) return QB200_ERR_INVALID;
  if (!std::strcmp(key, "swap_mode")) sv->swap_mode = value;
  else if (!std::strcmp(key, "reorder")) sv->reorder = value;
  else if (!std::strcmp(key, "push_kernel")) sv->push_kernel = value;
  else if (!std::strcmp(key, "push_ctas_per_sm")) sv->push_ctas_per_sm = value;
  else if (!std::strcmp(key, "overlap")) sv->overlap = value;
  else if (!std::strcmp(key, "overlap_chunks_log2")) sv->overlap_chunks_log2 = value;
  else if (!std::strcmp(key, "overlap_max_gates")) sv->overlap_max_gates = value;
  else if (!std::strcmp(key, "overlap_occ_reduce")) sv->overlap_occ_reduce = value;
  else if (!std::strcmp(key, "overlap_trace")) sv->overlap_trace = value;
  else if (!std::strcmp(key, "free_initial_map")) sv->free_initial_map = value;
  else if (!std::strcmp(key, "overlap_ce")) sv->overlap_ce = value;
  else if (!std::strcmp(key, "overlap_ctas_per_sm")) sv->overlap_ctas_per_sm = value;
  else if (!std::strcmp(key, "overlap_smem_kb")) sv->overlap_smem_kb = value;
  else if (!std::strcmp(key, "overlap_tile_bits")) sv->overlap_tile_bits = value;
  else if (!std::strcmp(key, "barrier_flags")) {
    if (sv->mp && !value) return QB200_ERR_INVALID;  // events do not cross processes
    sv->barrier_flags = value;
  } else {
    // anything else is a kernel tuning key of the per-shard contexts (qb200_ctx_set_tuning)
    for (auto& s : sv->sh) SV_TRY(qb200_ctx_set_tuning(s.ctx, key, value));
  }
  return QB200_OK;
}

int qb200_sv_sync(qb200_sv* sv) { return sv ? sync_all(sv) : QB200_ERR_INVALID; }

uint64_t qb200_sv_launch_count(const qb200_sv* sv) {
  uint64_t c = 0;
  if (sv) for (auto& s : sv->sh) c += s.ctx->launches;
  return c;
}

int qb200_sv_get_stats(qb200_sv* sv, qb200_sv_stats* out) {
  if (!sv || !out) return QB200_ERR_INVALID;
  SV_TRY(timing_collect(sv));
  out->swaps = sv->stats.swaps;
  out->local_swap_passes = sv->stats.local_passes;
  out->gate_passes = sv->stats.gate_passes;
  out->bytes_sent_per_shard = sv->stats.bytes_sent_per_shard;
  out->exchange_ms = sv->stats.exchange_ms;
  out->barrier_wait_ms = sv->stats.wait_ms;
  out->overlapped_swaps = sv->stats.overlapped_swaps;
  out->overlapped_gate_passes = sv->stats.overlapped_gate_passes;
  out->overlap_ms = sv->stats.overlap_ms;
  out->copy_engine_swaps = sv->stats.ce_swaps;
  return QB200_OK;
}

int qb200_sv_reset_stats(qb200_sv* sv) {
  if (!sv) return QB200_ERR_INVALID;
  SV_TRY(timing_collect(sv));
  sv->stats = SvStats();
  return QB200_OK;
}

// ---- state initialisation ---------------------------------------------------------------------------------------
int qb200_sv_set_all_zeros(qb200_sv* sv) {
  if (!sv) return QB200_ERR_INVALID;
  for (auto& s : sv->sh) SV_TRY(qb200_set_all_zeros(s.ctx, sv->dtype, cur_buf(sv, s), sv->nl));
  sv->fresh = true;
  return QB200_OK;
}

int qb200_sv_set_state_zero(qb200_sv* sv) {
  if (!sv) return QB200_ERR_INVALID;
  SV_TRY(qb200_sv_set_all_zeros(sv));
  // |0...0> is index 0 under every qubit map
  if (Shard* s = local_shard(sv, 0)) SV_TRY(qb200_set_ampl(s->ctx, sv->dtype, cur_buf(sv, *s), 0, 1.0, 0.0));
  sv->fresh = true;
  return QB200_OK;
}

int qb200_sv_set_state_uniform(qb200_sv* sv) {
  if (!sv) return QB200_ERR_INVALID;
  const double v = 1.0 / std::sqrt((double) (uint64_t{1} << sv->n));  // lib/statespace_cuda.h:122
  for (auto& s : sv->sh) SV_TRY(qb200_bulk_set_ampl(s.ctx, sv->dtype, cur_buf(sv, s), sv->nl, 0, 0, v, 0.0, 0));
  sv->fresh = true;
  return QB200_OK;
}

// resets the qubit map to the identity without moving data (only meaningful before the state is (re)initialised)
int qb200_sv_reset_map(qb200_sv* sv) {
  if (!sv) return QB200_ERR_INVALID;
  for (unsigned q = 0; q < sv->n; ++q) sv->pos[q] = q;
  return QB200_OK;
}

int qb200_sv_get_ampl(qb200_sv* sv, uint64_t i, double out[2]) {
  if (!sv || !out || (i >> sv->n)) return QB200_ERR_INVALID;
  const uint64_t p = to_physical(sv, i);
  double v[2] = {0, 0};
  if (Shard* s = local_shard(sv, (unsigned) (p >> sv->nl)))
    SV_TRY(qb200_get_ampl(s->ctx, sv->dtype, cur_buf(sv, *s), p & ((uint64_t{1} << sv->nl) - 1), v));
  SV_TRY(sum_over_ranks(sv, v, 2));
  out[0] = v[0];
  out[1] = v[1];
  return QB200_OK;
}

int qb200_sv_get_ampls(qb200_sv* sv, const uint64_t* indices, uint64_t count, double* out) {
  if (!sv || (count && (!indices || !out))) return QB200_ERR_INVALID;
  std::fill(out, out + 2 * count, 0.0);
  for (uint64_t j = 0; j < count; ++j) {
    if (indices[j] >> sv->n) return QB200_ERR_INVALID;
    const uint64_t p = to_physical(sv, indices[j]);
    if (Shard* s = local_shard(sv, (unsigned) (p >> sv->nl)))
      SV_TRY(qb200_get_ampl(s->ctx, sv->dtype, cur_buf(sv, *s), p & ((uint64_t{1} << sv->nl) - 1), out + 2 * j));
  }
  return sum_over_ranks(sv, out, 2 * count);  // one collective for all of them
}

int qb200_sv_set_ampl(qb200_sv* sv, uint64_t i, double re, double im) {
  if (!sv || (i >> sv->n)) return QB200_ERR_INVALID;
  sv->fresh = false;
  const uint64_t p = to_physical(sv, i);
  if (Shard* s = local_shard(sv, (unsigned) (p >> sv->nl)))
    SV_TRY(qb200_set_ampl(s->ctx, sv->dtype, cur_buf(sv, *s), p & ((uint64_t{1} << sv->nl) - 1), re, im));
  return QB200_OK;
}

int qb200_sv_bulk_set_ampl(qb200_sv* sv, uint64_t mask, uint64_t bits, double re, double im, int exclude) {
  if (!sv) return QB200_ERR_INVALID;
  sv->fresh = false;
  const uint64_t pm = to_physical(sv, mask), pb = to_physical(sv, bits);
  const uint64_t lmask = pm & ((uint64_t{1} << sv->nl) - 1), lbits = pb & ((uint64_t{1} << sv->nl) - 1);
  const uint64_t gmask = pm >> sv->nl, gbits = pb >> sv->nl;
  for (auto& s : sv->sh) {
    const bool gok = (s.rank & gmask) == gbits;
    if (gok) SV_TRY(qb200_bulk_set_ampl(s.ctx, sv->dtype, cur_buf(sv, s), sv->nl, lmask, lbits, re, im, exclude));
    else if (exclude) SV_TRY(qb200_bulk_set_ampl(s.ctx, sv->dtype, cur_buf(sv, s), sv->nl, 0, 0, re, im, 0));
  }
  return QB200_OK;
}

// ---- gates ----------------------------------------------------------------------------------------------------------
int qb200_sv_apply_controlled_gate(qb200_sv* sv, const unsigned* qs, unsigned nq, const unsigned* cqs, unsigned nc,
                                   uint64_t cvals, const void* matrix) {
  if (!sv || !matrix || (nq && !qs) || (nc && !cqs)) return QB200_ERR_INVALID;
  if (nq > kMaxTargets || (nc > 0 && nq > kMaxCtrlTargets)) return QB200_ERR_UNSUPPORTED;
  if (nq > sv->nl) return QB200_ERR_UNSUPPORTED;  // lib/simulator_custatevecex.h:67-71: a gate must fit a shard
  SV_TRY(make_local(sv, qs, nq));
  touch(sv, qs, nq);
  return local_gate(sv, qs, nq, cqs, nc, cvals, matrix, false, nullptr);
}

int qb200_sv_apply_gate(qb200_sv* sv, const unsigned* qs, unsigned nq, const void* matrix) {
  return qb200_sv_apply_controlled_gate(sv, qs, nq, nullptr, 0, 0, matrix);
}

int qb200_sv_expectation_value(qb200_sv* sv, const unsigned* qs, unsigned nq, const void* matrix, double out[2]) {
  if (!sv || !out || !matrix || !qs) return QB200_ERR_INVALID;
  out[0] = out[1] = 0;
  if (nq > kMaxTargets || nq > sv->nl) return QB200_ERR_UNSUPPORTED;
  SV_TRY(make_local(sv, qs, nq));
  double v[2] = {0, 0};
  SV_TRY(local_gate(sv, qs, nq, nullptr, 0, 0, matrix, true, v));
  SV_TRY(sum_over_ranks(sv, v, 2));
  out[0] = v[0];
  out[1] = v[1];
  return QB200_OK;
}

int qb200_sv_swap(qb200_sv* sv, const unsigned* victims, const unsigned* incoming, unsigned k) {
  if (!sv || !victims || !incoming) return QB200_ERR_INVALID;
  return exchange(sv, victims, incoming, k);
}

int qb200_sv_canonicalize(qb200_sv* sv) { return sv ? canonicalize(sv) : QB200_ERR_INVALID; }

static int gate_masks(unsigned num_qubits, unsigned num_global, const qb200_gate* gates, uint64_t count,
                      std::vector<uint64_t>* touch_m, std::vector<uint64_t>* need);

int qb200_sv_plan(unsigned num_qubits, unsigned num_global, const qb200_gate* gates, uint64_t count,
                  const unsigned* global_qubits, int reorder, int64_t* steps, uint64_t capacity, uint64_t* num_steps) {
  if (!gates && count) return QB200_ERR_INVALID;
  if (num_qubits > 63 || num_global > num_qubits) return QB200_ERR_INVALID;
  std::vector<uint64_t> touch_m, need;
  SV_TRY(gate_masks(num_qubits, num_global, gates, count, &touch_m, &need));
  uint64_t glob = 0;
  for (unsigned j = 0; j < num_global; ++j)
    glob |= uint64_t{1} << (global_qubits ? global_qubits[j] : num_qubits - num_global + j);
  if ((unsigned) __builtin_popcountll(glob) != num_global) return QB200_ERR_INVALID;
  SwapPlanner planner(num_qubits, num_global, touch_m, need, glob, reorder != 0);
  const auto plan = planner.Run();
  // encoding: a gate step is its op index (>= 0); a swap step is -(k) followed by k victims and k incoming qubits
  uint64_t w = 0;
  auto put = [&](int64_t v) {
    if (steps && w < capacity) steps[w] = v;
    ++w;
  };
  for (const auto& s : plan) {
    if (!s.is_swap) { put((int64_t) s.op); continue; }
    put(-(int64_t) s.victims.size());
    for (unsigned q : s.victims) put(q);
    for (unsigned q : s.incoming) put(q);
  }
  if (num_steps) *num_steps = w;
  return (steps && w > capacity) ? QB200_ERR_INVALID : QB200_OK;
}

// touch / need masks of a gate list (qb200_sv_plan's input); QB200_ERR_INVALID for a qubit out of range
static int gate_masks(unsigned num_qubits, unsigned num_global, const qb200_gate* gates, uint64_t count,
                      std::vector<uint64_t>* touch_m, std::vector<uint64_t>* need) {
  touch_m->assign(count, 0);
  need->assign(count, 0);
  for (uint64_t i = 0; i < count; ++i) {
    uint64_t t = 0, c = 0;
    for (unsigned j = 0; j < gates[i].num_targets; ++j) {
      if (gates[i].qs[j] >= num_qubits) return QB200_ERR_INVALID;
      t |= uint64_t{1} << gates[i].qs[j];
    }
    for (unsigned j = 0; j < gates[i].num_controls; ++j) {
      if (gates[i].cqs[j] >= num_qubits) return QB200_ERR_INVALID;
      c |= uint64_t{1} << gates[i].cqs[j];
    }
    if ((unsigned) __builtin_popcountll(t) > num_qubits - num_global) return QB200_ERR_UNSUPPORTED;
    (*touch_m)[i] = t | c;
    (*need)[i] = t;
  }
  return QB200_OK;
}

int qb200_sv_plan_initial(unsigned num_qubits, unsigned num_global, const qb200_gate* gates, uint64_t count, int reorder,
                          unsigned* global_qubits_out) {
  if ((!gates && count) || !global_qubits_out || num_qubits > 63 || num_global > num_qubits) return QB200_ERR_INVALID;
  std::vector<uint64_t> touch_m, need;
  SV_TRY(gate_masks(num_qubits, num_global, gates, count, &touch_m, &need));
  uint64_t def = 0;
  for (unsigned j = 0; j < num_global; ++j) def |= uint64_t{1} << (num_qubits - num_global + j);
  const uint64_t best = SwapPlanner::BestInitial(num_qubits, num_global, touch_m, need, def, reorder != 0);
  unsigned w = 0;
  for (unsigned q = 0; q < num_qubits; ++q)
    if ((best >> q) & 1) global_qubits_out[w++] = q;
  return QB200_OK;
}

// The state is fresh (|0...0>, zeros or uniform: the same under every qubit map): relabel the map so that the global
// qubits are the set whose schedule for THIS circuit exchanges the fewest shards.  No data moves.
static int choose_initial_map(qb200_sv* sv, const qb200_gate* gates, uint64_t count) {
  if (sv->g == 0) return QB200_OK;
  std::vector<uint64_t> touch_m, need;
  SV_TRY(gate_masks(sv->n, sv->g, gates, count, &touch_m, &need));
  std::vector<uint64_t> key;
  key.reserve(2 * count + 1);
  key.push_back((uint64_t) sv->reorder);
  for (uint64_t i = 0; i < count; ++i) { key.push_back(touch_m[i]); key.push_back(need[i]); }
  if (key != sv->init_key) {
    uint64_t cur = 0;
    for (unsigned t = 0; t < sv->g; ++t) cur |= uint64_t{1} << qubit_at(sv, sv->nl + t);
    sv->init_glob = SwapPlanner::BestInitial(sv->n, sv->g, touch_m, need, cur, sv->reorder != 0);
    sv->init_key = key;
  }
  unsigned il = 0, ig = 0;
  for (unsigned q = 0; q < sv->n; ++q) sv->pos[q] = ((sv->init_glob >> q) & 1) ? sv->nl + ig++ : il++;
  return QB200_OK;
}

int qb200_sv_run(qb200_sv* sv, const qb200_gate* gates, uint64_t count) {
  if (!sv || (!gates && count)) return QB200_ERR_INVALID;
  for (uint64_t i = 0; i < count; ++i) {
    if (gates[i].num_targets > kMaxTargets || (gates[i].num_controls && gates[i].num_targets > kMaxCtrlTargets))
      return QB200_ERR_UNSUPPORTED;
    if (gates[i].num_targets > sv->nl) return QB200_ERR_UNSUPPORTED;
  }
  if (sv->g == 0) {
    for (uint64_t i = 0; i < count; ++i)
      SV_TRY(local_gate(sv, gates[i].qs, gates[i].num_targets, gates[i].cqs, gates[i].num_controls, gates[i].cvals,
                        gates[i].matrix, false, nullptr));
    return QB200_OK;
  }
  if (sv->fresh && sv->free_initial_map && count > 0) SV_TRY(choose_initial_map(sv, gates, count));
  std::vector<unsigned> glob;
  for (unsigned t = 0; t < sv->g; ++t) glob.push_back(qubit_at(sv, sv->nl + t));
  // the schedule depends only on which qubits each gate touches and on the current global set: a circuit that
  // is run again (trajectories, benchmark steps) reuses it
  std::vector<uint64_t> key;
  key.reserve(2 * count + 2);
  key.push_back(sv->reorder);
  for (unsigned q : glob) key.push_back(q);
  for (uint64_t i = 0; i < count; ++i) {
    uint64_t t = 0, c = 0;
    for (unsigned j = 0; j < gates[i].num_targets; ++j) t |= uint64_t{1} << (gates[i].qs[j] & 63);
    for (unsigned j = 0; j < gates[i].num_controls; ++j) c |= uint64_t{1} << (gates[i].cqs[j] & 63);
    key.push_back(t);
    key.push_back(c);
  }
  if (key != sv->plan_key) {
    uint64_t need = 0;
    SV_TRY(qb200_sv_plan(sv->n, sv->g, gates, count, glob.data(), sv->reorder, nullptr, 0, &need));
    sv->plan_steps.assign(need, 0);
    sv->plan_key.clear();
    SV_TRY(qb200_sv_plan(sv->n, sv->g, gates, count, glob.data(), sv->reorder, sv->plan_steps.data(), need, &need));
    sv->plan_key = key;
  }
  const std::vector<int64_t>& steps = sv->plan_steps;
  const uint64_t need = steps.size();
  OverlapSpec spec;
  plan_overlap(sv, gates, steps, 0, &spec);
  for (uint64_t w = 0; w < need;) {
    if (spec.valid && w == spec.start) {
      PushPlan pp;
      // small tiles: the slim push kernel must not push the SMs into a larger shared-memory carveout (see push_launch)
      const unsigned tb = (unsigned) std::min<int>(kTileBits, std::max(8, sv->overlap_tile_bits));
      const int prc = push_prepare(sv, spec.victims, spec.incoming, &pp, tb);
      if (prc == QB200_OK) {
        SV_TRY(run_overlapped(sv, gates, steps, spec, pp));
        w = spec.swap + 1 + 2 * spec.victims.size();
        plan_overlap(sv, gates, steps, w, &spec);
        continue;
      }
      if (prc != QB200_ERR_UNSUPPORTED) return prc;
      spec.valid = false;   // no room for the spare buffers: plain gates, in-place exchange
    }
    const int64_t v = steps[w++];
    if (v >= 0) {
      const qb200_gate& gt = gates[v];
      touch(sv, gt.qs, gt.num_targets);
      SV_TRY(local_gate(sv, gt.qs, gt.num_targets, gt.cqs, gt.num_controls, gt.cvals, gt.matrix, false, nullptr));
    } else {
      const unsigned k = (unsigned) -v;
      unsigned vq[kMaxGlobal], iq[kMaxGlobal];
      for (unsigned j = 0; j < k; ++j) vq[j] = (unsigned) steps[w + j];
      for (unsigned j = 0; j < k; ++j) iq[j] = (unsigned) steps[w + k + j];
      w += 2 * k;
      SV_TRY(exchange(sv, vq, iq, k));
      plan_overlap(sv, gates, steps, w, &spec);
    }
  }
  return QB200_OK;
}

// ---- reductions ---------------------------------------------------------------------------------------------------
// out[r] for every rank r (zeros for the shards of other processes, then summed over the ranks)
static int per_shard_norms(qb200_sv* sv, uint64_t lmask, uint64_t lbits, uint64_t gmask, uint64_t gbits,
                           std::vector<double>* out) {
  out->assign(sv->P, 0.0);
  for (auto& s : sv->sh) SV_TRY(qb200_reduce_batch_begin(s.ctx, 1));
  int rc = QB200_OK;
  for (auto& s : sv->sh) {
    if ((s.rank & gmask) != gbits) continue;
    double dummy;
    rc = qb200_masked_norm(s.ctx, sv->dtype, cur_buf(sv, s), sv->nl, lmask, lbits, &dummy);
    if (rc) break;
  }
  for (auto& s : sv->sh) {
    double r[4] = {0, 0, 0, 0};
    uint32_t cnt = 0;
    int rc2 = qb200_reduce_batch_end(s.ctx, r, 2, &cnt);
    if (rc2 && !rc) rc = rc2;
    if (cnt) (*out)[s.rank] = r[0];
  }
  if (rc) return rc;
  return sum_over_ranks(sv, out->data(), sv->P);
}

int qb200_sv_norm(qb200_sv* sv, double* out) {
  if (!sv || !out) return QB200_ERR_INVALID;
  std::vector<double> v;
  SV_TRY(per_shard_norms(sv, 0, 0, 0, 0, &v));
  double t = 0;
  for (double x : v) t += x;
  *out = t;
  return QB200_OK;
}

static bool same_shape(const qb200_sv* a, const qb200_sv* b) {
  return a && b && a->n == b->n && a->P == b->P && a->dtype == b->dtype && a->mp == b->mp &&
         a->sh.size() == b->sh.size();
}

static int align_maps(qb200_sv* a, qb200_sv* b) {
  if (a->pos == b->pos) return QB200_OK;
  SV_TRY(canonicalize(a));
  return canonicalize(b);
}

int qb200_sv_inner_product(qb200_sv* a, qb200_sv* b, double out[2]) {
  if (!out || !same_shape(a, b)) return QB200_ERR_INVALID;
  SV_TRY(align_maps(a, b));
  for (auto& s : a->sh) SV_TRY(qb200_reduce_batch_begin(s.ctx, 1));
  int rc = QB200_OK;
  for (size_t i = 0; i < a->sh.size() && !rc; ++i) {
    // b's kernels run on b's streams: order a's read after them
    DevScope d(a->sh[i].device);
    cudaEventRecord(b->sh[i].ev, b->sh[i].stream);
    cudaStreamWaitEvent(a->sh[i].stream, b->sh[i].ev, 0);
    double dummy[2];
    rc = qb200_inner_product(a->sh[i].ctx, a->dtype, cur_buf(a, a->sh[i]), cur_buf(b, b->sh[i]), a->nl, dummy);
  }
  double v[2] = {0, 0};
  for (auto& s : a->sh) {
    double r[4] = {0, 0, 0, 0};
    uint32_t cnt = 0;
    int rc2 = qb200_reduce_batch_end(s.ctx, r, 2, &cnt);
    if (rc2 && !rc) rc = rc2;
    if (cnt) { v[0] += r[0]; v[1] += r[1]; }
  }
  if (rc) return rc;
  SV_TRY(sum_over_ranks(a, v, 2));
  out[0] = v[0];
  out[1] = v[1];
  return QB200_OK;
}

int qb200_sv_add(qb200_sv* src, qb200_sv* dest) {
  if (!same_shape(src, dest)) return QB200_ERR_INVALID;
  dest->fresh = false;
  SV_TRY(align_maps(src, dest));
  for (size_t i = 0; i < src->sh.size(); ++i) {
    DevScope d(dest->sh[i].device);
    cudaEventRecord(src->sh[i].ev, src->sh[i].stream);
    cudaStreamWaitEvent(dest->sh[i].stream, src->sh[i].ev, 0);
    SV_TRY(qb200_add(dest->sh[i].ctx, dest->dtype, cur_buf(src, src->sh[i]), cur_buf(dest, dest->sh[i]), dest->nl));
    // ... and src must not be overwritten before the add has read it
    cudaEventRecord(dest->sh[i].ev, dest->sh[i].stream);
    cudaStreamWaitEvent(src->sh[i].stream, dest->sh[i].ev, 0);
  }
  return QB200_OK;
}

int qb200_sv_copy(qb200_sv* src, qb200_sv* dest) {
  if (!same_shape(src, dest)) return QB200_ERR_INVALID;
  for (size_t i = 0; i < src->sh.size(); ++i) {
    DevScope d(dest->sh[i].device);
    cudaEventRecord(src->sh[i].ev, src->sh[i].stream);
    cudaStreamWaitEvent(dest->sh[i].stream, src->sh[i].ev, 0);
    if (cudaMemcpyAsync(cur_buf(dest, dest->sh[i]), cur_buf(src, src->sh[i]), shard_bytes(src), cudaMemcpyDeviceToDevice,
                        dest->sh[i].stream) != cudaSuccess) {
      (void) cudaGetLastError();
      return QB200_ERR_CUDA;
    }
    cudaEventRecord(dest->sh[i].ev, dest->sh[i].stream);
    cudaStreamWaitEvent(src->sh[i].stream, dest->sh[i].ev, 0);
  }
  dest->pos = src->pos;
  dest->fresh = src->fresh;
  return sync_all(dest);
}

int qb200_sv_multiply(qb200_sv* sv, double a) {
  if (!sv) return QB200_ERR_INVALID;
  sv->fresh = false;
  for (auto& s : sv->sh) SV_TRY(qb200_multiply(s.ctx, sv->dtype, a, cur_buf(sv, s), sv->nl));
  return QB200_OK;
}

// ---- sampling and measurement (canonical order: same cumulative sums as the unsharded state) --------------------
int qb200_sv_sample(qb200_sv* sv, const double* sorted_rs, uint64_t num_samples, uint64_t* out) {
  if (!sv || (num_samples && (!sorted_rs || !out))) return QB200_ERR_INVALID;
  if (num_samples == 0) return QB200_OK;
  SV_TRY(canonicalize(sv));
  std::vector<double> norms;
  SV_TRY(per_shard_norms(sv, 0, 0, 0, 0, &norms));
  std::vector<double> res(num_samples, 0.0);
  double lo = 0;
  uint64_t first = 0;
  for (unsigned r = 0; r < sv->P; ++r) {
    const double hi = lo + norms[r];
    uint64_t last = first;
    // the last shard also takes the draws that round-off left beyond the total (lib/statespace_basic.h:227-229)
    while (last < num_samples && (sorted_rs[last] < hi || r + 1 == sv->P)) ++last;
    if (last > first) {
      if (Shard* s = local_shard(sv, r)) {
        std::vector<double> rs(last - first);
        for (uint64_t i = first; i < last; ++i) rs[i - first] = sorted_rs[i] - lo;
        std::vector<uint64_t> idx(last - first);
        SV_TRY(qb200_sample(s->ctx, sv->dtype, cur_buf(sv, *s), sv->nl, rs.data(), last - first, idx.data()));
        for (uint64_t i = first; i < last; ++i) res[i] = (double) (idx[i - first] | (uint64_t{r} << sv->nl));
      }
    }
    first = last;
    lo = hi;
  }
  SV_TRY(sum_over_ranks(sv, res.data(), num_samples));  // indices < 2^53 are exact in double
  for (uint64_t i = 0; i < num_samples; ++i) out[i] = (uint64_t) res[i];
  return QB200_OK;
}

uint64_t qb200_sv_partial_norms_count(const qb200_sv* sv) {
  return sv ? qb200_partial_norms_count(sv->nl) * sv->P : 0;
}

int qb200_sv_partial_norms(qb200_sv* sv, double* out) {
  if (!sv || !out) return QB200_ERR_INVALID;
  SV_TRY(canonicalize(sv));
  const uint64_t per = qb200_partial_norms_count(sv->nl);
  std::fill(out, out + per * sv->P, 0.0);
  for (auto& s : sv->sh) SV_TRY(qb200_partial_norms(s.ctx, sv->dtype, cur_buf(sv, s), sv->nl, out + per * s.rank));
  return sum_over_ranks(sv, out, per * sv->P);
}

int qb200_sv_find_measured_bits(qb200_sv* sv, uint64_t m, double r, uint64_t mask, uint64_t* out_bits) {
  if (!sv || !out_bits) return QB200_ERR_INVALID;
  SV_TRY(canonicalize(sv));
  const uint64_t per = qb200_partial_norms_count(sv->nl);
  const unsigned rank = (unsigned) (m / per);
  if (rank >= sv->P) return QB200_ERR_INVALID;
  double v = 0;
  if (Shard* s = local_shard(sv, rank)) {
    uint64_t bits = 0;
    SV_TRY(qb200_find_measured_bits(s->ctx, sv->dtype, cur_buf(sv, *s), sv->nl, m % per, r,
                                    mask & ((uint64_t{1} << sv->nl) - 1), &bits));
    v = (double) (bits | ((uint64_t{rank} << sv->nl) & mask));
  }
  SV_TRY(sum_over_ranks(sv, &v, 1));
  *out_bits = (uint64_t) v;
  return QB200_OK;
}

int qb200_sv_collapse(qb200_sv* sv, uint64_t mask, uint64_t bits, double* out_norm) {
  if (!sv) return QB200_ERR_INVALID;
  sv->fresh = false;
  const uint64_t pm = to_physical(sv, mask), pb = to_physical(sv, bits);
  const uint64_t lmask = pm & ((uint64_t{1} << sv->nl) - 1), lbits = pb & ((uint64_t{1} << sv->nl) - 1);
  const uint64_t gmask = pm >> sv->nl, gbits = pb >> sv->nl;
  std::vector<double> norms;
  SV_TRY(per_shard_norms(sv, lmask, lbits, gmask, gbits, &norms));
  double t = 0;
  for (double x : norms) t += x;
  if (out_norm) *out_norm = t;
  const double renorm = 1.0 / std::sqrt(t);  // lib/statespace_cuda.h:319
  for (auto& s : sv->sh) {
    if ((s.rank & gmask) == gbits) SV_TRY(qb200_collapse_scaled(s.ctx, sv->dtype, cur_buf(sv, s), sv->nl, lmask, lbits, renorm));
    else SV_TRY(qb200_set_all_zeros(s.ctx, sv->dtype, cur_buf(sv, s), sv->nl));
  }
  return QB200_OK;
}

// ---- host transfers (canonical order; a multi-process state moves only this process's shard, at its offset) -----
int qb200_sv_copy_to_host(qb200_sv* sv, void* host) {
  if (!sv || !host) return QB200_ERR_INVALID;
  SV_TRY(canonicalize(sv));
  const size_t bytes = shard_bytes(sv);
  for (auto& s : sv->sh) {
    DevScope d(s.device);
    SV_CUDA(sv, cudaMemcpyAsync((char*) host + bytes * s.rank, cur_buf(sv, s), bytes, cudaMemcpyDeviceToHost, s.stream));
  }
  return sync_all(sv);
}

int qb200_sv_copy_from_host(qb200_sv* sv, const void* host) {
  if (!sv || !host) return QB200_ERR_INVALID;
  sv->fresh = false;
  for (unsigned q = 0; q < sv->n; ++q) sv->pos[q] = q;
  const size_t bytes = shard_bytes(sv);
  for (auto& s : sv->sh) {
    DevScope d(s.device);
    SV_CUDA(sv, cudaMemcpyAsync(cur_buf(sv, s), (const char*) host + bytes * s.rank, bytes, cudaMemcpyHostToDevice, s.stream));
  }
  return sync_all(sv);
}

}  // extern "C"
