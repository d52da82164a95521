// statespace.cu -- StateSpace entry points of the C ABI (replaces
// lib/statespace_cuda.h + lib/statespace_cuda_kernels.h).
//
// Every kernel is HBM-bound byte work: grid-stride loops over a fixed grid of
// kNumSMs * k blocks, 128-bit accesses (two fp32 amplitudes / one fp64
// amplitude per access), double accumulation with a fixed reduction tree, so
// results are run-to-run deterministic.
#include <algorithm>
#include <cmath>
#include <random>
#include <vector>

#include "gate_kernels.cuh"

namespace qb200 {

int finish_expectation(qb200_ctx* ctx, double* partials, uint32_t blocks, double out[2]);

constexpr int kNT = 256;
constexpr uint32_t kMapBlocks = kNumSMs * 16;
constexpr uint32_t kReduceBlocks = kNumSMs * 8;

template <typename FP>
inline bool aligned_ok(const void* p) {
  return (reinterpret_cast<uintptr_t>(p) & (2 * sizeof(FP) - 1)) == 0;
}
template <typename FP>
inline bool pair_ok(const void* p, unsigned n) {
  return n >= 1 && (reinterpret_cast<uintptr_t>(p) & 15) == 0;
}

inline uint32_t grid_for(uint64_t items, uint32_t max_blocks) {
  uint64_t b = (items + kNT - 1) / kNT;
  if (b < 1) b = 1;
  return (uint32_t) std::min<uint64_t>(b, max_blocks);
}

// ---------------------------------------------------------------------------
// element-wise maps.  F: void(uint64_t amp_index, FP& re, FP& im)
// PAIR: a thread moves amplitudes 2j and 2j+1 together.
// ---------------------------------------------------------------------------
template <typename FP, bool PAIR, bool LOAD, typename F>
__global__ void __launch_bounds__(kNT) k_map(FP* __restrict__ st, uint64_t items, F f) {
  for (uint64_t j = blockIdx.x * uint64_t{kNT} + threadIdx.x; j < items;
       j += uint64_t{gridDim.x} * kNT) {
    if constexpr (PAIR) {
      FP a = 0, b = 0, c = 0, d = 0;
      if constexpr (LOAD) ld2(st + 4 * j, a, b, c, d);
      f(2 * j, a, b);
      f(2 * j + 1, c, d);
      st2(st + 4 * j, a, b, c, d);
    } else {
      FP a = 0, b = 0;
      if constexpr (LOAD) ld1(st + 2 * j, a, b);
      f(j, a, b);
      st1(st + 2 * j, a, b);
    }
  }
}

template <typename FP, bool LOAD, typename F>
int launch_map(qb200_ctx* ctx, FP* st, unsigned n, F f) {
  if (!ctx || !st || n > kMaxQubits || !aligned_ok<FP>(st)) return QB200_ERR_INVALID;
  DeviceGuard guard(ctx);
  note_state_written();
  const uint64_t amps = uint64_t{1} << n;
  if (pair_ok<FP>(st, n)) {
    const uint64_t items = amps / 2;
    k_map<FP, true, LOAD, F><<<grid_for(items, kMapBlocks), kNT, 0, ctx->stream>>>(st, items, f);
  } else {
    k_map<FP, false, LOAD, F><<<grid_for(amps, kMapBlocks), kNT, 0, ctx->stream>>>(st, amps, f);
  }
  QB_LAUNCHED(ctx);
  return QB200_OK;
}

template <typename FP> struct FUniform {
  FP v;
  __device__ void operator()(uint64_t, FP& re, FP& im) const { re = v; im = 0; }
};
template <typename FP> struct FMultiply {
  FP a;
  __device__ void operator()(uint64_t, FP& re, FP& im) const { re *= a; im *= a; }
};
template <typename FP> struct FAdd {
  const FP* src;
  __device__ void operator()(uint64_t i, FP& re, FP& im) const {
    FP a, b;
    ld1(src + 2 * i, a, b);
    re += a; im += b;
  }
};
template <typename FP> struct FBulkSet {
  uint64_t mask, bits;
  FP vre, vim;
  bool exclude;
  __device__ void operator()(uint64_t i, FP& re, FP& im) const {
    const bool in_mask = ((i & mask) == bits) != exclude;
    re = in_mask ? vre : re;
    im = in_mask ? vim : im;
  }
};
template <typename FP> struct FCollapse {
  uint64_t mask, bits;
  FP renorm;
  __device__ void operator()(uint64_t i, FP& re, FP& im) const {
    const bool keep = (i & mask) == bits;
    re = keep ? re * renorm : FP(0);
    im = keep ? im * renorm : FP(0);
  }
};

template <typename FP>
__global__ void k_set_ampl(FP* st, uint64_t i, FP re, FP im) {
  st[2 * i] = re;
  st[2 * i + 1] = im;
}

// ---------------------------------------------------------------------------
// reductions.  OP 0: sum |s1|^2 where (i & mask) == bits (norm, masked norm)
//              OP 1: sum conj(s1) s2 (complex; RealInnerProduct takes .re)
// Products in FP like the reference functors (lib/util_cuda.h:94-125),
// accumulation in double.
// ---------------------------------------------------------------------------
template <typename FP, int OP>
__device__ __forceinline__ void red_term(uint64_t i, uint64_t mask, uint64_t bits, FP a, FP b,
                                         FP c, FP d, double& re, double& im) {
  if constexpr (OP == 0) {
    const FP t = a * a + b * b;
    if ((i & mask) == bits) re += t;
  } else {
    re += a * c + b * d;
    im += a * d - b * c;
  }
}

template <typename FP, int OP, bool PAIR>
__global__ void __launch_bounds__(kNT)
k_reduce(const FP* __restrict__ s1, const FP* __restrict__ s2, uint64_t items, uint64_t mask,
         uint64_t bits, double* __restrict__ partials) {
  double re = 0, im = 0;
  for (uint64_t j = blockIdx.x * uint64_t{kNT} + threadIdx.x; j < items;
       j += uint64_t{gridDim.x} * kNT) {
    if constexpr (PAIR) {
      FP a0, b0, a1, b1, c0 = 0, d0 = 0, c1 = 0, d1 = 0;
      ld2(s1 + 4 * j, a0, b0, a1, b1);
      if constexpr (OP == 1) ld2(s2 + 4 * j, c0, d0, c1, d1);
      red_term<FP, OP>(2 * j, mask, bits, a0, b0, c0, d0, re, im);
      red_term<FP, OP>(2 * j + 1, mask, bits, a1, b1, c1, d1, re, im);
    } else {
      FP a, b, c = 0, d = 0;
      ld1(s1 + 2 * j, a, b);
      if constexpr (OP == 1) ld1(s2 + 2 * j, c, d);
      red_term<FP, OP>(j, mask, bits, a, b, c, d, re, im);
    }
  }
  block_sum2<kNT>(re, im);
  if (threadIdx.x == 0) {
    partials[2 * blockIdx.x] = re;
    partials[2 * blockIdx.x + 1] = im;
  }
}

template <typename FP, int OP>
int reduce(qb200_ctx* ctx, const FP* s1, const FP* s2, unsigned n, uint64_t mask, uint64_t bits,
           double out[2]) {
  if (!ctx || !s1 || !s2 || n > kMaxQubits || !aligned_ok<FP>(s1) || !aligned_ok<FP>(s2))
    return QB200_ERR_INVALID;
  DeviceGuard guard(ctx);
  const uint64_t amps = uint64_t{1} << n;
  const bool pair = pair_ok<FP>(s1, n) && pair_ok<FP>(s2, n);
  const uint64_t items = pair ? amps / 2 : amps;
  const uint32_t blocks = grid_for(items, kReduceBlocks);
  int rc = ensure_scratch(ctx, (2 * size_t{blocks} + 2) * sizeof(double));
  if (rc) return rc;
  double* partials = (double*) ctx->scratch;
  if (pair)
    k_reduce<FP, OP, true><<<blocks, kNT, 0, ctx->stream>>>(s1, s2, items, mask, bits, partials);
  else
    k_reduce<FP, OP, false><<<blocks, kNT, 0, ctx->stream>>>(s1, s2, items, mask, bits, partials);
  QB_LAUNCHED(ctx);
  return finish_expectation(ctx, partials, blocks, out);
}

// ---------------------------------------------------------------------------
// chunk norms / prefix / locate (PartialNorms, FindMeasuredBits, Sample)
// ---------------------------------------------------------------------------
inline unsigned chunk_bits(unsigned n) { return n < kChunkBits ? n : kChunkBits; }

// one block per chunk; thread t sums the contiguous slice [t*per, (t+1)*per)
template <typename FP>
__device__ __forceinline__ double slice_sum(const FP* __restrict__ chunk, uint64_t per, uint64_t t) {
  double s = 0;
  const FP* p = chunk + 2 * t * per;
  for (uint64_t k = 0; k < per; ++k) {
    FP a, b;
    ld1(p + 2 * k, a, b);
    s += (double) a * a + (double) b * b;
  }
  return s;
}

// out[m] = sum |a|^2 over chunk m.  One block per chunk, 16-byte loads in address order (a warp reads 512 contiguous
// bytes per instruction): the pass runs at the read roofline.  (Until round 2 thread t summed the contiguous slice
// [t*per, (t+1)*per) with 8-byte loads 256 bytes apart: 2.5 read-pass equivalents.)  The summation order differs from
// k_locate's slice-by-slice scan; a draw within round-off of a chunk boundary falls back to the chunk's last index there.
template <typename FP>
__global__ void __launch_bounds__(kNT)
k_chunk_norms(const FP* __restrict__ st, uint64_t chunk_amps, uint64_t nchunks,
              double* __restrict__ out) {
  constexpr int APT = sizeof(FP) == 4 ? 2 : 1;   // amplitudes per 16-byte item
  const uint64_t items = chunk_amps / APT;
  for (uint64_t m = blockIdx.x; m < nchunks; m += gridDim.x) {
    double s = 0, z = 0;
    if (items >= 1 && (reinterpret_cast<uintptr_t>(st) & 15) == 0) {
      const uint4* p = reinterpret_cast<const uint4*>(st + 2 * m * chunk_amps);
      for (uint64_t i = threadIdx.x; i < items; i += kNT) {
        const uint4 v = __ldg(p + i);
        if constexpr (APT == 2) {
          const float a = __uint_as_float(v.x), b = __uint_as_float(v.y), c = __uint_as_float(v.z), d = __uint_as_float(v.w);
          s += (double) a * a + (double) b * b;
          s += (double) c * c + (double) d * d;
        } else {
          const double a = __hiloint2double(v.y, v.x), b = __hiloint2double(v.w, v.z);
          s += a * a + b * b;
        }
      }
    } else {
      const uint64_t per = chunk_amps >= kNT ? chunk_amps / kNT : 1;
      if (threadIdx.x < chunk_amps) s = slice_sum(st + 2 * m * chunk_amps, per, threadIdx.x);
    }
    block_sum2<kNT>(s, z);
    if (threadIdx.x == 0) out[m] = s;
  }
}

// exclusive prefix over `count` doubles with one block: prefix[0]=0 ... prefix[count]=total
__global__ void __launch_bounds__(1024)
k_exclusive_scan(const double* __restrict__ in, uint64_t count, double* __restrict__ prefix) {
  __shared__ double seg[1024];
  const uint64_t per = (count + 1023) / 1024;
  const uint64_t lo = threadIdx.x * per < count ? threadIdx.x * per : count;
  const uint64_t hi = lo + per < count ? lo + per : count;
  double s = 0;
  for (uint64_t k = lo; k < hi; ++k) s += in[k];
  seg[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double run = 0;
    for (int t = 0; t < 1024; ++t) {
      const double v = seg[t];
      seg[t] = run;
      run += v;
    }
    prefix[count] = run;
  }
  __syncthreads();
  double run = seg[threadIdx.x];
  for (uint64_t k = lo; k < hi; ++k) {
    prefix[k] = run;
    run += in[k];
  }
}

// Finds, for each query, the first amplitude index whose running sum of |a|^2
// exceeds the query value.  SAMPLE: queries are sorted_rs (global cumulative
// values), the chunk is found by binary search over the chunk prefix.
// !SAMPLE: one query (chunk m, value r relative to the chunk start).
template <typename FP, bool SAMPLE>
__global__ void __launch_bounds__(kNT)
k_locate(const FP* __restrict__ st, unsigned n, uint64_t chunk_amps, uint64_t nchunks,
         const double* __restrict__ prefix, const double* __restrict__ rs, uint64_t nqueries,
         uint64_t m_in, double r_in, uint64_t* __restrict__ out) {
  __shared__ double excl[kNT];
  __shared__ unsigned long long best;
  const uint64_t per = chunk_amps >= kNT ? chunk_amps / kNT : 1;
  const uint64_t last = (uint64_t{1} << n) - 1;

  for (uint64_t q = blockIdx.x; q < nqueries; q += gridDim.x) {
    uint64_t m = m_in;
    double r = r_in;
    if constexpr (SAMPLE) {
      r = rs[q];
      if (!(r < prefix[nchunks])) {  // round-off tail (lib/statespace_basic.h:227-229)
        if (threadIdx.x == 0) out[q] = last;
        continue;
      }
      uint64_t lo = 0, hi = nchunks;  // largest m with prefix[m] <= r
      while (hi - lo > 1) {
        const uint64_t mid = (lo + hi) / 2;
        if (prefix[mid] <= r) lo = mid; else hi = mid;
      }
      m = lo;
      r -= prefix[m];
    }
    const FP* chunk = st + 2 * m * chunk_amps;
    const bool active = threadIdx.x < chunk_amps;
    const double s = active ? slice_sum(chunk, per, threadIdx.x) : 0.0;
    excl[threadIdx.x] = s;
    if (threadIdx.x == 0) best = ~0ull;
    __syncthreads();
    if (threadIdx.x == 0) {
      double run = 0;
      for (int t = 0; t < kNT; ++t) {
        const double v = excl[t];
        excl[t] = run;
        run += v;
      }
    }
    __syncthreads();
    if (active && r < excl[threadIdx.x] + s) {
      double c = excl[threadIdx.x];
      const FP* p = chunk + 2 * threadIdx.x * per;
      for (uint64_t k = 0; k < per; ++k) {
        FP a, b;
        ld1(p + 2 * k, a, b);
        c += (double) a * a + (double) b * b;
        if (r < c) {
          atomicMin(&best, (unsigned long long) (threadIdx.x * per + k));
          break;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      // no hit: "return the last bitstring in the unlikely case of underflow"
      // (lib/statespace_basic.h:292-293)
      const uint64_t k = best == ~0ull ? chunk_amps - 1 : (uint64_t) best;
      out[q] = m * chunk_amps + k;
    }
    __syncthreads();
  }
}

template <typename FP>
int chunk_norms_device(qb200_ctx* ctx, const FP* st, unsigned n, double* d_out) {
  const uint64_t chunk_amps = uint64_t{1} << chunk_bits(n);
  const uint64_t nchunks = (uint64_t{1} << n) / chunk_amps;
  const uint32_t blocks = (uint32_t) std::min<uint64_t>(nchunks, 1u << 20);
  k_chunk_norms<FP><<<blocks, kNT, 0, ctx->stream>>>(st, chunk_amps, nchunks, d_out);
  QB_LAUNCHED(ctx);
  return QB200_OK;
}

template <typename FP>
int partial_norms(qb200_ctx* ctx, const FP* st, unsigned n, double* out) {
  if (!ctx || !st || !out || n > kMaxQubits || !aligned_ok<FP>(st)) return QB200_ERR_INVALID;
  DeviceGuard guard(ctx);
  const uint64_t nchunks = qb200_partial_norms_count(n);
  int rc = ensure_scratch(ctx, nchunks * sizeof(double));
  if (rc) return rc;
  rc = chunk_norms_device(ctx, st, n, (double*) ctx->scratch);
  if (rc) return rc;
  QB_CUDA(ctx, cudaMemcpyAsync(out, ctx->scratch, nchunks * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return QB200_OK;
}

template <typename FP>
int find_measured_bits(qb200_ctx* ctx, const FP* st, unsigned n, uint64_t m, double r,
                       uint64_t mask, uint64_t* out_bits) {
  if (!ctx || !st || !out_bits || n > kMaxQubits || !aligned_ok<FP>(st)) return QB200_ERR_INVALID;
  if (m >= qb200_partial_norms_count(n)) return QB200_ERR_INVALID;
  DeviceGuard guard(ctx);
  int rc = ensure_scratch(ctx, sizeof(uint64_t));
  if (rc) return rc;
  rc = ensure_pinned(ctx, sizeof(uint64_t));
  if (rc) return rc;
  const uint64_t chunk_amps = uint64_t{1} << chunk_bits(n);
  k_locate<FP, false><<<1, kNT, 0, ctx->stream>>>(st, n, chunk_amps, 0, nullptr, nullptr, 1, m, r,
                                                  (uint64_t*) ctx->scratch);
  QB_LAUNCHED(ctx);
  QB_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, ctx->scratch, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *out_bits = *(uint64_t*) ctx->pinned & mask;
  return QB200_OK;
}

// sample_rng.cu
size_t sorted_uniform_temp_bytes(uint64_t ns);
int sorted_uniform_device(qb200_ctx* ctx, unsigned seed, uint64_t ns, double max_value, const double* d_max_value,
                          double* d_draws, double* d_sorted, void* d_temp, size_t temp_bytes);

// sorted_rs != nullptr: the caller's sorted values (host memory); else they are drawn on the device from
// (seed, max_value) exactly as GenerateRandomValues<double> draws them (sample_rng.cu).
template <typename FP>
int sample(qb200_ctx* ctx, const FP* st, unsigned n, const double* sorted_rs, unsigned seed, double max_value,
           uint64_t ns, uint64_t* out) {
  if (!ctx || !st || n > kMaxQubits || !aligned_ok<FP>(st)) return QB200_ERR_INVALID;
  if (ns == 0) return QB200_OK;
  if (!out) return QB200_ERR_INVALID;
  DeviceGuard guard(ctx);
  const uint64_t chunk_amps = uint64_t{1} << chunk_bits(n);
  const uint64_t nchunks = (uint64_t{1} << n) / chunk_amps;
  const size_t temp_bytes = sorted_rs ? 0 : sorted_uniform_temp_bytes(ns);
  const size_t bytes = (2 * nchunks + 1 + 2 * ns + (sorted_rs ? 0 : ns)) * sizeof(double) + 256 + temp_bytes;
  int rc = ensure_scratch(ctx, bytes);
  if (rc) return rc;
  double* d_sums = (double*) ctx->scratch;
  double* d_prefix = d_sums + nchunks;
  double* d_rs = d_prefix + nchunks + 1;
  uint64_t* d_out = (uint64_t*) (d_rs + ns);
  rc = chunk_norms_device(ctx, st, n, d_sums);
  if (rc) return rc;
  k_exclusive_scan<<<1, 1024, 0, ctx->stream>>>(d_sums, nchunks, d_prefix);
  QB_LAUNCHED(ctx);
  if (sorted_rs) {
    QB_CUDA(ctx, cudaMemcpyAsync(d_rs, sorted_rs, ns * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  } else {
    // max_value < 0: draw in [0, total of the chunk sums) -- the state's norm without a separate Norm pass
    double* d_draws = (double*) (d_out + ns);
    void* d_temp = (void*) (((uintptr_t) (d_draws + ns) + 255) & ~uintptr_t{255});
    rc = sorted_uniform_device(ctx, seed, ns, max_value, max_value < 0 ? d_prefix + nchunks : nullptr, d_draws, d_rs,
                               d_temp, temp_bytes);
    if (rc) return rc;
  }
  const uint32_t blocks = (uint32_t) std::min<uint64_t>(ns, 1u << 20);
  k_locate<FP, true><<<blocks, kNT, 0, ctx->stream>>>(st, n, chunk_amps, nchunks, d_prefix, d_rs, ns,
                                                      0, 0.0, d_out);
  QB_LAUNCHED(ctx);
  QB_CUDA(ctx, cudaMemcpyAsync(out, d_out, ns * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return QB200_OK;
}

template <typename FP>
int collapse(qb200_ctx* ctx, FP* st, unsigned n, uint64_t mask, uint64_t bits, double* out_norm) {
  // the masked norm is needed on the host NOW: not available inside qb200_reduce_batch_begin/end, where every
  // reduction of the context is deferred (the state would be scaled by NaN)
  if (ctx && ctx->batching) return QB200_ERR_INVALID;
  double r[2];
  int rc = reduce<FP, 0>(ctx, st, st, n, mask, bits, r);
  if (rc) return rc;
  if (out_norm) *out_norm = r[0];
  // fp_type renorm = 1 / std::sqrt(r)  (lib/statespace_cuda.h:319)
  const FP renorm = (FP) (1.0 / std::sqrt(r[0]));
  return launch_map<FP, true>(ctx, st, n, FCollapse<FP>{mask, bits, renorm});
}

template <typename FP>
int get_ampl(qb200_ctx* ctx, const FP* st, uint64_t i, double out[2]) {
  if (!ctx || !st || !out) return QB200_ERR_INVALID;
  DeviceGuard guard(ctx);
  int rc = ensure_pinned(ctx, 2 * sizeof(FP));
  if (rc) return rc;
  QB_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, st + 2 * i, 2 * sizeof(FP), cudaMemcpyDeviceToHost, ctx->stream));
  QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  out[0] = ((FP*) ctx->pinned)[0];
  out[1] = ((FP*) ctx->pinned)[1];
  return QB200_OK;
}

template <typename FP>
int set_ampl(qb200_ctx* ctx, FP* st, uint64_t i, double re, double im) {
  if (!ctx || !st) return QB200_ERR_INVALID;
  DeviceGuard guard(ctx);
  k_set_ampl<FP><<<1, 1, 0, ctx->stream>>>(st, i, (FP) re, (FP) im);
  QB_LAUNCHED(ctx);
  return QB200_OK;
}

template <typename FP>
int set_all_zeros(qb200_ctx* ctx, FP* st, unsigned n) {
  if (!ctx || !st || n > kMaxQubits) return QB200_ERR_INVALID;
  DeviceGuard guard(ctx);
  QB_CUDA(ctx, cudaMemsetAsync(st, 0, (size_t{2} << n) * sizeof(FP), ctx->stream));
  return QB200_OK;
}

}  // namespace qb200

using namespace qb200;

#define QB_DISPATCH(dtype, EXPR_F32, EXPR_F64)        \
  do {                                                \
    if ((dtype) == QB200_F32) { using FP = float; return EXPR_F32; }  \
    if ((dtype) == QB200_F64) { using FP = double; return EXPR_F64; } \
    return QB200_ERR_INVALID;                         \
  } while (0)

extern "C" {

int qb200_set_all_zeros(qb200_ctx* ctx, int dtype, void* state, unsigned n) {
  note_state_written();
  QB_DISPATCH(dtype, set_all_zeros<FP>(ctx, (FP*) state, n), set_all_zeros<FP>(ctx, (FP*) state, n));
}

int qb200_set_state_zero(qb200_ctx* ctx, int dtype, void* state, unsigned n) {
  int rc = qb200_set_all_zeros(ctx, dtype, state, n);
  if (rc) return rc;
  return qb200_set_ampl(ctx, dtype, state, 0, 1.0, 0.0);
}

int qb200_set_state_uniform(qb200_ctx* ctx, int dtype, void* state, unsigned n) {
  if (n > kMaxQubits) return QB200_ERR_INVALID;
  // fp_type v = double{1} / std::sqrt(hsize)  (lib/statespace_cuda.h:122)
  const double v = 1.0 / std::sqrt((double) (uint64_t{1} << n));
  QB_DISPATCH(dtype, (launch_map<FP, false>(ctx, (FP*) state, n, FUniform<FP>{(FP) v})),
              (launch_map<FP, false>(ctx, (FP*) state, n, FUniform<FP>{(FP) v})));
}

int qb200_get_ampl(qb200_ctx* ctx, int dtype, const void* state, uint64_t i, double out[2]) {
  QB_DISPATCH(dtype, get_ampl<FP>(ctx, (const FP*) state, i, out), get_ampl<FP>(ctx, (const FP*) state, i, out));
}

int qb200_set_ampl(qb200_ctx* ctx, int dtype, void* state, uint64_t i, double re, double im) {
  note_state_written();
  QB_DISPATCH(dtype, set_ampl<FP>(ctx, (FP*) state, i, re, im), set_ampl<FP>(ctx, (FP*) state, i, re, im));
}

int qb200_bulk_set_ampl(qb200_ctx* ctx, int dtype, void* state, unsigned n, uint64_t mask,
                        uint64_t bits, double re, double im, int exclude) {
  QB_DISPATCH(dtype,
              (launch_map<FP, true>(ctx, (FP*) state, n, FBulkSet<FP>{mask, bits, (FP) re, (FP) im, exclude != 0})),
              (launch_map<FP, true>(ctx, (FP*) state, n, FBulkSet<FP>{mask, bits, (FP) re, (FP) im, exclude != 0})));
}

int qb200_add(qb200_ctx* ctx, int dtype, const void* src, void* dest, unsigned n) {
  if (!src) return QB200_ERR_INVALID;
  QB_DISPATCH(dtype, (launch_map<FP, true>(ctx, (FP*) dest, n, FAdd<FP>{(const FP*) src})),
              (launch_map<FP, true>(ctx, (FP*) dest, n, FAdd<FP>{(const FP*) src})));
}

int qb200_multiply(qb200_ctx* ctx, int dtype, double a, void* state, unsigned n) {
  QB_DISPATCH(dtype, (launch_map<FP, true>(ctx, (FP*) state, n, FMultiply<FP>{(FP) a})),
              (launch_map<FP, true>(ctx, (FP*) state, n, FMultiply<FP>{(FP) a})));
}

int qb200_inner_product(qb200_ctx* ctx, int dtype, const void* s1, const void* s2, unsigned n,
                        double out[2]) {
  if (!out) return QB200_ERR_INVALID;
  QB_DISPATCH(dtype, (reduce<FP, 1>(ctx, (const FP*) s1, (const FP*) s2, n, 0, 0, out)),
              (reduce<FP, 1>(ctx, (const FP*) s1, (const FP*) s2, n, 0, 0, out)));
}

int qb200_real_inner_product(qb200_ctx* ctx, int dtype, const void* s1, const void* s2,
                             unsigned n, double* out) {
  if (!out) return QB200_ERR_INVALID;
  double r[2] = {0, 0};
  int rc = qb200_inner_product(ctx, dtype, s1, s2, n, r);
  *out = r[0];
  return rc;
}

int qb200_norm(qb200_ctx* ctx, int dtype, const void* state, unsigned n, double* out) {
  if (!out) return QB200_ERR_INVALID;
  double r[2] = {0, 0};
  int rc;
  if (dtype == QB200_F32) rc = reduce<float, 0>(ctx, (const float*) state, (const float*) state, n, 0, 0, r);
  else if (dtype == QB200_F64) rc = reduce<double, 0>(ctx, (const double*) state, (const double*) state, n, 0, 0, r);
  else rc = QB200_ERR_INVALID;
  *out = r[0];
  return rc;
}

int qb200_sample(qb200_ctx* ctx, int dtype, const void* state, unsigned n, const double* sorted_rs,
                 uint64_t num_samples, uint64_t* out) {
  if (num_samples && !sorted_rs) return QB200_ERR_INVALID;
  QB_DISPATCH(dtype, sample<FP>(ctx, (const FP*) state, n, sorted_rs, 0, 0.0, num_samples, out),
              sample<FP>(ctx, (const FP*) state, n, sorted_rs, 0, 0.0, num_samples, out));
}

int qb200_sample_seeded(qb200_ctx* ctx, int dtype, const void* state, unsigned n, uint64_t num_samples, unsigned seed,
                        double norm, uint64_t* out) {
  QB_DISPATCH(dtype, sample<FP>(ctx, (const FP*) state, n, nullptr, seed, norm, num_samples, out),
              sample<FP>(ctx, (const FP*) state, n, nullptr, seed, norm, num_samples, out));
}

int qb200_generate_random_values_device(qb200_ctx* ctx, uint64_t num_samples, unsigned seed, double max_value, double* out) {
  if (!ctx || (num_samples && !out)) return QB200_ERR_INVALID;
  if (num_samples == 0) return QB200_OK;
  DeviceGuard guard(ctx);
  const size_t temp_bytes = sorted_uniform_temp_bytes(num_samples);
  int rc = ensure_scratch(ctx, 2 * num_samples * sizeof(double) + 256 + temp_bytes);
  if (rc) return rc;
  double* d_draws = (double*) ctx->scratch;
  double* d_sorted = d_draws + num_samples;
  void* d_temp = (void*) (((uintptr_t) (d_sorted + num_samples) + 255) & ~uintptr_t{255});
  rc = sorted_uniform_device(ctx, seed, num_samples, max_value, nullptr, d_draws, d_sorted, d_temp, temp_bytes);
  if (rc) return rc;
  QB_CUDA(ctx, cudaMemcpyAsync(out, d_sorted, num_samples * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return QB200_OK;
}

int qb200_generate_random_values(uint64_t num_samples, unsigned seed, double max_value, double* out) {
  if (num_samples && !out) return QB200_ERR_INVALID;
  std::mt19937 rgen(seed);
  std::uniform_real_distribution<double> distr(0.0, max_value);
  for (uint64_t i = 0; i < num_samples; ++i) out[i] = distr(rgen);
  std::sort(out, out + num_samples);
  return QB200_OK;
}

uint64_t qb200_partial_norms_count(unsigned n) {
  return (uint64_t{1} << n) >> chunk_bits(n);
}

int qb200_partial_norms(qb200_ctx* ctx, int dtype, const void* state, unsigned n, double* out) {
  QB_DISPATCH(dtype, partial_norms<FP>(ctx, (const FP*) state, n, out),
              partial_norms<FP>(ctx, (const FP*) state, n, out));
}

int qb200_find_measured_bits(qb200_ctx* ctx, int dtype, const void* state, unsigned n, uint64_t m,
                             double r, uint64_t mask, uint64_t* out_bits) {
  QB_DISPATCH(dtype, find_measured_bits<FP>(ctx, (const FP*) state, n, m, r, mask, out_bits),
              find_measured_bits<FP>(ctx, (const FP*) state, n, m, r, mask, out_bits));
}

int qb200_collapse(qb200_ctx* ctx, int dtype, void* state, unsigned n, uint64_t mask, uint64_t bits,
                   double* out_norm) {
  QB_DISPATCH(dtype, collapse<FP>(ctx, (FP*) state, n, mask, bits, out_norm),
              collapse<FP>(ctx, (FP*) state, n, mask, bits, out_norm));
}

int qb200_masked_norm(qb200_ctx* ctx, int dtype, const void* state, unsigned n, uint64_t mask, uint64_t bits,
                      double* out) {
  if (!out) return QB200_ERR_INVALID;
  double r[2] = {0, 0};
  int rc;
  if (dtype == QB200_F32) rc = reduce<float, 0>(ctx, (const float*) state, (const float*) state, n, mask, bits, r);
  else if (dtype == QB200_F64) rc = reduce<double, 0>(ctx, (const double*) state, (const double*) state, n, mask, bits, r);
  else rc = QB200_ERR_INVALID;
  *out = r[0];
  return rc;
}

int qb200_collapse_scaled(qb200_ctx* ctx, int dtype, void* state, unsigned n, uint64_t mask, uint64_t bits,
                          double renorm) {
  QB_DISPATCH(dtype, (launch_map<FP, true>(ctx, (FP*) state, n, FCollapse<FP>{mask, bits, (FP) renorm})),
              (launch_map<FP, true>(ctx, (FP*) state, n, FCollapse<FP>{mask, bits, (FP) renorm})));
}

int qb200_internal_to_normal_order(qb200_ctx* ctx, int, void*, unsigned) {
  return ctx ? QB200_OK : QB200_ERR_INVALID;
}
int qb200_normal_to_internal_order(qb200_ctx* ctx, int, void*, unsigned) {
  return ctx ? QB200_OK : QB200_ERR_INVALID;
}

}  // extern "C"
