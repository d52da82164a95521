// sample_rng.cu -- the sorted random values of StateSpace::Sample, drawn ON THE DEVICE.
//
// The reference draws them on the host and copies them over ("TODO: generate random values on the device",
// lib/statespace_cuda.h:292): GenerateRandomValues<double> (lib/util.h:67-85) = std::mt19937(seed) +
// std::uniform_real_distribution<double>(0, norm), then std::sort.  At 10^6 samples that is ~90 ms of host work
// around a 40 ms device scan of a 34-qubit state.  Here the SAME values come from a device Mersenne Twister:
//   * MT19937 exactly (seeding recurrence, 624-word twist, tempering): one CTA keeps the generator state in shared
//     memory; a twist has three phases whose elements are independent (words 0..226 depend on old words only,
//     227..453 on words of phase one, 454..623 on words of phase two), old and new state in two arrays so that a
//     phase is read -> write with one barrier;
//   * std::uniform_real_distribution<double> exactly as libstdc++ evaluates it (bits/random.tcc generate_canonical):
//     two 32-bit draws x0, x1 -> (double(x0) + double(x1) * 2^32) / 2^64, clamped below 1, times (b - a), plus a;
//   * an ascending radix sort (cub::DeviceRadixSort, keys only; not on the state's data path).
// Same seed => bit-identical sorted values => the very same sample indices as the host path (tests/test_statespace_gpu.py).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace qb200 {

namespace {

constexpr int kMtN = 624, kMtM = 397;
constexpr int kMtThreads = 256;

__device__ __forceinline__ uint32_t mt_twist(uint32_t cur, uint32_t next, uint32_t far) {
  const uint32_t y = (cur & 0x80000000u) | (next & 0x7fffffffu);
  return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

// rs[j] = uniform_real_distribution<double>(0, max_value)(mt19937(seed)), j = 0 .. ns-1, in draw order
__global__ void __launch_bounds__(kMtThreads)
k_mt19937_uniform(uint32_t seed, uint64_t ns, double max_value, const double* __restrict__ d_max_value,
                  double* __restrict__ rs) {
  __shared__ uint32_t st[2][kMtN];
  const int t = threadIdx.x;
  if (d_max_value != nullptr) max_value = *d_max_value;   // computed earlier on this stream (the sampler's chunk prefix)
  if (t == 0) {
    uint32_t x = seed;
    st[0][0] = x;
    for (int i = 1; i < kMtN; ++i) {
      x = 1812433253u * (x ^ (x >> 30)) + (uint32_t) i;
      st[0][i] = x;
    }
  }
  __syncthreads();
  int cur = 0;
  for (uint64_t base = 0; base < ns; base += kMtN / 2) {
    const uint32_t* o = st[cur];
    uint32_t* nw = st[cur ^ 1];
    // phase 1: words 0 .. N-M-1 (old words only)
    if (t < kMtN - kMtM) nw[t] = mt_twist(o[t], o[t + 1], o[t + kMtM]);
    __syncthreads();
    // phase 2: words N-M .. 2(N-M)-1 (far word = a NEW word of phase 1)
    if (t < kMtN - kMtM) {
      const int i = t + (kMtN - kMtM);
      nw[i] = mt_twist(o[i], o[i + 1], nw[i - (kMtN - kMtM)]);
    }
    __syncthreads();
    // phase 3: words 2(N-M) .. N-1 (far word = a NEW word of phase 2; the last word wraps to NEW word 0)
    {
      const int i = t + 2 * (kMtN - kMtM);
      if (i < kMtN - 1) nw[i] = mt_twist(o[i], o[i + 1], nw[i - (kMtN - kMtM)]);
      else if (i == kMtN - 1) nw[i] = mt_twist(o[i], nw[0], nw[i - (kMtN - kMtM)]);
    }
    __syncthreads();
    // 312 doubles per twist: libstdc++ generate_canonical<double, 53> takes two draws, low word first
    for (int p = t; p < kMtN / 2; p += kMtThreads) {
      const uint64_t j = base + (uint64_t) p;
      if (j < ns) {
        const double x0 = (double) mt_temper(nw[2 * p]);
        const double x1 = (double) mt_temper(nw[2 * p + 1]);
        const double sum = __dadd_rn(x0, __dmul_rn(x1, 4294967296.0));
        double c = __dmul_rn(sum, 5.421010862427522170037e-20);  // / 2^64, exact
        if (c >= 1.0) c = 0.99999999999999988897769753748;        // nextafter(1, 0)
        rs[j] = __dadd_rn(__dmul_rn(c, max_value), 0.0);
      }
    }
    cur ^= 1;
    // (the next twist writes st[cur ^ 1], which the loop above has finished reading only after this barrier)
    __syncthreads();
  }
}

}  // namespace

size_t sorted_uniform_temp_bytes(uint64_t ns) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, bytes, (const double*) nullptr, (double*) nullptr, (int64_t) ns);
  return (bytes + 255) & ~size_t{255};
}

// d_sorted[0 .. ns) = the values GenerateRandomValues<double>(ns, seed, max_value) returns, in device memory.
// d_max_value != nullptr: the upper bound is read from device memory when the kernel runs.
// d_draws: ns doubles of scratch (draw order); d_temp: sorted_uniform_temp_bytes(ns) bytes.  Enqueued on ctx->stream.
int sorted_uniform_device(qb200_ctx* ctx, unsigned seed, uint64_t ns, double max_value, const double* d_max_value,
                          double* d_draws, double* d_sorted, void* d_temp, size_t temp_bytes) {
  k_mt19937_uniform<<<1, kMtThreads, 0, ctx->stream>>>((uint32_t) seed, ns, max_value, d_max_value, d_draws);
  QB_LAUNCHED(ctx);
  if (cub::DeviceRadixSort::SortKeys(d_temp, temp_bytes, (const double*) d_draws, d_sorted, (int64_t) ns, 0, 64,
                                     ctx->stream) != cudaSuccess) {
    (void) cudaGetLastError();
    return QB200_ERR_CUDA;
  }
  return QB200_OK;
}

}  // namespace qb200
