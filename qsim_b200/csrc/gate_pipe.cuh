// gate_pipe.cuh -- deep-prefetch fused-gate kernel for the compute-heavy gates
// (fp32, G >= 4).
//
// Why: a G=4 fp32 pass needs ~32 FFMA2 per amplitude, which is within ~25% of
// the FMA-pipe time per HBM byte, so the pass is only HBM-bound if the loads of
// later groups are in flight *while* the current group is in the FMA pipe.  The
// plain register kernel (gate_kernels.cuh) holds one group per thread in ~140
// registers: 12 warps/SM, ~48 KB of loads in flight per SM, and ncu shows
// long-scoreboard stalls with DRAM at 60%.  Here every thread streams its groups
// through a private D-deep ring in shared memory filled by cp.async (LDGSTS):
// up to (D-1) x NT x 128 B of HBM requests stay in flight per CTA independent of
// register count, no block-level synchronisation is needed (a thread only ever
// reads the bytes it copied itself), and the mat-vec is the same FFMA2 code.
#pragma once

#include "gate_kernels.cuh"

namespace qb200 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t) __cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// smem ring: stage s, slot q (8-byte slots for kV1, 16-byte for kV2/kV2T), thread t
//   byte address = ((s * SLOTS + q) * NT + t) * SLOT_BYTES        (conflict free)
template <int G, int MODE, int NT, int D, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_gate_pipe(float* __restrict__ st, const __grid_constant__ Geom g,
            const __grid_constant__ MatParam<float, G> mat) {
  constexpr int N = 1 << G;
  constexpr int NV = MODE == kV2 ? 2 : 1;
  constexpr int SLOT_BYTES = MODE == kV1 ? 8 : 16;
  constexpr int SLOTS = MODE == kV2T ? N / 2 : N;
  using C = CT<float>::type;
  extern __shared__ __align__(16) unsigned char ring[];
  __shared__ uint64_t base_ring[D][NT];  // amplitude index of the group's base, per stage

  const uint32_t ring0 = smem_u32(ring) + threadIdx.x * SLOT_BYTES;
  const uint64_t stride = uint64_t{gridDim.x} * NT;
  const uint64_t first = blockIdx.x * uint64_t{NT} + threadIdx.x;

  auto issue = [&](uint64_t i, int s) {
    if (i < g.work) {
      const uint64_t base = expand_index(i, g);
      base_ring[s][threadIdx.x] = base;
      const float* p = st + 2 * base;
      const uint32_t dst = ring0 + (uint32_t) (s * SLOTS) * NT * SLOT_BYTES;
#pragma unroll
      for (int q = 0; q < SLOTS; ++q) {
        const int k = MODE == kV2T ? 2 * q : q;
        if constexpr (MODE == kV1) cp_async8(dst + q * NT * SLOT_BYTES, p + 2 * elem_offset<G>(k, g));
        else cp_async16(dst + q * NT * SLOT_BYTES, p + 2 * elem_offset<G>(k, g));
      }
    }
    cp_async_commit();
  };

#pragma unroll
  for (int s = 0; s < D - 1; ++s) issue(first + s * stride, s);

  int s = 0;
  for (uint64_t i = first; i < g.work; i += stride) {
    int sp = s + D - 1;
    if (sp >= D) sp -= D;
    issue(i + (D - 1) * stride, sp);
    cp_async_wait<D - 1>();

    C x[NV][N], ix[NV][N];
    const unsigned char* src = ring + ((size_t) (s * SLOTS) * NT + threadIdx.x) * SLOT_BYTES;
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) {
      if constexpr (MODE == kV1) {
        x[0][q] = *reinterpret_cast<const uint64_t*>(src + (size_t) q * NT * SLOT_BYTES);
      } else {
        const uint2* v = reinterpret_cast<const uint2*>(src + (size_t) q * NT * SLOT_BYTES);
        const uint4 w = *reinterpret_cast<const uint4*>(v);
        const uint64_t lo = (uint64_t) w.x | ((uint64_t) w.y << 32);
        const uint64_t hi = (uint64_t) w.z | ((uint64_t) w.w << 32);
        if constexpr (MODE == kV2) { x[0][q] = lo; x[1][q] = hi; }
        else { x[0][2 * q] = lo; x[0][2 * q + 1] = hi; }
      }
    }
    float* const p = st + 2 * base_ring[s][threadIdx.x];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int k = 0; k < N; ++k) ix[v][k] = CT<float>::rot(x[v][k]);

    if constexpr (MODE == kV1) {
#pragma unroll
      for (int r = 0; r < N; ++r) {
        if (r % kRowBatch == 0 && r > 0) CT<float>::fence(x[0][0]);
        stc1<float>(p + 2 * elem_offset<G>(r, g), row_dot<float, G>(x[0], ix[0], mat, r));
      }
    } else if constexpr (MODE == kV2) {
#pragma unroll
      for (int r = 0; r < N; ++r) {
        if (r % kRowBatch == 0 && r > 0) { CT<float>::fence(x[0][0]); CT<float>::fence(x[1][0]); }
        const C a = row_dot<float, G>(x[0], ix[0], mat, r);
        const C b = row_dot<float, G>(x[1], ix[1], mat, r);
        stc2<float>(p + 2 * elem_offset<G>(r, g), a, b);
      }
    } else {
#pragma unroll
      for (int r = 0; r < N; r += 2) {
        if (r % kRowBatch == 0 && r > 0) CT<float>::fence(x[0][0]);
        const C a = row_dot<float, G>(x[0], ix[0], mat, r);
        const C b = row_dot<float, G>(x[0], ix[0], mat, r + 1);
        stc2<float>(p + 2 * elem_offset<G>(r, g), a, b);
      }
    }
    if (++s == D) s = 0;
  }
  cp_async_wait<0>();
}

template <int G, int MODE, int NT, int D>
constexpr size_t pipe_smem_bytes() {
  return (size_t) D * (1 << G) * NT * 8 * (MODE == kV2 ? 2 : 1);
}

}  // namespace qb200
