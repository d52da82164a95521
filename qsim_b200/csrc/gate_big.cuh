// gate_big.cuh -- fp32 fused gates and expectation values on 5 and 6 qubits
// (ApplyGateH/L<5,6>, ExpectationValueH/L<5,6> of lib/simulator_cuda_kernels.h).
//
// Regime.  8*2^G flops per 16 bytes: G=5 is 16 flop/B, G=6 is 32 flop/B.  Against
// ~70 TFLOP/s of FFMA2 (tools/ubench/fma_rate.cu) and ~6.5 TB/s of HBM the ridge sits
// at ~11 flop/B, so these passes are bound by the FMA pipe, not by HBM: the job here is
// to keep the pipe issuing FFMA2 back to back with nothing else in the way.
//
// Shape.  Same warp tile as gate_tile.cuh (32 groups, coalesced 16-byte cp.async in,
// STG.128 out, XOR-swizzled shared memory, warp-level synchronisation only).  Each lane
// pulls its whole group (2^G amplitudes = 64/128 registers) out of the tile and runs the
// mat-vec in row blocks of kBigRows rows: rows inside a block are unrolled, the block
// index is a run-time (warp-uniform) value, so the instruction footprint is
// 2*kBigRows*2^G FFMA2 (16 KB-32 KB of SASS) instead of 131 KB for a fully unrolled G=6.
// Matrix elements are broadcast scalars fetched by LDCU from the kernel parameter
// (G=5, 8 KB) or from __constant__ memory (G=6: 32 KB does not fit the parameter space).
//
// Complex MAC without a rotated copy of x (that would double the register footprint):
//     p += x * (mr, mr);   q += x * (mi, mi);      result = (p.re - q.im, p.im + q.re)
// two independent FFMA2 chains per row, one FADD pair per row at the end.
//
// EXPECT: read-only pass, <x|M|x> accumulated in double; x[r] is re-read from the tile.
#pragma once

#include "gate_tile.cuh"

namespace qb200 {

constexpr int kBigRows = 16;

// 6-qubit fp32 matrix (row-major, interleaved).  One per device: launch_big serialises
// its users across streams with an event (gates_f32_big.cu).
__constant__ __align__(16) float c_mat6[2 << 12];

template <int G>
struct alignas(16) BigMat {  // G == 5: by-value kernel parameter; G == 6: empty tag, data in c_mat6
  float m[G == 5 ? (2 << 10) : 2];
};

// two adjacent matrix elements (r, c), (r, c+1) with one 128-bit uniform load
template <int G>
__device__ __forceinline__ float4 big_elem2(const BigMat<G>& mat, int r, int c) {
  if constexpr (G == 6) return reinterpret_cast<const float4*>(c_mat6)[((r << 6) + c) >> 1];
  else return reinterpret_cast<const float4*>(mat.m)[((r << G) + c) >> 1];
}

template <int G>
__device__ __forceinline__ uint64_t big_row(const uint64_t (&x)[1 << G], const BigMat<G>& mat, int r) {
  constexpr int N = 1 << G;
  uint64_t p = 0, q = 0;  // (+0.0f, +0.0f)
#pragma unroll
  for (int c = 0; c < N; c += 2) {
    const float4 e = big_elem2<G>(mat, r, c);
    p = fma2(x[c], pack2(e.x, e.x), p);
    q = fma2(x[c], pack2(e.y, e.y), q);
    p = fma2(x[c + 1], pack2(e.z, e.z), p);
    q = fma2(x[c + 1], pack2(e.w, e.w), q);
  }
  float pr, pi, qr, qi;
  unpack2(p, pr, pi);
  unpack2(q, qr, qi);
  return pack2(pr - qi, pi + qr);
}

template <int G, bool PAIR, bool EXPECT, int NT, int D>
__global__ void __launch_bounds__(NT, 1)
k_gate_big(float* __restrict__ st, const __grid_constant__ TileGeom t,
           const __grid_constant__ BigMat<G> mat, double* __restrict__ partials) {
  constexpr int N = 1 << G;
  constexpr int TILE_BYTES = 8 << (G + 5);
  constexpr int ROWS = N / 2;  // 16-byte chunk rows per lane
  constexpr int WARPS = NT / 32;
  extern __shared__ __align__(1024) unsigned char ring_raw[];  // [WARPS][D][TILE_BYTES] + pad
  __shared__ uint64_t base_ring[WARPS][D];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t pad = (TILE_BYTES - (smem_u32(ring_raw) & (TILE_BYTES - 1))) & (TILE_BYTES - 1);
  unsigned char* const wring = ring_raw + pad + (size_t) warp * D * TILE_BYTES;
  const uint32_t wring_s = smem_u32(wring);

  uint64_t goff_l = 0;
  uint32_t jl = (uint32_t) lane << 1, jw = 0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    goff_l |= (uint64_t) ((lane >> i) & 1) << t.bpos[i + 1];
    jw |= (uint32_t) ((lane >> i) & 1) << t.fl[i];
  }
  const uint32_t chunk_l = tile_swz(jl, t) << 3;
  const uint32_t group_l = tile_swz(jw, t) << 3;

  const uint64_t stride = uint64_t{gridDim.x} * WARPS;
  const uint64_t first = blockIdx.x * uint64_t{WARPS} + warp;
  double ere = 0, eim = 0;

  auto tile_base = [&](uint64_t i) {
    for (uint32_t k = 0; k < t.npos; ++k) {
      const uint64_t lo = i & ((uint64_t{1} << t.pos[k]) - 1);
      i = ((i - lo) << 1) | lo;
    }
    return i | t.cbits;
  };

  auto issue = [&](uint64_t i, int s) {
    if (i < t.work) {
      const uint64_t base = tile_base(i);
      if (lane == 0) base_ring[warp][s] = base;
      const float* p = st + 2 * (base + goff_l);
      const uint32_t dst = (wring_s + s * TILE_BYTES) | chunk_l;
#pragma unroll
      for (int m = 0; m < ROWS; ++m) cp_async16(dst ^ t.sm_m[m], p + 2 * t.goff_m[m]);
    }
    cp_async_commit();
  };

#pragma unroll
  for (int s = 0; s < D - 1; ++s) issue(first + s * stride, s);

  int s = 0;
  for (uint64_t i = first; i < t.work; i += stride) {
    int sp = s + D - 1;
    if (sp >= D) sp -= D;
    __syncwarp();  // every lane is done with stage sp (consumed last iteration)
    issue(i + (D - 1) * stride, sp);
    cp_async_wait<D - 1>();
    __syncwarp();  // all lanes' copies of stage s have landed

    unsigned char* const tile = wring + s * TILE_BYTES;
    const uint32_t ga = group_l;
    uint64_t x[N];
    if constexpr (PAIR) {
#pragma unroll
      for (int k = 0; k < N; k += 2) {
        const uint4 w = *reinterpret_cast<const uint4*>(tile + (ga ^ t.skb[k]));
        x[k] = (uint64_t) w.x | ((uint64_t) w.y << 32);
        x[k + 1] = (uint64_t) w.z | ((uint64_t) w.w << 32);
      }
    } else {
#pragma unroll
      for (int k = 0; k < N; ++k) x[k] = *reinterpret_cast<const uint64_t*>(tile + (ga ^ t.skb[k]));
    }

#pragma unroll 1
    for (int rb = 0; rb < N; rb += kBigRows) {
#pragma unroll
      for (int rr = 0; rr < kBigRows; rr += (PAIR ? 2 : 1)) {
        const int r = rb + rr;
        if (rr % kRowBatch == 0 && rr > 0) CT<float>::fence(x[0]);
        const uint64_t a = big_row<G>(x, mat, r);
        if constexpr (PAIR) {
          const uint64_t b = big_row<G>(x, mat, r + 1);
          if constexpr (EXPECT) {
            const uint4 w = *reinterpret_cast<const uint4*>(tile + (ga ^ t.skb[r]));
            const float xr0 = __uint_as_float(w.x), xi0 = __uint_as_float(w.y);
            const float xr1 = __uint_as_float(w.z), xi1 = __uint_as_float(w.w);
            float re, im;
            unpack2(a, re, im);
            ere += xr0 * re + xi0 * im;
            eim += xr0 * im - xi0 * re;
            unpack2(b, re, im);
            ere += xr1 * re + xi1 * im;
            eim += xr1 * im - xi1 * re;
          } else {
            *reinterpret_cast<uint4*>(tile + (ga ^ t.skb[r])) =
                make_uint4((uint32_t) a, (uint32_t) (a >> 32), (uint32_t) b, (uint32_t) (b >> 32));
          }
        } else {
          if constexpr (EXPECT) {
            const float2 w = *reinterpret_cast<const float2*>(tile + (ga ^ t.skb[r]));
            float re, im;
            unpack2(a, re, im);
            // products in FP, accumulation in double (lib/simulator_basic.h:323-324)
            ere += w.x * re + w.y * im;
            eim += w.x * im - w.y * re;
          } else {
            *reinterpret_cast<uint64_t*>(tile + (ga ^ t.skb[r])) = a;
          }
        }
      }
    }

    if constexpr (!EXPECT) {
      __syncwarp();
      // coalesced write-back: 16 bytes per lane, 512 contiguous bytes per request
      float* const p = st + 2 * (base_ring[warp][s] + goff_l);
#pragma unroll
      for (int m = 0; m < ROWS; ++m) {
        const uint4 v = *reinterpret_cast<const uint4*>(tile + (chunk_l ^ t.sm_m[m]));
        *reinterpret_cast<uint4*>(p + 2 * t.goff_m[m]) = v;
      }
    }
    if (++s == D) s = 0;
  }
  cp_async_wait<0>();

  if constexpr (EXPECT) {
    block_sum2<NT>(ere, eim);
    if (threadIdx.x == 0) {
      partials[2 * blockIdx.x] = ere;
      partials[2 * blockIdx.x + 1] = eim;
    }
  }
}

}  // namespace qb200
