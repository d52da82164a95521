// gate_kernels.cuh -- fused-gate kernels (replace lib/simulator_cuda_kernels.h:
// ApplyGateH/L, ApplyControlledGateH/LH/L, ExpectationValueH/L).
//
// Design (B200-first):
//  * state in normal order, one group (2^G amplitudes) per thread, the whole
//    2^G x 2^G complex mat-vec in registers;
//  * the gate matrix is a __grid_constant__ kernel parameter: with rows and
//    columns fully unrolled every matrix element is a constant-bank / uniform
//    register FFMA operand, so the inner loop is pure FFMA with no loads;
//  * two adjacent amplitudes per thread moved with one 128-bit access when
//    bit 0 is free (kV2) or when bit 0 is the lowest target (kV2T): a warp
//    request then covers 512 contiguous bytes;
//  * controls cost nothing extra: they only shrink the work-index space
//    (gate_geom.cuh), one kernel serves ApplyGate and ApplyControlledGate.
#pragma once

#include "gate_geom.cuh"

namespace qb200 {

enum GateMode : int { kV1 = 0, kV2 = 1, kV2T = 2 };
constexpr int kRowBatch = 4;  // rows of the unrolled mat-vec scheduled together

// ---- matrix operand sources ------------------------------------------------
template <typename FP, int G>
struct MatParam {  // passed by value as a __grid_constant__ parameter
  FP m[2 << (2 * G)];
  __device__ __forceinline__ FP re(int r, int c) const { return m[2 * ((r << G) + c)]; }
  __device__ __forceinline__ FP im(int r, int c) const { return m[2 * ((r << G) + c) + 1]; }
};

template <typename FP, int G>
struct MatGlobal {  // device pointer (matrices too big for the parameter space)
  const FP* m;
  __device__ __forceinline__ FP re(int r, int c) const { return __ldg(m + 2 * ((r << G) + c)); }
  __device__ __forceinline__ FP im(int r, int c) const { return __ldg(m + 2 * ((r << G) + c) + 1); }
};

// ---- 128-bit / 64-bit accessors -------------------------------------------
__device__ __forceinline__ void ld1(const float* p, float& a, float& b) {
  const float2 v = *reinterpret_cast<const float2*>(p); a = v.x; b = v.y;
}
__device__ __forceinline__ void ld1(const double* p, double& a, double& b) {
  const double2 v = *reinterpret_cast<const double2*>(p); a = v.x; b = v.y;
}
__device__ __forceinline__ void st1(float* p, float a, float b) {
  *reinterpret_cast<float2*>(p) = make_float2(a, b);
}
__device__ __forceinline__ void st1(double* p, double a, double b) {
  *reinterpret_cast<double2*>(p) = make_double2(a, b);
}
__device__ __forceinline__ void ld2(const float* p, float& a, float& b, float& c, float& d) {
  const float4 v = *reinterpret_cast<const float4*>(p); a = v.x; b = v.y; c = v.z; d = v.w;
}
__device__ __forceinline__ void st2(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
// double: two adjacent amplitudes = 32 bytes = two 128-bit accesses
__device__ __forceinline__ void ld2(const double* p, double& a, double& b, double& c, double& d) {
  ld1(p, a, b); ld1(p + 2, c, d);
}
__device__ __forceinline__ void st2(double* p, double a, double b, double c, double d) {
  st1(p, a, b); st1(p + 2, c, d);
}

// Store addresses equal load addresses; left alone, the compiler keeps all 2^G
// 64-bit addresses live across the whole mat-vec (32 registers at G=4).  This
// opaque move makes the store address a fresh value computed next to its use.
__device__ __forceinline__ uint64_t late(uint64_t v) {
  asm volatile("mov.b64 %0, %0;" : "+l"(v));
  return v;
}

// Scheduling fence between row batches of the unrolled mat-vec: the first
// column operand of every later row "changes" here, so ptxas cannot start the
// dependent FFMA chains of later rows early and blow up the register count.
// Emits no instruction.
__device__ __forceinline__ void row_fence(float& a, float& b) { asm volatile("" : "+f"(a), "+f"(b)); }
__device__ __forceinline__ void row_fence(double& a, double& b) { asm volatile("" : "+d"(a), "+d"(b)); }

// one output row of the complex mat-vec: (re,im) = sum_c M[r][c] * x[c]
template <typename FP, int G, typename Mat>
__device__ __forceinline__ void row_dot(const FP (&xr)[1 << G], const FP (&xi)[1 << G],
                                        const Mat& mat, int r, FP& re, FP& im) {
  constexpr int N = 1 << G;
  re = 0;
  im = 0;
#pragma unroll
  for (int c = 0; c < N; ++c) {
    const FP mr = mat.re(r, c), mi = mat.im(r, c);
    re = fma(xr[c], mr, re);
    re = fma(-xi[c], mi, re);
    im = fma(xr[c], mi, im);
    im = fma(xi[c], mr, im);
  }
}

// ---------------------------------------------------------------------------
// Register kernel.  FP, G, MODE compile-time; UNROLL = rows unrolled too.
// EXPECT: read-only pass accumulating <x|M|x> into per-block partials.
// ---------------------------------------------------------------------------
template <typename FP, int G, int MODE, bool UNROLL, bool EXPECT, int NT, int MINB, typename Mat>
__global__ void __launch_bounds__(NT, MINB)
k_gate_reg(FP* __restrict__ st, const __grid_constant__ Geom g,
           const __grid_constant__ Mat mat, double* __restrict__ partials) {
  constexpr int N = 1 << G;
  constexpr int NV = MODE == kV2 ? 2 : 1;
  double ere = 0, eim = 0;

  for (uint64_t i = blockIdx.x * uint64_t{NT} + threadIdx.x; i < g.work;
       i += uint64_t{gridDim.x} * NT) {
    const uint64_t base = expand_index(i, g);
    FP* const p = st + 2 * base;

    FP xr[NV][N], xi[NV][N];
    if constexpr (MODE == kV1) {
#pragma unroll
      for (int k = 0; k < N; ++k) ld1(p + 2 * elem_offset<G>(k, g), xr[0][k], xi[0][k]);
    } else if constexpr (MODE == kV2) {
#pragma unroll
      for (int k = 0; k < N; ++k)
        ld2(p + 2 * elem_offset<G>(k, g), xr[0][k], xi[0][k], xr[1][k], xi[1][k]);
    } else {  // kV2T: target 0 is bit 0 -> elements k and k+1 are adjacent
#pragma unroll
      for (int k = 0; k < N; k += 2)
        ld2(p + 2 * elem_offset<G>(k, g), xr[0][k], xi[0][k], xr[0][k + 1], xi[0][k + 1]);
    }

    if constexpr (EXPECT) {
      static_assert(UNROLL, "expectation kernels index x[r] at compile time");
#pragma unroll
      for (int v = 0; v < NV; ++v) {
#pragma unroll
        for (int r = 0; r < N; ++r) {
          if (r % kRowBatch == 0 && r > 0) row_fence(xr[v][0], xi[v][0]);
          FP re, im;
          row_dot<FP, G>(xr[v], xi[v], mat, r, re, im);
          // products in FP, accumulation in double (lib/simulator_basic.h:323-324)
          ere += xr[v][r] * re + xi[v][r] * im;
          eim += xr[v][r] * im - xi[v][r] * re;
        }
      }
    } else if constexpr (MODE == kV1) {
      auto body = [&](int r) {
        FP re, im;
        row_dot<FP, G>(xr[0], xi[0], mat, r, re, im);
        st1(p + 2 * late(elem_offset<G>(r, g)), re, im);
      };
      if constexpr (UNROLL) {
#pragma unroll
        for (int r = 0; r < N; ++r) {
          if (r % kRowBatch == 0 && r > 0) row_fence(xr[0][0], xi[0][0]);
          body(r);
        }
      } else {
#pragma unroll 2
        for (int r = 0; r < N; ++r) body(r);
      }
    } else if constexpr (MODE == kV2) {
      auto body = [&](int r) {
        FP re0, im0, re1, im1;
        row_dot<FP, G>(xr[0], xi[0], mat, r, re0, im0);
        row_dot<FP, G>(xr[1], xi[1], mat, r, re1, im1);
        st2(p + 2 * late(elem_offset<G>(r, g)), re0, im0, re1, im1);
      };
      if constexpr (UNROLL) {
#pragma unroll
        for (int r = 0; r < N; ++r) {
          if (r % kRowBatch == 0 && r > 0) {
            row_fence(xr[0][0], xi[0][0]);
            row_fence(xr[1][0], xi[1][0]);
          }
          body(r);
        }
      } else {
#pragma unroll 1
        for (int r = 0; r < N; ++r) body(r);
      }
    } else {
      auto body = [&](int r) {
        FP re0, im0, re1, im1;
        row_dot<FP, G>(xr[0], xi[0], mat, r, re0, im0);
        row_dot<FP, G>(xr[0], xi[0], mat, r + 1, re1, im1);
        st2(p + 2 * late(elem_offset<G>(r, g)), re0, im0, re1, im1);
      };
      if constexpr (UNROLL) {
#pragma unroll
        for (int r = 0; r < N; r += 2) {
          if (r % kRowBatch == 0 && r > 0) row_fence(xr[0][0], xi[0][0]);
          body(r);
        }
      } else {
#pragma unroll 1
        for (int r = 0; r < N; r += 2) body(r);
      }
    }
  }

  if constexpr (EXPECT) {
    block_sum2<NT>(ere, eim);
    if (threadIdx.x == 0) {
      partials[2 * blockIdx.x] = ere;
      partials[2 * blockIdx.x + 1] = eim;
    }
  }
}

// ---------------------------------------------------------------------------
// Runtime-generic kernel: any G <= 6, any dtype; each thread's group lives in
// shared memory (column-major over threads -> conflict free), matrix read from
// global memory through the read-only path.  Used for fp32 G=6, fp64 G>=5 and
// as the forced fallback (tuning "force_generic") that cross-checks the
// register kernels.
// ---------------------------------------------------------------------------
template <typename FP, bool EXPECT, int NT>
__global__ void __launch_bounds__(NT)
k_gate_generic(FP* __restrict__ st, const __grid_constant__ Geom g,
               const FP* __restrict__ mat, double* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FP* const sx = reinterpret_cast<FP*>(smem_raw);  // [2][N][NT]
  const int G = g.nt;
  const int N = 1 << G;
  FP* const sr = sx + threadIdx.x;
  FP* const si = sx + N * NT + threadIdx.x;
  double ere = 0, eim = 0;

  for (uint64_t i = blockIdx.x * uint64_t{NT} + threadIdx.x; i < g.work;
       i += uint64_t{gridDim.x} * NT) {
    const uint64_t base = expand_index(i, g);
    FP* const p = st + 2 * base;
    for (int k = 0; k < N; ++k) {
      uint64_t o = 0;
      for (int j = 0; j < G; ++j)
        if ((k >> j) & 1) o += g.xs[j];
      FP a, b;
      ld1(p + 2 * o, a, b);
      sr[k * NT] = a;
      si[k * NT] = b;
    }
    for (int r = 0; r < N; ++r) {
      const FP* mrow = mat + 2 * (size_t) r * N;
      FP re0 = 0, re1 = 0, im0 = 0, im1 = 0;
#pragma unroll 4
      for (int c = 0; c < N; ++c) {
        const FP mr = __ldg(mrow + 2 * c), mi = __ldg(mrow + 2 * c + 1);
        const FP a = sr[c * NT], b = si[c * NT];
        re0 = fma(a, mr, re0);
        re1 = fma(-b, mi, re1);
        im0 = fma(a, mi, im0);
        im1 = fma(b, mr, im1);
      }
      const FP re = re0 + re1, im = im0 + im1;
      if constexpr (EXPECT) {
        const FP a = sr[r * NT], b = si[r * NT];
        ere += a * re + b * im;
        eim += a * im - b * re;
      } else {
        uint64_t o = 0;
        for (int j = 0; j < G; ++j)
          if ((r >> j) & 1) o += g.xs[j];
        st1(p + 2 * o, re, im);
      }
    }
  }

  if constexpr (EXPECT) {
    block_sum2<NT>(ere, eim);
    if (threadIdx.x == 0) {
      partials[2 * blockIdx.x] = ere;
      partials[2 * blockIdx.x + 1] = eim;
    }
  }
}

}  // namespace qb200
