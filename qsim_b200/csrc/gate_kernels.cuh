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

// ---- matrix operand -----------------------------------------------------------
// Passed by value as a __grid_constant__ kernel parameter: with rows and columns
// unrolled every element is fetched by LDCU straight into uniform registers and
// used as a broadcast scalar operand -- no per-block matrix staging at all.
template <typename FP, int G>
struct MatParam {
  FP m[2 << (2 * G)];  // row-major, interleaved (re, im), as the caller passes it
  void fill(const FP* src) {
    for (int i = 0; i < (2 << (2 * G)); ++i) m[i] = src[i];
  }
};

// ---- complex values in registers ------------------------------------------------
// fp32: one amplitude = one 64-bit register pair (re, im), exactly as an LDG.64 /
// LDG.128 delivers it.  On sm_100 the scalar FFMA issues at half rate; the packed
// FFMA2 (fma.rn.f32x2) does two FMAs per lane per issue.  One complex MAC
//     acc += x * mr;   acc += (i x) * mi        with  i x = (-im, re)
// is two FFMA2 whose first operand is the natural (or lane-swapped, ptxas folds
// the swap/negate into operand modifiers) amplitude pair and whose second operand
// is a broadcast uniform-register scalar: no packing moves, one LDCU.128 feeds
// four FFMA2 per amplitude group.
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

template <typename FP> struct CT;
template <> struct CT<float> {
  using type = uint64_t;
  static __device__ __forceinline__ type make(float re, float im) { return pack2(re, im); }
  static __device__ __forceinline__ void get(type v, float& re, float& im) { unpack2(v, re, im); }
  static __device__ __forceinline__ type rot(type v) { float a, b; unpack2(v, a, b); return pack2(-b, a); }
  static __device__ __forceinline__ type mac(type acc, type x, type ix, float mr, float mi) {
    acc = fma2(x, pack2(mr, mr), acc);
    return fma2(ix, pack2(mi, mi), acc);
  }
  static __device__ __forceinline__ void fence(type& v) { asm volatile("" : "+l"(v)); }
};
template <> struct CT<double> {
  struct type { double re, im; };
  static __device__ __forceinline__ type make(double re, double im) { return type{re, im}; }
  static __device__ __forceinline__ void get(type v, double& re, double& im) { re = v.re; im = v.im; }
  static __device__ __forceinline__ type rot(type v) { return v; }  // unused by mac
  static __device__ __forceinline__ type mac(type acc, type x, type, double mr, double mi) {
    acc.re = fma(x.re, mr, acc.re);
    acc.re = fma(-x.im, mi, acc.re);
    acc.im = fma(x.re, mi, acc.im);
    acc.im = fma(x.im, mr, acc.im);
    return acc;
  }
  static __device__ __forceinline__ void fence(type& v) { asm volatile("" : "+d"(v.re), "+d"(v.im)); }
};

// ---- 64-bit / 128-bit accessors -----------------------------------------------
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
template <typename FP>
__device__ __forceinline__ typename CT<FP>::type ldc1(const FP* p) {
  FP a, b; ld1(p, a, b); return CT<FP>::make(a, b);
}
template <typename FP>
__device__ __forceinline__ void ldc2(const FP* p, typename CT<FP>::type& u, typename CT<FP>::type& v) {
  FP a, b, c, d; ld2(p, a, b, c, d); u = CT<FP>::make(a, b); v = CT<FP>::make(c, d);
}
template <typename FP>
__device__ __forceinline__ void stc1(FP* p, typename CT<FP>::type u) {
  FP a, b; CT<FP>::get(u, a, b); st1(p, a, b);
}
template <typename FP>
__device__ __forceinline__ void stc2(FP* p, typename CT<FP>::type u, typename CT<FP>::type v) {
  FP a, b, c, d; CT<FP>::get(u, a, b); CT<FP>::get(v, c, d); st2(p, a, b, c, d);
}

// Store addresses equal load addresses; left alone, the compiler keeps all 2^G
// 64-bit addresses live across the whole mat-vec (32 registers at G=4).  This
// opaque move makes the store address a fresh value computed next to its use.
__device__ __forceinline__ uint64_t late(uint64_t v) {
  asm volatile("mov.b64 %0, %0;" : "+l"(v));
  return v;
}

// one output row of the complex mat-vec: sum_c M[r][c] * x[c]
template <typename FP, int G>
__device__ __forceinline__ typename CT<FP>::type
row_dot(const typename CT<FP>::type (&x)[1 << G], const typename CT<FP>::type (&ix)[1 << G],
        const MatParam<FP, G>& mat, int r) {
  constexpr int N = 1 << G;
  typename CT<FP>::type acc = CT<FP>::make(0, 0);
#pragma unroll
  for (int c = 0; c < N; ++c) {
    const FP mr = mat.m[2 * ((r << G) + c)], mi = mat.m[2 * ((r << G) + c) + 1];
    acc = CT<FP>::mac(acc, x[c], ix[c], mr, mi);
  }
  return acc;
}

// ---------------------------------------------------------------------------
// Register kernel.  FP, G, MODE compile-time; UNROLL = rows unrolled too.
// EXPECT: read-only pass accumulating <x|M|x> into per-block partials.
//   kV1  one group per thread, one amplitude per access (64-bit fp32 / 128-bit fp64)
//   kV2  two groups per thread = the two amplitudes of one 128-bit access (bit 0 free)
//   kV2T one group per thread, bit 0 is the lowest target: elements k, k+1 share a
//        128-bit access
// ---------------------------------------------------------------------------
template <typename FP, int G, int MODE, bool UNROLL, bool EXPECT, bool PREFETCH, int NT, int MINB, typename Mat>
__global__ void __launch_bounds__(NT, MINB)
k_gate_reg(FP* __restrict__ st, const __grid_constant__ Geom g,
           const __grid_constant__ Mat mat, double* __restrict__ partials) {
  constexpr int N = 1 << G;
  constexpr int NV = MODE == kV2 ? 2 : 1;
  using C = typename CT<FP>::type;
  double ere = 0, eim = 0;

  auto load_group = [&](uint64_t i, FP*& p, C (&x)[NV][N]) {
    p = st + 2 * expand_index(i, g);
    if constexpr (MODE == kV1) {
#pragma unroll
      for (int k = 0; k < N; ++k) x[0][k] = ldc1<FP>(p + 2 * elem_offset<G>(k, g));
    } else if constexpr (MODE == kV2) {
#pragma unroll
      for (int k = 0; k < N; ++k) ldc2<FP>(p + 2 * elem_offset<G>(k, g), x[0][k], x[1][k]);
    } else {
#pragma unroll
      for (int k = 0; k < N; k += 2) ldc2<FP>(p + 2 * elem_offset<G>(k, g), x[0][k], x[0][k + 1]);
    }
  };

  // PREFETCH: persistent grid-stride loop, the next group's loads are issued
  // before the current group's mat-vec so every warp keeps a full group of
  // HBM requests in flight while its FMA pipe is busy.
  const uint64_t stride = uint64_t{gridDim.x} * NT;
  uint64_t i = blockIdx.x * uint64_t{NT} + threadIdx.x;
  FP* pn = st;
  C xn[NV][N];
  if constexpr (PREFETCH) {
    if (i < g.work) load_group(i, pn, xn);
  }

  for (; i < g.work; i += stride) {
    FP* p;
    C x[NV][N], ix[NV][N];
    if constexpr (PREFETCH) {
      p = pn;
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < N; ++k) x[v][k] = xn[v][k];
      if (i + stride < g.work) load_group(i + stride, pn, xn);
    } else {
      load_group(i, p, x);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int k = 0; k < N; ++k) ix[v][k] = CT<FP>::rot(x[v][k]);

    if constexpr (EXPECT) {
      static_assert(UNROLL, "expectation kernels index x[r] at compile time");
#pragma unroll
      for (int v = 0; v < NV; ++v) {
#pragma unroll
        for (int r = 0; r < N; ++r) {
          if (r % kRowBatch == 0 && r > 0) CT<FP>::fence(x[v][0]);
          FP re, im, xr, xi;
          CT<FP>::get(row_dot<FP, G>(x[v], ix[v], mat, r), re, im);
          CT<FP>::get(x[v][r], xr, xi);
          // products in FP, accumulation in double (lib/simulator_basic.h:323-324)
          ere += xr * re + xi * im;
          eim += xr * im - xi * re;
        }
      }
    } else if constexpr (MODE == kV1) {
      auto body = [&](int r) {
        stc1<FP>(p + 2 * late(elem_offset<G>(r, g)), row_dot<FP, G>(x[0], ix[0], mat, r));
      };
      if constexpr (UNROLL) {
#pragma unroll
        for (int r = 0; r < N; ++r) {
          if (r % kRowBatch == 0 && r > 0) CT<FP>::fence(x[0][0]);
          body(r);
        }
      } else {
#pragma unroll 2
        for (int r = 0; r < N; ++r) body(r);
      }
    } else if constexpr (MODE == kV2) {
      auto body = [&](int r) {
        const C a = row_dot<FP, G>(x[0], ix[0], mat, r);
        const C b = row_dot<FP, G>(x[1], ix[1], mat, r);
        stc2<FP>(p + 2 * late(elem_offset<G>(r, g)), a, b);
      };
      if constexpr (UNROLL) {
#pragma unroll
        for (int r = 0; r < N; ++r) {
          if (r % kRowBatch == 0 && r > 0) {
            CT<FP>::fence(x[0][0]);
            CT<FP>::fence(x[1][0]);
          }
          body(r);
        }
      } else {
#pragma unroll 1
        for (int r = 0; r < N; ++r) body(r);
      }
    } else {
      auto body = [&](int r) {
        const C a = row_dot<FP, G>(x[0], ix[0], mat, r);
        const C b = row_dot<FP, G>(x[0], ix[0], mat, r + 1);
        stc2<FP>(p + 2 * late(elem_offset<G>(r, g)), a, b);
      };
      if constexpr (UNROLL) {
#pragma unroll
        for (int r = 0; r < N; r += 2) {
          if (r % kRowBatch == 0 && r > 0) CT<FP>::fence(x[0][0]);
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
// Streaming expectation kernel for small gates (fp32 G <= 2): a read-only pass
// with almost no arithmetic, so what sets its speed is the number of bytes each
// thread keeps in flight.  UG independent groups (grid-strided, so every access
// stays coalesced) are loaded per iteration, chosen by the launcher so that a
// thread issues four 128-bit (or 64-bit, kV1) loads per iteration whatever the
// layout, and the next iteration's loads are issued before the current one is
// consumed.  Persistent grid.  Arithmetic as in k_gate_reg<EXPECT>: products in
// FP, accumulation in double (lib/simulator_basic.h:323-324).
// ---------------------------------------------------------------------------
template <typename FP, int G, int MODE, int UG, int NT, int MINB, typename Mat>
__global__ void __launch_bounds__(NT, MINB)
k_expect_stream(const FP* __restrict__ st, const __grid_constant__ Geom g,
                const __grid_constant__ Mat mat, double* __restrict__ partials, const int contiguous) {
  constexpr int N = 1 << G;
  constexpr int NV = MODE == kV2 ? 2 : 1;
  using C = typename CT<FP>::type;
  double ere = 0, eim = 0;
  // a thread's UG groups per iteration: NT apart inside the block's contiguous chunk of UG * NT work items
  // (`contiguous`), or one grid stride apart
  const uint64_t stride = contiguous ? uint64_t{NT} : uint64_t{gridDim.x} * NT;
  const uint64_t advance = uint64_t{gridDim.x} * NT * UG;

  // branch-free: a group beyond the end re-reads the thread's first (valid) group and is zeroed by selects,
  // so all of an iteration's loads issue back to back
  auto load = [&](uint64_t i, C (&x)[UG][NV][N]) {
#pragma unroll
    for (int u = 0; u < UG; ++u) {
      const uint64_t iu = i + u * stride;
      const bool ok = u == 0 || iu < g.work;
      const FP* p = st + 2 * expand_index(ok ? iu : i, g);
      if constexpr (MODE == kV1) {
#pragma unroll
        for (int k = 0; k < N; ++k) x[u][0][k] = ldc1<FP>(p + 2 * elem_offset<G>(k, g));
      } else if constexpr (MODE == kV2) {
#pragma unroll
        for (int k = 0; k < N; ++k) ldc2<FP>(p + 2 * elem_offset<G>(k, g), x[u][0][k], x[u][1][k]);
      } else {
#pragma unroll
        for (int k = 0; k < N; k += 2) ldc2<FP>(p + 2 * elem_offset<G>(k, g), x[u][0][k], x[u][0][k + 1]);
      }
      if (u > 0) {
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
          for (int k = 0; k < N; ++k) x[u][v][k] = ok ? x[u][v][k] : CT<FP>::make(0, 0);  // contributes nothing
      }
    }
  };

  uint64_t i = blockIdx.x * uint64_t{NT} * (contiguous ? UG : 1) + threadIdx.x;
  C xn[UG][NV][N];
  if (i < g.work) load(i, xn);
  for (; i < g.work; i += advance) {
    C x[UG][NV][N];
#pragma unroll
    for (int u = 0; u < UG; ++u)
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int k = 0; k < N; ++k) x[u][v][k] = xn[u][v][k];
    if (i + advance < g.work) load(i + advance, xn);
#pragma unroll
    for (int u = 0; u < UG; ++u) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        C ix[N];
#pragma unroll
        for (int k = 0; k < N; ++k) ix[k] = CT<FP>::rot(x[u][v][k]);
#pragma unroll
        for (int r = 0; r < N; ++r) {
          FP re, im, xr, xi;
          CT<FP>::get(row_dot<FP, G>(x[u][v], ix, mat, r), re, im);
          CT<FP>::get(x[u][v][r], xr, xi);
          ere += xr * re + xi * im;
          eim += xr * im - xi * re;
        }
      }
    }
  }

  block_sum2<NT>(ere, eim);
  if (threadIdx.x == 0) {
    partials[2 * blockIdx.x] = ere;
    partials[2 * blockIdx.x + 1] = eim;
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
