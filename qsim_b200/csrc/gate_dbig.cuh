// gate_dbig.cuh -- fp64 fused gates on 5 and 6 qubits and fp64 expectation values on 4..6 qubits
// (ApplyGateH/L<5,6>, ExpectationValueH/L<4..6> of lib/simulator_cuda_kernels.h in double).
//
// fp64 has no packed FMA and no tensor-core path at this accuracy, so these passes are bound
// by the DFMA pipe: 8*2^G flop per 32 B against ~37 TFLOP/s is 7.4 ms at n = 29 for G = 6, far
// above the 2.6 ms of HBM time.  The job is to keep DFMA issuing: the runtime-generic kernel
// (one LDS pair and one LDG pair per complex MAC, 138 ms) does not.
//
// Shape.  One group per thread.  The thread's 2^G amplitudes live in shared memory,
// column-major over the threads of the block (re and im planes, conflict free) -- 2^6 complex
// doubles do not fit the register file.  The mat-vec runs in row blocks of kDRows rows held in
// registers (2 * kDRows independent accumulator chains); per column one LDS.64 pair feeds
// 4 * kDRows DFMA, matrix elements are warp-uniform 128-bit loads from device memory (L1
// resident: 64 KB at most).  Results go straight from registers to global memory; expectation
// values fold <x|y> per row instead, accumulated in double like lib/simulator_basic.h:323-324.
#pragma once

#include "gate_kernels.cuh"

namespace qb200 {

constexpr int kDRows = 8;
constexpr int kDThreads = 64;

template <int G, bool EXPECT>
__global__ void __launch_bounds__(kDThreads)
k_gate_dbig(double* __restrict__ st, const __grid_constant__ Geom g, const double2* __restrict__ mat,
            double* __restrict__ partials) {
  constexpr int N = 1 << G;
  extern __shared__ __align__(16) unsigned char dsm_raw[];
  double* const sre = reinterpret_cast<double*>(dsm_raw) + threadIdx.x;        // [N][kDThreads]
  double* const sim = sre + N * kDThreads;
  double ere = 0, eim = 0;

  for (uint64_t i = blockIdx.x * uint64_t{kDThreads} + threadIdx.x; i < g.work; i += uint64_t{gridDim.x} * kDThreads) {
    double* const p = st + 2 * expand_index(i, g);
#pragma unroll 8
    for (int k = 0; k < N; ++k) {
      const double2 v = *reinterpret_cast<const double2*>(p + 2 * elem_offset<G>(k, g));
      sre[k * kDThreads] = v.x;
      sim[k * kDThreads] = v.y;
    }
#pragma unroll 1
    for (int rb = 0; rb < N; rb += kDRows) {
      double ar[kDRows], ai[kDRows];
#pragma unroll
      for (int r = 0; r < kDRows; ++r) ar[r] = ai[r] = 0.0;
      const double2* mrow = mat + (size_t) rb * N;
#pragma unroll 4
      for (int c = 0; c < N; ++c) {
        const double xr = sre[c * kDThreads], xi = sim[c * kDThreads];
#pragma unroll
        for (int r = 0; r < kDRows; ++r) {
          const double2 m = __ldg(mrow + r * N + c);  // warp-uniform address: one broadcast transaction
          ar[r] = fma(xr, m.x, ar[r]);
          ar[r] = fma(-xi, m.y, ar[r]);
          ai[r] = fma(xr, m.y, ai[r]);
          ai[r] = fma(xi, m.x, ai[r]);
        }
      }
      if constexpr (EXPECT) {
#pragma unroll
        for (int r = 0; r < kDRows; ++r) {
          const double xr = sre[(rb + r) * kDThreads], xi = sim[(rb + r) * kDThreads];
          ere += xr * ar[r] + xi * ai[r];
          eim += xr * ai[r] - xi * ar[r];
        }
      } else {
#pragma unroll
        for (int r = 0; r < kDRows; ++r) {
          // row index is run-time: the element offset is evaluated from its bits
          const int k = rb + r;
          uint64_t o = 0;
#pragma unroll
          for (int j = 0; j < G; ++j)
            if ((k >> j) & 1) o += g.xs[j];
          *reinterpret_cast<double2*>(p + 2 * o) = make_double2(ar[r], ai[r]);
        }
      }
    }
  }

  if constexpr (EXPECT) {
    block_sum2<kDThreads>(ere, eim);
    if (threadIdx.x == 0) {
      partials[2 * blockIdx.x] = ere;
      partials[2 * blockIdx.x + 1] = eim;
    }
  }
}

template <int G, bool EXPECT>
int launch_dbig(qb200_ctx* ctx, double* st, const Geom& g, const double* m, double* out) {
  auto kern = k_gate_dbig<G, EXPECT>;
  constexpr size_t smem = (size_t{2} << G) * kDThreads * sizeof(double);
  static PerDevice occ_cache;
  const int occ = occ_cache.get(ctx, [&] {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kDThreads, smem) != cudaSuccess || nb < 1) {
      (void) cudaGetLastError();
      nb = 1;
    }
    return nb;
  });
  const void* dmat = nullptr;
  int rc = stage_matrix(ctx, m, sizeof(double) * (size_t{2} << (2 * G)), &dmat);
  if (rc) return rc;
  const uint64_t need = (g.work + kDThreads - 1) / kDThreads;
  uint64_t persistent = uint64_t(grid_sms(ctx)) * grid_occ(ctx, occ);
  if (EXPECT && persistent > kExpectMaxBlocks) persistent = kExpectMaxBlocks;
  const uint32_t blocks = (uint32_t) (need < persistent ? need : persistent);
  double* partials = nullptr;
  if constexpr (EXPECT) {
    rc = ensure_scratch(ctx, (2 * size_t{blocks} + 2) * sizeof(double));
    if (rc) return rc;
    partials = (double*) ctx->scratch;
  }
  kern<<<blocks, kDThreads, smem, ctx->stream>>>(st, g, (const double2*) dmat, partials);
  QB_LAUNCHED(ctx);
  stage_matrix_done(ctx);
  if constexpr (EXPECT) return finish_expectation(ctx, partials, blocks, out);
  return QB200_OK;
}

}  // namespace qb200
