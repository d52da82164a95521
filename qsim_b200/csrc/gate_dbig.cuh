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

// ---------------------------------------------------------------------------------------------------------
// k_gate_d6: fp64 6-qubit gates and expectation values as a register-blocked complex GEMM.
//
// k_gate_dbig<6> keeps one group per thread in 64 KB of shared memory per 64 threads: 6 warps per SM, one
// warp-uniform LDG per complex MAC -- 53 ms at 29 qubits, 7x above the DFMA bound (8 * 64 * 2^n flop at ~37 TFLOP/s
// = 7.4 ms).  Here a CTA of 256 threads multiplies the 64 x 64 matrix (64 KB of shared memory, staged once) with a
// tile of 64 groups (64 KB, [element][group], filled by cp.async, double-buffered so that the next tile's loads run
// under this tile's 1 M DFMAs); thread (ty, tx) owns rows ty + 16 r and groups tx + 16 q, r, q = 0..3: 16 complex
// accumulators = 64 independent DFMA chains, and per matrix column 4 + 4 LDS.128 feed 64 DFMAs (the matrix loads
// are two-address broadcasts, the state loads 16 consecutive double2 per half warp: conflict free).  Results go
// straight from registers to global memory (16 consecutive groups per store instruction).
// ---------------------------------------------------------------------------------------------------------
constexpr int kD6Threads = 256;
constexpr size_t kD6Tile = 64 * 64 * sizeof(double2);   // one state tile: 2^G elements x 4096 / 2^G groups = 64 KB for every G

// G = 6 as described; G = 5 / 4: the same kernel with 8 / 4 row slots per column of threads, 32 / 64 thread columns and
// tiles of 128 / 256 groups (the state tile stays 64 KB, the matrix shrinks to 16 / 4 KB).
//
// (Tried and dropped: touching global memory in address order whatever the targets are -- tile amplitudes enumerated
// by their 12 index bits sorted by position, results written back through the tile buffer -- costs two more barriers
// and bank-conflicted shared-memory traffic: 3.89 -> 4.51 ms for G = 4, 11.4 -> 12.1 ms for G = 6, low targets no better.)
template <int G, bool EXPECT>
__global__ void __launch_bounds__(kD6Threads, 1)
k_gate_d6(double* __restrict__ st, const __grid_constant__ Geom g, const double2* __restrict__ mat,
          double* __restrict__ partials) {
  constexpr int N = 1 << G;            // rows = columns = elements per group
  constexpr int TY = N / 4;            // thread rows: thread (ty, tx) owns rows ty + TY r, r = 0..3
  constexpr int TX = kD6Threads / TY;  // thread columns: ... and groups tx + TX q, q = 0..3
  constexpr int GT = 4 * TX;           // groups per tile
  constexpr int GTB = 12 - G;          // log2(GT)
  static_assert(N * GT == 4096, "a tile is 4096 amplitudes");
  extern __shared__ __align__(16) unsigned char d6_raw[];
  double2* const X = reinterpret_cast<double2*>(d6_raw);   // two buffers of [element][group]
  double2* const M = X + 2 * 4096;                          // [row][col]
  const uint32_t t = threadIdx.x, tx = t % TX, ty = t / TX;
  for (uint32_t e = t; e < N * N; e += kD6Threads) M[e] = mat[e];
  const uint64_t ntiles = g.work >> GTB;   // the launcher guarantees whole tiles

  // loads: thread t always fetches group t % GT of the tile, elements t / GT + (256 / GT) j
  const uint32_t lgrp = t % GT, lk0 = t / GT;
  constexpr uint32_t kStep = kD6Threads / GT > 0 ? kD6Threads / GT : 1;
  auto issue = [&](uint64_t tile, uint32_t b) {
    const double* p = st + 2 * expand_index((tile << GTB) + lgrp, g);
    const uint32_t dst = (uint32_t) __cvta_generic_to_shared(X + b * 4096 + lgrp);
#pragma unroll 4
    for (uint32_t j = 0; j < 16; ++j) {
      const uint32_t k = lk0 + kStep * j;
      uint64_t o = 0;
#pragma unroll
      for (int q = 0; q < G; ++q)
        if ((k >> q) & 1) o += g.xs[q];
      cp_async16(dst + k * GT * (uint32_t) sizeof(double2), p + 2 * o);
    }
    cp_async_commit();
  };
  double ere = 0, eim = 0;
  uint64_t tile = blockIdx.x;
  if (tile < ntiles) issue(tile, 0);
  for (uint32_t it = 0; tile < ntiles; tile += gridDim.x, ++it) {
    const uint32_t b = it & 1;
    if (tile + gridDim.x < ntiles) {
      issue(tile + gridDim.x, b ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();   // this tile (and, first time round, the matrix) is in shared memory
    const double2* const Xb = X + b * 4096;
    double ar[4][4], ai[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int q = 0; q < 4; ++q) ar[r][q] = ai[r][q] = 0.0;
#pragma unroll 2
    for (int c = 0; c < N; ++c) {
      double2 m[4], x[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) m[r] = M[(ty + TY * r) * N + c];
#pragma unroll
      for (int q = 0; q < 4; ++q) x[q] = Xb[c * GT + tx + TX * q];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          ar[r][q] = fma(x[q].x, m[r].x, ar[r][q]);
          ar[r][q] = fma(-x[q].y, m[r].y, ar[r][q]);
          ai[r][q] = fma(x[q].x, m[r].y, ai[r][q]);
          ai[r][q] = fma(x[q].y, m[r].x, ai[r][q]);
        }
    }
    if constexpr (EXPECT) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double2 xv = Xb[(ty + TY * r) * GT + tx + TX * q];
          ere += xv.x * ar[r][q] + xv.y * ai[r][q];
          eim += xv.x * ai[r][q] - xv.y * ar[r][q];
        }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        double* const p = st + 2 * expand_index((tile << GTB) + tx + TX * q, g);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const uint32_t k = ty + TY * r;
          uint64_t o = 0;
#pragma unroll
          for (int j = 0; j < G; ++j)
            if ((k >> j) & 1) o += g.xs[j];
          *reinterpret_cast<double2*>(p + 2 * o) = make_double2(ar[r][q], ai[r][q]);
        }
      }
    }
    __syncthreads();   // everybody is done with buffer b before the next iteration's loads overwrite it
  }
  if constexpr (EXPECT) {
    block_sum2<kD6Threads>(ere, eim);
    if (t == 0) {
      partials[2 * blockIdx.x] = ere;
      partials[2 * blockIdx.x + 1] = eim;
    }
  }
}

// whole tiles of 2^(12 - G) groups?
template <int G>
inline bool d6_fits(const Geom& g) { return g.work >= (uint64_t{1} << (12 - G)); }

template <int G, bool EXPECT>
int launch_d6(qb200_ctx* ctx, double* st, const Geom& g, const double* m, double* out) {
  auto kern = k_gate_d6<G, EXPECT>;
  constexpr size_t smem = 2 * kD6Tile + (size_t{1} << (2 * G)) * sizeof(double2);
  static PerDevice attr;
  attr.get(ctx, [&] {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    return 1;
  });
  const void* dmat = nullptr;
  int rc = stage_matrix(ctx, m, sizeof(double) * (size_t{2} << (2 * G)), &dmat);
  if (rc) return rc;
  const uint64_t tiles = g.work >> (12 - G);
  uint64_t persistent = uint64_t(grid_sms(ctx));
  if (EXPECT && persistent > kExpectMaxBlocks) persistent = kExpectMaxBlocks;
  const uint32_t blocks = (uint32_t) (tiles < persistent ? tiles : persistent);
  double* partials = nullptr;
  if constexpr (EXPECT) {
    rc = ensure_scratch(ctx, (2 * size_t{blocks} + 2) * sizeof(double));
    if (rc) return rc;
    partials = (double*) ctx->scratch;
  }
  kern<<<blocks, kD6Threads, smem, ctx->stream>>>(st, g, (const double2*) dmat, partials);
  QB_LAUNCHED(ctx);
  stage_matrix_done(ctx);
  if constexpr (EXPECT) return finish_expectation(ctx, partials, blocks, out);
  return QB200_OK;
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
