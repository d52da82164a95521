// gate_launch.cuh -- host dispatch of one fused-gate pass (replaces the
// H/L/HH/LH/L switch ladders of lib/simulator_cuda.h:70-260).
#pragma once

#include <cstring>

#include "gate_kernels.cuh"

namespace qb200 {

// pinned/device ring for matrices that do not fit the kernel parameter space
int stage_matrix(qb200_ctx* ctx, const void* host, size_t bytes, const void** dev);
void stage_matrix_done(qb200_ctx* ctx);
int finish_expectation(qb200_ctx* ctx, double* partials, uint32_t blocks, double out[2]);

template <typename FP> struct RegLimits;
template <> struct RegLimits<float>  { static constexpr int kMaxG = 5; static constexpr int kMaxUnrollG = 4; };
template <> struct RegLimits<double> { static constexpr int kMaxG = 4; static constexpr int kMaxUnrollG = 3; };

// Launch shape per (precision, G): the unrolled G>=4 (fp32) / G>=3 (fp64)
// kernels want ~160 registers (data + one 64-bit address per group element),
// so they run 128-thread blocks, 3 per SM; smaller gates run 256 x 2.
template <typename FP, int G> constexpr int block_threads() {
  return (sizeof(FP) == 4 ? G <= 3 : G <= 2) ? 256 : 128;
}
template <typename FP, int G> constexpr int min_blocks() {
  return (sizeof(FP) == 4 ? G <= 3 : G <= 2) ? 2 : 3;
}

constexpr uint32_t kExpectMaxBlocks = kNumSMs * 8;

template <typename FP, int G, int MODE, bool EXPECT>
int launch_reg(qb200_ctx* ctx, FP* st, const Geom& g, const FP* m, double* out) {
  constexpr int NT = block_threads<FP, G>();
  constexpr int MINB = min_blocks<FP, G>();
  constexpr bool UNROLL = G <= RegLimits<FP>::kMaxUnrollG;
  using Mat = MatParam<FP, G>;
  Mat mat;
  std::memcpy(mat.m, m, sizeof(mat.m));
  uint64_t blocks64 = (g.work + NT - 1) / NT;
  if (blocks64 > 0x7fffffffull) blocks64 = 0x7fffffffull;
  if constexpr (EXPECT) {
    if constexpr (!UNROLL) {
      return QB200_ERR_UNSUPPORTED;
    } else {
      uint32_t blocks = (uint32_t) (blocks64 < kExpectMaxBlocks ? blocks64 : kExpectMaxBlocks);
      int rc = ensure_scratch(ctx, (2 * size_t{blocks} + 2) * sizeof(double));
      if (rc) return rc;
      double* partials = (double*) ctx->scratch;
      k_gate_reg<FP, G, MODE, true, true, NT, MINB, Mat><<<blocks, NT, 0, ctx->stream>>>(st, g, mat, partials);
      QB_LAUNCHED(ctx);
      return finish_expectation(ctx, partials, blocks, out);
    }
  } else {
    k_gate_reg<FP, G, MODE, UNROLL, false, NT, MINB, Mat><<<(uint32_t) blocks64, NT, 0, ctx->stream>>>(st, g, mat, nullptr);
    QB_LAUNCHED(ctx);
    return QB200_OK;
  }
}

template <typename FP, bool EXPECT>
int launch_generic(qb200_ctx* ctx, FP* st, const Geom& g, unsigned nq, const FP* m, double* out) {
  constexpr int NT = 64;
  const size_t mbytes = (size_t{2} << (2 * nq)) * sizeof(FP);
  const void* dmat = nullptr;
  int rc = stage_matrix(ctx, m, mbytes, &dmat);
  if (rc) return rc;
  const size_t smem = (size_t{2} << nq) * NT * sizeof(FP);
  auto kern = k_gate_generic<FP, EXPECT, NT>;
  if (smem > 48 * 1024) {
    QB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  }
  uint64_t blocks64 = (g.work + NT - 1) / NT;
  if (blocks64 > 0x7fffffffull) blocks64 = 0x7fffffffull;
  if constexpr (EXPECT) {
    uint32_t blocks = (uint32_t) (blocks64 < kExpectMaxBlocks ? blocks64 : kExpectMaxBlocks);
    rc = ensure_scratch(ctx, (2 * size_t{blocks} + 2) * sizeof(double));
    if (rc) return rc;
    double* partials = (double*) ctx->scratch;
    kern<<<blocks, NT, smem, ctx->stream>>>(st, g, (const FP*) dmat, partials);
    QB_LAUNCHED(ctx);
    stage_matrix_done(ctx);
    return finish_expectation(ctx, partials, blocks, out);
  } else {
    kern<<<(uint32_t) blocks64, NT, smem, ctx->stream>>>(st, g, (const FP*) dmat, nullptr);
    QB_LAUNCHED(ctx);
    stage_matrix_done(ctx);
    return QB200_OK;
  }
}

template <typename FP, int G, bool EXPECT>
int launch_reg_mode(qb200_ctx* ctx, int mode, FP* st, const Geom& g, const FP* m, double* out) {
  if constexpr (sizeof(FP) == 4) {
    if (mode == kV2) {
      if constexpr (G <= 4) return launch_reg<FP, G, kV2, EXPECT>(ctx, st, g, m, out);
    }
    if (mode == kV2T) {
      if constexpr (G >= 1) return launch_reg<FP, G, kV2T, EXPECT>(ctx, st, g, m, out);
    }
  }
  return launch_reg<FP, G, kV1, EXPECT>(ctx, st, g, m, out);
}

// One fused-gate pass: gate (EXPECT=false, in place) or expectation value.
template <typename FP, bool EXPECT>
int gate_pass(qb200_ctx* ctx, FP* st, unsigned n, const unsigned* qs, unsigned nq,
              const unsigned* cqs, unsigned nc, uint64_t cvals, const FP* m, double* out) {
  if (!ctx || !st || !m || (nq && !qs) || (nc && !cqs)) return QB200_ERR_INVALID;
  if (nq > kMaxTargets) return QB200_ERR_UNSUPPORTED;
  DeviceGuard guard(ctx);

  const bool aligned16 = (reinterpret_cast<uintptr_t>(st) & 15) == 0;
  bool generic = ctx->tune.force_generic || (int) nq > RegLimits<FP>::kMaxG ||
                 (EXPECT && (int) nq > RegLimits<FP>::kMaxUnrollG);
  int mode = kV1;
  if (!generic && sizeof(FP) == 4 && aligned16 && n >= 1 && ctx->tune.gate_mode != 0) {
    bool bit0_ctrl = false;
    for (unsigned j = 0; j < nc; ++j) bit0_ctrl |= cqs[j] == 0;
    if (nq >= 1 && qs[0] == 0) mode = kV2T;
    else if (!bit0_ctrl && nq <= 4 && nq + nc + 1 <= n) mode = kV2;
  }

  Geom g;
  int rc = make_geom(n, qs, nq, cqs, nc, cvals, mode == kV2, &g);
  if (rc) return rc;
  if (mode == kV2T) {
    // the pair (k, k+1) is moved by one thread: hide target 0 from the index
    // expansion is NOT needed (bit 0 is already a special position); nothing to do.
  }

  if (generic) return launch_generic<FP, EXPECT>(ctx, st, g, nq, m, out);

  switch (nq) {
    case 0: return launch_reg_mode<FP, 0, EXPECT>(ctx, mode, st, g, m, out);
    case 1: return launch_reg_mode<FP, 1, EXPECT>(ctx, mode, st, g, m, out);
    case 2: return launch_reg_mode<FP, 2, EXPECT>(ctx, mode, st, g, m, out);
    case 3: return launch_reg_mode<FP, 3, EXPECT>(ctx, mode, st, g, m, out);
    case 4: return launch_reg_mode<FP, 4, EXPECT>(ctx, mode, st, g, m, out);
    case 5:
      if constexpr (RegLimits<FP>::kMaxG >= 5 && !EXPECT)
        return launch_reg_mode<FP, 5, EXPECT>(ctx, mode, st, g, m, out);
      break;
    default: break;
  }
  return launch_generic<FP, EXPECT>(ctx, st, g, nq, m, out);
}

}  // namespace qb200
