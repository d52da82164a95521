// gate_launch.cuh -- host dispatch of one fused-gate pass (replaces the
// H/L/HH/LH/L switch ladders of lib/simulator_cuda.h:70-260).
#pragma once

#include <cstring>

#include "gate_tile.cuh"

namespace qb200 {

// pinned/device ring for matrices that do not fit the kernel parameter space
int stage_matrix(qb200_ctx* ctx, const void* host, size_t bytes, const void** dev);
void stage_matrix_done(qb200_ctx* ctx);
int finish_expectation(qb200_ctx* ctx, double* partials, uint32_t blocks, double out[2]);

// fp32 G = 5, 6 gate / expectation passes (gate_big.cuh, instantiated in gates_f32_big.cu)
int launch_big_f32(qb200_ctx* ctx, float* st, const TileGeom& t, unsigned nq, const float* m,
                   bool expect, double* out);

// fp32 G = 4, 5 gate passes on the tensor cores (gate_tc.cuh, instantiated in gates_f32_tc.cu)
int launch_tc_f32(qb200_ctx* ctx, float* st, const Geom& g, unsigned nq, bool pair, const float* m);
// fp32 G = 6 gates and G = 4, 5, 6 expectation values on the tensor cores (k_gate_tcx)
int launch_tcx_f32(qb200_ctx* ctx, float* st, const Geom& g, unsigned nq, bool pair, const float* m,
                   bool expect, double* out);

template <typename FP> struct RegLimits;
template <> struct RegLimits<float>  { static constexpr int kMaxG = 5; static constexpr int kMaxUnrollG = 4; };
template <> struct RegLimits<double> { static constexpr int kMaxG = 5; static constexpr int kMaxUnrollG = 5; };

// Launch shape per (precision, G): the unrolled G>=4 (fp32) / G>=3 (fp64)
// kernels want ~160 registers (data + one 64-bit address per group element),
// so they run 128-thread blocks, 3 per SM; smaller gates run 256 x 2.
template <typename FP, int G> constexpr int block_threads() {
  return (sizeof(FP) == 4 ? G <= 3 : G <= 2) ? 256 : 128;
}
template <typename FP, int G> constexpr int min_blocks() {
  if (sizeof(FP) == 8 && G >= 5) return 1;  // 2^5 complex doubles = 128 registers of data alone
  return (sizeof(FP) == 4 ? G <= 3 : G <= 2) ? 2 : 3;
}

constexpr uint32_t kExpectMaxBlocks = kNumSMs * 8;

// Largest shared-memory carveout for the warp-tile kernel: k_gate_tile<4> 3.43 -> 2.92 ms per pass at 30 qubits.  (The
// tensor-core kernels want the opposite: their per-thread 64-byte row loads rely on L1 sector merging, 2.82 -> 3.34 ms
// with the largest carveout.)
template <typename K>
inline void prefer_max_smem(K kern) {
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

// resident blocks per SM of a kernel (queried once per instantiation)
template <typename K>
int resident_blocks(K kern, int threads) {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, 0) != cudaSuccess || nb < 1) {
    (void) cudaGetLastError();
    nb = 1;
  }
  return nb;
}

template <int G, int MODE, int PNT, int PD, int PMINB>
int launch_pipe(qb200_ctx* ctx, float* st, const Geom& g, const MatParam<float, G>& mat) {
  auto kern = k_gate_pipe<G, MODE, PNT, PD, PMINB>;
  constexpr size_t smem = pipe_smem_bytes<G, MODE, PNT, PD>();
  static PerDevice occ_cache;
  const int occ = occ_cache.get(ctx, [&] {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, PNT, smem) != cudaSuccess || nb < 1) {
      (void) cudaGetLastError();
      nb = 1;
    }
    return nb;
  });
  const uint64_t need = (g.work + PNT - 1) / PNT;
  const uint64_t persistent = uint64_t(grid_sms(ctx)) * grid_occ(ctx, occ);
  const uint32_t blocks = (uint32_t) (need < persistent ? need : persistent);
  kern<<<blocks, PNT, smem, ctx->stream>>>(st, g, mat);
  QB_LAUNCHED(ctx);
  return QB200_OK;
}

template <int G, bool PAIR, int TNT, int TD, int TMINB>
int launch_tile(qb200_ctx* ctx, float* st, const TileGeom& t, const float* m) {
  MatParam<float, G> mat;
  mat.fill(m);
  auto kern = k_gate_tile<G, PAIR, TNT, TD, TMINB>;
  constexpr size_t smem = tile_smem_bytes<G, TNT, TD>();
  static PerDevice occ_cache;
  const int occ = occ_cache.get(ctx, [&] {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    prefer_max_smem(kern);
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, TNT, smem) != cudaSuccess || nb < 1) {
      (void) cudaGetLastError();
      nb = 1;
    }
    return nb;
  });
  constexpr int warps = TNT / 32;
  const uint64_t need = (t.work + warps - 1) / warps;
  const uint64_t persistent = uint64_t(grid_sms(ctx)) * grid_occ(ctx, occ);
  const uint32_t blocks = (uint32_t) (need < persistent ? need : persistent);
  kern<<<blocks, TNT, smem, ctx->stream>>>(st, t, mat);
  QB_LAUNCHED(ctx);
  return QB200_OK;
}

template <typename FP, int G, int MODE, bool EXPECT>
int launch_reg(qb200_ctx* ctx, FP* st, const Geom& g, const FP* m, double* out) {
  constexpr int NT = block_threads<FP, G>();
  constexpr int MINB = min_blocks<FP, G>();
  // fp64: rows are unrolled (matrix elements become constant-bank operands of the DFMAs) where that
  // measured faster at n = 29: G <= 3, G = 5 gates (11 ms against 16.5 ms through k_gate_dbig) and G = 4
  // expectation values (5.1 ms against 8.0 ms); the G = 4 gate keeps the rolled row loop (5.0 ms against 7.2 ms)
  constexpr bool UNROLL = sizeof(FP) == 8 ? (G <= 3 || G == 5 || EXPECT) : G <= RegLimits<FP>::kMaxUnrollG;
  using Mat = MatParam<FP, G>;
  Mat mat;
  mat.fill(m);
  uint64_t blocks64 = (g.work + NT - 1) / NT;
  if (blocks64 > 0x7fffffffull) blocks64 = 0x7fffffffull;
  if constexpr (EXPECT) {
    if constexpr (!UNROLL) {
      return QB200_ERR_UNSUPPORTED;
    } else {
      uint32_t blocks = (uint32_t) (blocks64 < kExpectMaxBlocks ? blocks64 : kExpectMaxBlocks);
      // Small read-only passes: fp32 G <= 2 runs k_expect_stream (the next iteration's loads issued before the
      // current group is consumed) on a persistent grid of four 256-thread blocks per SM.  One group per thread
      // per iteration is the default: several groups per iteration (tuning expect_ug = 2 grid-strided, 3
      // contiguous per block) measured 15-25 % slower at n = 26 and 30 in every layout
      // (profiles/r01_expect_variants.txt) -- the pass is not short of bytes in flight.
      if constexpr (sizeof(FP) == 4 && G <= 2) {
        constexpr int kLoads = MODE == kV2T ? (1 << G) / 2 : (1 << G);  // per group
        constexpr int kUG = kLoads >= 4 ? 1 : 4 / kLoads;
        auto run = [&](auto kern, int occ, int ug, int contiguous) -> int {
          const uint64_t persistent = uint64_t(grid_sms(ctx)) * grid_occ(ctx, occ);
          const uint64_t need = (blocks64 + ug - 1) / ug;
          const uint32_t nb = (uint32_t) (need < persistent ? need : persistent);
          int rc = ensure_scratch(ctx, (2 * size_t{nb} + 2) * sizeof(double));
          if (rc) return rc;
          double* partials = (double*) ctx->scratch;
          kern<<<nb, NT, 0, ctx->stream>>>(st, g, mat, partials, contiguous);
          QB_LAUNCHED(ctx);
          return finish_expectation(ctx, partials, nb, out);
        };
        static const int occ1 = resident_blocks(k_expect_stream<FP, G, MODE, 1, NT, 4, Mat>, NT);
        static const int occu = resident_blocks(k_expect_stream<FP, G, MODE, kUG, NT, 4, Mat>, NT);
        if (kUG == 1 || ctx->tune.expect_ug < 2) return run(k_expect_stream<FP, G, MODE, 1, NT, 4, Mat>, occ1, 1, 0);
        return run(k_expect_stream<FP, G, MODE, kUG, NT, 4, Mat>, occu, kUG, ctx->tune.expect_ug == 2 ? 0 : 1);
      }
      int rc = ensure_scratch(ctx, (2 * size_t{blocks} + 2) * sizeof(double));
      if (rc) return rc;
      double* partials = (double*) ctx->scratch;
      k_gate_reg<FP, G, MODE, true, true, false, NT, MINB, Mat><<<blocks, NT, 0, ctx->stream>>>(st, g, mat, partials);
      QB_LAUNCHED(ctx);
      return finish_expectation(ctx, partials, blocks, out);
    }
  } else {
    // fp32 G == 4: deep cp.async ring through shared memory (gate_pipe.cuh)
    if constexpr (sizeof(FP) == 4 && G == 4) {
      if (ctx->tune.tile != 0) {  // 0 = register kernels only
        switch (ctx->tune.block) {
          case 1: return launch_pipe<G, MODE, 128, (MODE == kV2 ? 3 : 6), 2>(ctx, st, g, mat);
          case 2: return launch_pipe<G, MODE, 256, (MODE == kV2 ? 2 : 4), 1>(ctx, st, g, mat);
          case 3: return launch_pipe<G, MODE, 128, (MODE == kV2 ? 2 : 4), 3>(ctx, st, g, mat);
          default: return launch_pipe<G, MODE, 256, (MODE == kV2 ? 2 : 3), 2>(ctx, st, g, mat);
        }
      }
    }
    // G >= 4 fp32 (single group per thread): software-pipelined persistent grid
    constexpr bool kCanPrefetch = sizeof(FP) == 4 && G >= 4 && UNROLL && MODE != kV2;
    if constexpr (kCanPrefetch) {
      if (ctx->tune.prefetch != 0) {
        auto kern = k_gate_reg<FP, G, MODE, UNROLL, false, true, NT, MINB, Mat>;
        static PerDevice occ_cache;
        const int occ = occ_cache.get(ctx, [&] { return resident_blocks(kern, NT); });
        const uint64_t persistent = uint64_t(grid_sms(ctx)) * grid_occ(ctx, occ);
        const uint32_t blocks = (uint32_t) (blocks64 < persistent ? blocks64 : persistent);
        kern<<<blocks, NT, 0, ctx->stream>>>(st, g, mat, nullptr);
        QB_LAUNCHED(ctx);
        return QB200_OK;
      }
    }
    k_gate_reg<FP, G, MODE, UNROLL, false, false, NT, MINB, Mat><<<(uint32_t) blocks64, NT, 0, ctx->stream>>>(st, g, mat, nullptr);
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

}  // namespace qb200

#include "gate_dbig.cuh"

namespace qb200 {

// One fused-gate pass: gate (EXPECT=false, in place) or expectation value.
template <typename FP, bool EXPECT>
int gate_pass(qb200_ctx* ctx, FP* st, unsigned n, const unsigned* qs, unsigned nq,
              const unsigned* cqs, unsigned nc, uint64_t cvals, const FP* m, double* out) {
  if (!ctx || !st || !m || (nq && !qs) || (nc && !cqs)) return QB200_ERR_INVALID;
  if (nq > kMaxTargets) return QB200_ERR_UNSUPPORTED;
  DeviceGuard guard(ctx);
  if (int drc = check_state_device(ctx, st)) return drc;
  if constexpr (!EXPECT) note_state_written();

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

  // fp32 G = 4, 5 gates: tcgen05 3xTF32 kernel (gate_tc.cuh); tiles of 128 groups
  if constexpr (sizeof(FP) == 4 && !EXPECT) {
    // auto (-1): every G = 5 pass (the tensor-core kernel wins for all 32 low-bit layouts, 0.33-0.9x the time of
    // k_gate_big<5>) and every G = 4 pass except the layouts with at least two targets on bits 1, 2, 3: what decides
    // is which targets sit inside the warp's lane bits 0..4 (the per-thread row pieces then form short runs across
    // neighbouring lanes), not the qubit count of the state -- the same 15 of 31 masks lose by 2-110 % at n = 26, 30
    // and 33 (tools/dispatch_sweep.py, profiles/r02_dispatch_sweep.txt) and go to the warp-tile FFMA2 kernel below.
    // tc_low >= 0 replaces the table by a plain threshold on the lowest non-zero target (experiments).
    bool g4_tc = false;
    if (nq == 4) {
      if (ctx->tune.tc_low >= 0) {
        g4_tc = (int) (qs[0] == 0 ? qs[1] : qs[0]) >= ctx->tune.tc_low;
      } else {
        unsigned low_mask = 0;
        for (unsigned j = 0; j < nq; ++j)
          if (qs[j] < 5) low_mask |= 1u << qs[j];
        g4_tc = __builtin_popcount(low_mask & 0b01110u) < 2;
      }
    }
    const bool use_tc = ctx->tune.tc > 0 || (ctx->tune.tc < 0 && (nq == 5 || g4_tc));
    if (!ctx->tune.force_generic && (nq == 4 || nq == 5) && aligned16 && use_tc && n >= nq + nc + 7) {
      Geom tg;
      int trc = make_geom(n, qs, nq, cqs, nc, cvals, false, &tg);
      if (trc) return trc;
      ctx->last_kernel = ctx->tune.tc > 0 && ctx->tune.tc < 3 ? (nq == 4 ? "k_gate_tc<4>" : "k_gate_tc<5>")
                                                                : (nq == 4 ? "k_gate_tca<4>" : "k_gate_tca<5>");
      return launch_tc_f32(ctx, st, tg, nq, qs[0] == 0, m);
    }
  }

  // fp32 G == 4: warp-cooperative swizzled-tile kernel (gate_tile.cuh) when the lowest
  // target sits below bit 3 (the per-thread kernels would issue 16-byte chunks at a
  // 128-byte lane stride there), or everywhere when forced by tuning tile=2.
  if constexpr (sizeof(FP) == 4 && !EXPECT) {
    if (!generic && nq == 4 && aligned16 && (ctx->tune.tile == 2 || (ctx->tune.tile == -1 && qs[0] <= 2))) {
      TileGeom t;
      int trc = make_tile_geom(n, qs, nq, cqs, nc, cvals, &t);
      if (trc == QB200_OK) {
        ctx->last_kernel = "k_gate_tile<4>";
        return t.pair ? launch_tile<4, true, 256, 3, 2>(ctx, st, t, m)
                      : launch_tile<4, false, 256, 3, 2>(ctx, st, t, m);
      }
      if (trc != QB200_ERR_UNSUPPORTED) return trc;
    }
  }

  // fp32 G = 6 gates and G = 4, 5, 6 expectation values: tensor cores too (k_gate_tcx); tuning tcx = 0 keeps
  // them on the FFMA2 kernels below
  if constexpr (sizeof(FP) == 4) {
    const bool want = EXPECT ? (nq >= 4 && nq <= 6) : nq == 6;
    if (!ctx->tune.force_generic && want && aligned16 && ctx->tune.tc != 0 && ctx->tune.tcx != 0 && nc == 0 &&
        n >= nq + 7) {
      Geom tg;
      int trc = make_geom(n, qs, nq, cqs, nc, cvals, false, &tg);
      if (trc) return trc;
      ctx->last_kernel = EXPECT ? (nq == 4 ? "k_gate_tcx<4,expect>" : nq == 5 ? "k_gate_tcx<5,expect>" : "k_gate_tcx<6,expect>")
                                : "k_gate_tcx<6>";
      return launch_tcx_f32(ctx, st, tg, nq, qs[0] == 0, m, EXPECT, out);
    }
  }

  // fp64 G = 4, 5, 6 gates and expectation values: register-blocked complex GEMM (gate_dbig.cuh k_gate_d6) once the
  // pass has whole tiles of 2^(12 - G) groups.  Tuning big: 1 = the round-1 selection (G = 4, 5 gates and G = 4
  // expectation values on the unrolled register kernels with constant-bank matrix operands, G = 5 expectation values
  // on k_gate_dbig, G = 6 on k_gate_d6); 2 = k_gate_dbig for G = 5 gates and G = 4 expectation values too; 3 = like 1
  // with k_gate_dbig for G = 6 (the cross-check in the tests); 0 = none of these kernels.
  if constexpr (sizeof(FP) == 8) {
    if (!ctx->tune.force_generic && nq >= 4 && nq <= 6 && ctx->tune.big != 0) {
      Geom dg;
      int drc = make_geom(n, qs, nq, cqs, nc, cvals, false, &dg);
      if (drc) return drc;
      const int big = ctx->tune.big;
      const bool gemm = big == -1 || (nq == 6 && big != 3);
      if (gemm) {
        ctx->last_kernel = "k_gate_d6";
        // (G = 4 with bit 0 among the targets: group elements 16 bytes apart, the per-group loads do not coalesce --
        //  5.9-7.6 ms against 5.0-5.6 on the register kernels)
        if (nq == 4 && d6_fits<4>(dg) && qs[0] >= 1) return launch_d6<4, EXPECT>(ctx, st, dg, m, out);
        if (nq == 5 && d6_fits<5>(dg)) return launch_d6<5, EXPECT>(ctx, st, dg, m, out);
        if (nq == 6 && d6_fits<6>(dg)) return launch_d6<6, EXPECT>(ctx, st, dg, m, out);
      }
      const bool want = nq == 6 || (EXPECT && nq == 5) || (big == 2 && (EXPECT ? nq >= 4 : nq == 5));
      if (want) {
        ctx->last_kernel = "k_gate_dbig";
        if (nq == 4) { if constexpr (EXPECT) return launch_dbig<4, true>(ctx, st, dg, m, out); }
        if (nq == 5) return launch_dbig<5, EXPECT>(ctx, st, dg, m, out);
        if (nq == 6) return launch_dbig<6, EXPECT>(ctx, st, dg, m, out);
      }
    }
  }

  // fp32 G = 5, 6 (gate and expectation): row-blocked warp-tile kernel (gate_big.cuh)
  if constexpr (sizeof(FP) == 4) {
    if (!ctx->tune.force_generic && (nq == 5 || nq == 6) && aligned16 && ctx->tune.big != 0) {
      TileGeom t;
      int trc = make_tile_geom(n, qs, nq, cqs, nc, cvals, &t);
      if (trc == QB200_OK) {
        ctx->last_kernel = nq == 5 ? "k_gate_big<5>" : "k_gate_big<6>";
        return launch_big_f32(ctx, st, t, nq, m, EXPECT, out);
      }
      if (trc != QB200_ERR_UNSUPPORTED) return trc;
    }
  }

  Geom g;
  int rc = make_geom(n, qs, nq, cqs, nc, cvals, mode == kV2, &g);
  if (rc) return rc;
  if (mode == kV2T) {
    // the pair (k, k+1) is moved by one thread: hide target 0 from the index
    // expansion is NOT needed (bit 0 is already a special position); nothing to do.
  }

  if (generic) {
    ctx->last_kernel = "k_gate_generic";
    return launch_generic<FP, EXPECT>(ctx, st, g, nq, m, out);
  }

  {
    static const char* const kRegNames[2][6] = {
        {"k_gate_reg<0>", "k_gate_reg<1>", "k_gate_reg<2>", "k_gate_reg<3>", "k_gate_reg<4>", "k_gate_reg<5>"},
        {"k_expect<0>", "k_expect<1>", "k_expect<2>", "k_expect<3>", "k_expect<4>", "k_expect<5>"}};
    if (nq <= 5) ctx->last_kernel = kRegNames[EXPECT ? 1 : 0][nq];
    if (sizeof(FP) == 4 && !EXPECT && nq == 4 && ctx->tune.tile != 0) ctx->last_kernel = "k_gate_pipe<4>";
  }
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
  ctx->last_kernel = "k_gate_generic";
  return launch_generic<FP, EXPECT>(ctx, st, g, nq, m, out);
}

}  // namespace qb200
