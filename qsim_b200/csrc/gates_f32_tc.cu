// gates_f32_tc.cu -- instantiates and launches the tcgen05 (3xTF32) gate kernels (gate_tc.cuh).
#include <cstdio>
#include <cstdlib>

#include "gate_tc.cuh"
#include "gate_launch.cuh"

namespace qb200 {

namespace {

// At most one non-zero per row and per column: permutations, Pauli strings, diagonal phases, CZ/SWAP-like fused
// gates.  Every output amplitude is then ONE product, there is no accumulation whose truncation the compensation
// term was calibrated to cancel (tc_bias: mean norm drift of DENSE unitaries), so for these matrices the term
// would inflate the norm by ~1e-7 per pass.  Without it a permutation with entries in {0, +-1, +-i} is exact on the
// tensor cores (A_hi + A_lo == A in fp32).
static bool is_monomial(const float* m, unsigned nq) {
  const unsigned dim = 1u << nq;
  unsigned col_used[64] = {};
  for (unsigned r = 0; r < dim; ++r) {
    unsigned nz = 0;
    for (unsigned c = 0; c < dim; ++c) {
      if (m[2 * (r * dim + c)] != 0.f || m[2 * (r * dim + c) + 1] != 0.f) {
        if (++nz > 1 || col_used[c]++) return false;
      }
    }
  }
  return true;
}

template <int G, bool PAIR, int NBUF, int PF, int MINB>
int launch_tc_shape(qb200_ctx* ctx, float* st, const Geom& g, const float* m) {
  auto kern = k_gate_tc<G, PAIR, NBUF, PF, MINB>;
  constexpr size_t smem = tc_smem_bytes<G, NBUF>();
  static PerDevice occ_cache;
  const int occ = occ_cache.get(ctx, [&] {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    // resident CTAs per SM: shared memory (227 KB usable, 1 KB reserved per CTA) and TMEM
    // (512 columns) are the limits; the runtime's occupancy query answers 1 for kernels that
    // allocate tensor memory, so it is not used here
    int nb = (int) ((227 * 1024) / (smem + 1024 + 64));
    const int tmem_limit = 512 / (NBUF * TcShape<G>::KF);
    if (nb > tmem_limit) nb = tmem_limit;
    if (nb > MINB) nb = MINB;
    if (nb < 1) nb = 1;
    if (getenv("QB200_VERBOSE")) fprintf(stderr, "k_gate_tc<%d>: smem %zu, blocks per SM %d\n", G, smem, nb);
    return nb;
  });
  MatParam<float, G> mat;
  mat.fill(m);
  const uint64_t tiles = g.work >> 7;
  const uint64_t persistent = uint64_t(grid_sms(ctx)) * grid_occ(ctx, occ);
  const uint32_t blocks = (uint32_t) (tiles < persistent ? tiles : persistent);
  kern<<<blocks, kTcThreads, smem, ctx->stream>>>(st, g, mat);
  QB_LAUNCHED(ctx);
  return QB200_OK;
}


template <int G, bool PAIR, int NBUF, int PF, int MT, bool COMP, int RING, int MINB>
int launch_tca_shape(qb200_ctx* ctx, float* st, const Geom& g, const float* m) {
  auto kern = k_gate_tca<G, PAIR, NBUF, PF, MT, COMP, RING, MINB>;
  constexpr size_t smem = tca_smem_bytes<G, RING, MT>();
  static PerDevice occ_cache;
  const int occ = occ_cache.get(ctx, [&] {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    // (no shared-memory carveout preference: with the largest carveout the per-thread 64-byte row loads lose their
    //  L1 sector merging, 2.82 -> 3.34 ms per pass at 30 qubits)
    int nb = 512 / tca_tmem_cols<G, NBUF * MT>();  // tensor memory (and 2048 threads) bound the residency
    if (nb * kTcThreads * MT > 2048) nb = 2048 / (kTcThreads * MT);
    const int smem_limit = (int) ((227 * 1024) / (smem + 1024 + 64));
    if (nb > smem_limit) nb = smem_limit;
    if (nb > MINB) nb = MINB;
    if (nb < 1) nb = 1;
    if (getenv("QB200_VERBOSE")) fprintf(stderr, "k_gate_tca<%d>: smem %zu, blocks per SM %d\n", G, smem, nb);
    return nb;
  });
  MatParam<float, G> mat;
  mat.fill(m);
  const uint64_t tiles = g.work >> 7 >> (MT - 1);
  const uint64_t persistent = uint64_t(grid_sms(ctx)) * grid_occ(ctx, occ);
  const uint32_t blocks = (uint32_t) (tiles < persistent ? tiles : persistent);
  kern<<<blocks, kTcThreads * MT, smem, ctx->stream>>>(st, g, mat);
  QB_LAUNCHED(ctx);
  return QB200_OK;
}


// G = 6 gates and tensor-core expectation values (k_gate_tcx): one CTA per SM, matrix from the
// pinned->device staging ring
template <int G, bool PAIR, bool EXPECT>
int launch_tcx_shape(qb200_ctx* ctx, float* st, const Geom& g, const float* m, double* out) {
  auto kern = k_gate_tcx<G, PAIR, EXPECT>;
  constexpr size_t smem = tcx_smem_bytes<G>();
  // resident CTAs per SM: TMEM columns (the allocation is the next power of two of 3 * KF), shared
  // memory and registers (the runtime's occupancy query answers 1 for kernels that allocate TMEM)
  static PerDevice occ_cache;
  const int occ = occ_cache.get(ctx, [&] {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    int nb = 512 / tca_tmem_cols<G, 1>();
    const int smem_limit = (int) ((227 * 1024) / (smem + 1024 + 64));
    if (nb > smem_limit) nb = smem_limit;
    cudaFuncAttributes fa{};
    if (cudaFuncGetAttributes(&fa, kern) == cudaSuccess && fa.numRegs > 0) {
      const int reg_limit = 65536 / (((fa.numRegs + 7) / 8 * 8) * kTcThreads);
      if (nb > reg_limit) nb = reg_limit;
    }
    if (nb < 1) nb = 1;
    if (getenv("QB200_VERBOSE")) fprintf(stderr, "k_gate_tcx<%d,%d>: smem %zu, blocks per SM %d\n", G, (int) EXPECT, smem, nb);
    return nb;
  });
  const uint64_t tiles = g.work >> 7;
  uint64_t persistent = uint64_t(grid_sms(ctx)) * grid_occ(ctx, occ);
  if (EXPECT && persistent > kExpectMaxBlocks) persistent = kExpectMaxBlocks;
  const uint32_t blocks = (uint32_t) (tiles < persistent ? tiles : persistent);
  double* partials = nullptr;
  if constexpr (EXPECT) {
    int rc = ensure_scratch(ctx, (2 * size_t{blocks} + 2) * sizeof(double));
    if (rc) return rc;
    partials = (double*) ctx->scratch;
  }
  const void* dmat = nullptr;
  int rc = stage_matrix(ctx, m, sizeof(float) * (size_t{2} << (2 * G)), &dmat);
  if (rc) return rc;
  // accumulation-bias compensation (see tc_bias): measured per G; none for expectation values
  const float comp = EXPECT || is_monomial(m, G) ? 0.f
                   : (G == 4 ? tc_bias<4>() : G == 5 ? tc_bias<5>() : (float) ctx->tune.tc_comp6 * 1e-9f);
  kern<<<blocks, kTcThreads, smem, ctx->stream>>>(st, g, (const float*) dmat, comp, partials);
  QB_LAUNCHED(ctx);
  stage_matrix_done(ctx);
  if constexpr (EXPECT) return finish_expectation(ctx, partials, blocks, out);
  return QB200_OK;
}

}  // namespace

int launch_tcx_f32(qb200_ctx* ctx, float* st, const Geom& g, unsigned nq, bool pair, const float* m,
                   bool expect, double* out) {
  if (expect) {
    if (nq == 4) return pair ? launch_tcx_shape<4, true, true>(ctx, st, g, m, out) : launch_tcx_shape<4, false, true>(ctx, st, g, m, out);
    if (nq == 5) return pair ? launch_tcx_shape<5, true, true>(ctx, st, g, m, out) : launch_tcx_shape<5, false, true>(ctx, st, g, m, out);
    if (nq == 6) return pair ? launch_tcx_shape<6, true, true>(ctx, st, g, m, out) : launch_tcx_shape<6, false, true>(ctx, st, g, m, out);
  } else if (nq == 6) {
    return pair ? launch_tcx_shape<6, true, false>(ctx, st, g, m, out) : launch_tcx_shape<6, false, false>(ctx, st, g, m, out);
  }
  return QB200_ERR_UNSUPPORTED;
}

int launch_tc_f32(qb200_ctx* ctx, float* st, const Geom& g, unsigned nq, bool pair, const float* m) {
  const bool alt = ctx->tune.tc == 2;  // alternative shapes (tools/tc_check.py)
  // default: A operand in tensor memory, bias-compensated -- except for matrices with one non-zero per row
  // (is_monomial): plain 3xTF32, which is exact for permutations
  const bool plain = ctx->tune.tc == 4 || ((ctx->tune.tc < 0 || ctx->tune.tc == 3) && is_monomial(m, nq));
  if (!plain && (ctx->tune.tc < 0 || ctx->tune.tc == 3)) {
    if (nq == 4) return pair ? launch_tca_shape<4, true, 1, 2, 1, true, 0, 3>(ctx, st, g, m) : launch_tca_shape<4, false, 1, 2, 1, true, 0, 3>(ctx, st, g, m);
    if (nq == 5) return pair ? launch_tca_shape<5, true, 1, 1, 1, true, 0, 2>(ctx, st, g, m) : launch_tca_shape<5, false, 1, 1, 1, true, 0, 2>(ctx, st, g, m);
  }
  if (plain) {  // ... without the compensation term (plain 3xTF32)
    if (nq == 4) return pair ? launch_tca_shape<4, true, 1, 2, 1, false, 0, 3>(ctx, st, g, m) : launch_tca_shape<4, false, 1, 2, 1, false, 0, 3>(ctx, st, g, m);
    if (nq == 5) return pair ? launch_tca_shape<5, true, 1, 1, 1, false, 0, 2>(ctx, st, g, m) : launch_tca_shape<5, false, 1, 1, 1, false, 0, 2>(ctx, st, g, m);
  }
  if (ctx->tune.tc == 5) {  // ... cp.async staging ring, 3 tiles ahead, four CTAs per SM
    if (nq == 4) return pair ? launch_tca_shape<4, true, 1, 1, 1, true, 3, 4>(ctx, st, g, m) : launch_tca_shape<4, false, 1, 1, 1, true, 3, 4>(ctx, st, g, m);
    if (nq == 5) return pair ? launch_tca_shape<5, true, 1, 1, 1, true, 3, 2>(ctx, st, g, m) : launch_tca_shape<5, false, 1, 1, 1, true, 3, 2>(ctx, st, g, m);
  }
  if (ctx->tune.tc == 6) {  // ... ring of 4, three CTAs per SM
    if (nq == 4) return pair ? launch_tca_shape<4, true, 1, 1, 1, true, 4, 3>(ctx, st, g, m) : launch_tca_shape<4, false, 1, 1, 1, true, 4, 3>(ctx, st, g, m);
    if (nq == 5) return pair ? launch_tca_shape<5, true, 1, 1, 1, true, 2, 2>(ctx, st, g, m) : launch_tca_shape<5, false, 1, 1, 1, true, 2, 2>(ctx, st, g, m);
  }
  if (nq == 4) {
    if (alt) {
      return pair ? launch_tc_shape<4, true, 1, 2, 4>(ctx, st, g, m) : launch_tc_shape<4, false, 1, 2, 4>(ctx, st, g, m);
    }
    return pair ? launch_tc_shape<4, true, 2, 2, 3>(ctx, st, g, m) : launch_tc_shape<4, false, 2, 2, 3>(ctx, st, g, m);
  }
  if (nq == 5) {
    if (alt) {
      return pair ? launch_tc_shape<5, true, 2, 2, 1>(ctx, st, g, m) : launch_tc_shape<5, false, 2, 2, 1>(ctx, st, g, m);
    }
    return pair ? launch_tc_shape<5, true, 1, 1, 2>(ctx, st, g, m) : launch_tc_shape<5, false, 1, 1, 2>(ctx, st, g, m);
  }
  return QB200_ERR_UNSUPPORTED;
}

}  // namespace qb200
