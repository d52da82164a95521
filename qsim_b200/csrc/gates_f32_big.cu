// gates_f32_big.cu -- instantiates and launches the fp32 5- and 6-qubit kernels (gate_big.cuh).
#include <mutex>

#include "gate_big.cuh"
#include "gate_launch.cuh"

namespace qb200 {

namespace {

// c_mat6 is one __constant__ buffer per device: a pass that overwrites it must wait for the
// previous 6-qubit pass on that device (possibly issued on another context's stream).
constexpr int kMaxDevices = 64;
std::mutex g_mat6_mu;
cudaEvent_t g_mat6_done[kMaxDevices] = {};

template <int G, bool PAIR, bool EXPECT, int NT, int D>
int launch_big_shape(qb200_ctx* ctx, float* st, const TileGeom& t, const float* m, double* out) {
  auto kern = k_gate_big<G, PAIR, EXPECT, NT, D>;
  constexpr int warps = NT / 32;
  constexpr size_t smem = (size_t) warps * D * (8 << (G + 5)) + (8 << (G + 5));
  static PerDevice occ_cache;
  const int occ = occ_cache.get(ctx, [&] {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, NT, smem) != cudaSuccess || nb < 1) {
      (void) cudaGetLastError();
      nb = 1;
    }
    return nb;
  });
  const uint64_t need = (t.work + warps - 1) / warps;
  uint64_t persistent = uint64_t(grid_sms(ctx)) * grid_occ(ctx, occ);
  if (EXPECT && persistent > kExpectMaxBlocks) persistent = kExpectMaxBlocks;
  const uint32_t blocks = (uint32_t) (need < persistent ? need : persistent);

  double* partials = nullptr;
  if constexpr (EXPECT) {
    int rc = ensure_scratch(ctx, (2 * size_t{blocks} + 2) * sizeof(double));
    if (rc) return rc;
    partials = (double*) ctx->scratch;
  }

  if constexpr (G == 6) {
    if (ctx->device < 0 || ctx->device >= kMaxDevices) return QB200_ERR_INVALID;
    const void* dmat = nullptr;
    int rc = stage_matrix(ctx, m, sizeof(float) * (2 << 12), &dmat);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(g_mat6_mu);
    cudaEvent_t& ev = g_mat6_done[ctx->device];
    if (ev) QB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev, 0));
    else QB_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    QB_CUDA(ctx, cudaMemcpyToSymbolAsync(c_mat6, dmat, sizeof(float) * (2 << 12), 0,
                                         cudaMemcpyDeviceToDevice, ctx->stream));
    BigMat<6> tag{};
    kern<<<blocks, NT, smem, ctx->stream>>>(st, t, tag, partials);
    QB_LAUNCHED(ctx);
    QB_CUDA(ctx, cudaEventRecord(ev, ctx->stream));
    stage_matrix_done(ctx);
  } else {
    BigMat<G> mat;
    std::memcpy(mat.m, m, sizeof(mat.m));
    kern<<<blocks, NT, smem, ctx->stream>>>(st, t, mat, partials);
    QB_LAUNCHED(ctx);
  }
  if constexpr (EXPECT) return finish_expectation(ctx, partials, blocks, out);
  return QB200_OK;
}

template <int G, bool EXPECT>
int launch_big_g(qb200_ctx* ctx, float* st, const TileGeom& t, const float* m, double* out) {
  // launch shapes: 8 warps x 1 block per SM; ring depth D trades tile prefetch for warps
  // (G=6: 16 KB per tile stage, G=5: 8 KB)
  const int shape = ctx->tune.big;
  if (t.pair) {
    if (G == 5 && shape == 1) return launch_big_shape<G, true, EXPECT, 256, 1>(ctx, st, t, m, out);
    return launch_big_shape<G, true, EXPECT, 256, G == 5 ? 2 : 1>(ctx, st, t, m, out);
  }
  if (G == 5 && shape == 1) return launch_big_shape<G, false, EXPECT, 256, 1>(ctx, st, t, m, out);
  return launch_big_shape<G, false, EXPECT, 256, G == 5 ? 2 : 1>(ctx, st, t, m, out);
}

}  // namespace

int launch_big_f32(qb200_ctx* ctx, float* st, const TileGeom& t, unsigned nq, const float* m,
                   bool expect, double* out) {
  if (nq == 5) return expect ? launch_big_g<5, true>(ctx, st, t, m, out) : launch_big_g<5, false>(ctx, st, t, m, out);
  if (nq == 6) return expect ? launch_big_g<6, true>(ctx, st, t, m, out) : launch_big_g<6, false>(ctx, st, t, m, out);
  return QB200_ERR_UNSUPPORTED;
}

}  // namespace qb200
