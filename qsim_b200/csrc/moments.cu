// moments.cu -- qb200_one_qubit_moments: the reduced density matrix of EVERY qubit from a handful of
// read-only passes (the observable loop of lib/expect.h:106-151 / apps/qsim_qtrajectory_cuda.cu
// costs one full pass per single-qubit operator; 52 of the 56 passes per trajectory of BASELINE config 5).
//
// For qubit q:  S00 = sum |a_i|^2 over bit_q(i) = 0,  S11 = the same over bit_q(i) = 1,
//               S01 = sum conj(a_i0) * a_i1 over the pairs (i0, i1 = i0 | 1 << q),
// and any single-qubit operator M has  <psi|M|psi> = m00 S00 + m11 S11 + m01 S01 + m10 conj(S01).
//
// One pass = one set B of T tile bits (T = 12 fp32 / 11 fp64: a 32 KB tile): the low 4 index bits (128-byte
// runs -> fully coalesced 128-bit loads) plus up to T - 4 further bit positions.  A block stages the 2^T
// amplitudes whose other index bits equal the tile number in shared memory (next tile prefetched into
// registers meanwhile) and each warp walks the 2^(T-1) pairs of "its" measured tile bits: products in FP,
// per-tile sums in FP, running sums in double (the reference's precision contract, lib/simulator_basic.h:323-324),
// fixed reduction tree.  Pass 0 measures bits 0..T-1, every later pass T - 4 new qubits: 3 passes at 26 qubits,
// 4 at 30 -- against one pass per operator.  Bound: pass 0 by shared-memory bandwidth (12 x 32 KB read per
// 32 KB loaded, about 2.3x the HBM time of a read pass), the later passes by HBM.
#include <algorithm>

#include "gate_kernels.cuh"

namespace qb200 {

constexpr int kMomNT = 256;
constexpr int kMomMaxT = 12;

struct MomGeom {
  uint32_t T = 0;                // tile bits
  uint32_t pos[kMomMaxT] = {};   // their positions in the amplitude index, ascending
  uint32_t nm = 0;               // measured tile bits in this pass
  uint32_t mk[kMomMaxT] = {};    // tile-bit number of each
  uint32_t qubit[kMomMaxT] = {};  // = pos[mk[.]]
  uint64_t ntiles = 0;
};

template <typename FP, int TMAX>
__global__ void __launch_bounds__(kMomNT, 2)
k_moments(const FP* __restrict__ st, const __grid_constant__ MomGeom g, double* __restrict__ partials) {
  using V2 = typename Vec2<FP>::type;
  constexpr int V = 16 / (int) sizeof(V2);                // amplitudes per 128-bit access
  constexpr int NLD = (1 << TMAX) / V / kMomNT;            // 128-bit loads per thread per tile
  __shared__ __align__(16) V2 s[1 << TMAX];
  const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t tile_amps = 1u << g.T;
  const uint32_t nvec = tile_amps / V;

  auto deposit = [&](uint32_t j) {
    uint64_t o = 0;
    for (uint32_t k = 0; k < g.T; ++k) o |= uint64_t((j >> k) & 1u) << g.pos[k];
    return o;
  };
  auto tile_base = [&](uint64_t t) {
    for (uint32_t k = 0; k < g.T; ++k) {
      const uint64_t lo = t & ((uint64_t{1} << g.pos[k]) - 1);
      t = ((t - lo) << 1) | lo;
    }
    return t;
  };
  uint64_t off[NLD];
#pragma unroll
  for (int r = 0; r < NLD; ++r) off[r] = deposit((tid + kMomNT * r) * V);

  uint4 buf[NLD];
  auto load = [&](uint64_t t) {
    const FP* base = st + 2 * tile_base(t);
#pragma unroll
    for (int r = 0; r < NLD; ++r)
      if (tid + kMomNT * r < nvec) buf[r] = *reinterpret_cast<const uint4*>(base + 2 * off[r]);
  };

  double acc[2][4] = {};
  uint64_t t = blockIdx.x;
  if (t < g.ntiles) load(t);
  for (; t < g.ntiles; t += gridDim.x) {
#pragma unroll
    for (int r = 0; r < NLD; ++r)
      if (tid + kMomNT * r < nvec) reinterpret_cast<uint4*>(s)[tid + kMomNT * r] = buf[r];
    __syncthreads();
    if (t + gridDim.x < g.ntiles) load(t + gridDim.x);
#pragma unroll
    for (int slot = 0; slot < 2; ++slot) {
      const uint32_t m = w + 8 * slot;
      if (m < g.nm) {  // warp-uniform
        const uint32_t k = g.mk[m], low = (1u << k) - 1, bit = 1u << k;
        FP s00 = 0, s11 = 0, re = 0, im = 0;
        for (uint32_t p = lane; p < tile_amps / 2; p += 32) {
          const uint32_t j0 = ((p & ~low) << 1) | (p & low);
          const V2 a0 = s[j0], a1 = s[j0 | bit];
          s00 = fma(a0.x, a0.x, fma(a0.y, a0.y, s00));
          s11 = fma(a1.x, a1.x, fma(a1.y, a1.y, s11));
          re = fma(a0.x, a1.x, fma(a0.y, a1.y, re));
          im = fma(a0.x, a1.y, fma(-a0.y, a1.x, im));
        }
        acc[slot][0] += s00;
        acc[slot][1] += s11;
        acc[slot][2] += re;
        acc[slot][3] += im;
      }
    }
    __syncthreads();
  }

#pragma unroll
  for (int slot = 0; slot < 2; ++slot) {
    const uint32_t m = w + 8 * slot;
    if (m < g.nm) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double v = warp_sum(acc[slot][c]);
        if (lane == 0) partials[(size_t{blockIdx.x} * kMomMaxT + m) * 4 + c] = v;
      }
    }
  }
}

// partials[blocks][kMomMaxT][4] -> out[4 * qubit + c], summed in block order (deterministic)
__global__ void __launch_bounds__(64)
k_moments_finish(const double* __restrict__ partials, uint32_t blocks, const __grid_constant__ MomGeom g,
                 double* __restrict__ out) {
  const uint32_t m = threadIdx.x >> 2, c = threadIdx.x & 3;
  if (m >= g.nm) return;
  double v = 0;
  for (uint32_t b = 0; b < blocks; ++b) v += partials[(size_t{b} * kMomMaxT + m) * 4 + c];
  out[4 * g.qubit[m] + c] = v;
}

template <typename FP, int TMAX>
int one_qubit_moments(qb200_ctx* ctx, const FP* st, unsigned n, double* out) {
  if (n == 0) return QB200_OK;
  if (reinterpret_cast<uintptr_t>(st) & 15) return QB200_ERR_UNSUPPORTED;  // 128-bit loads
  const unsigned T = std::min<unsigned>(TMAX, n), L = std::min<unsigned>(4, T);
  auto kern = k_moments<FP, TMAX>;
  static const int occ = [&] {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kMomNT, 0) != cudaSuccess || nb < 1) {
      (void) cudaGetLastError();
      nb = 1;
    }
    return nb;
  }();
  const uint64_t ntiles = uint64_t{1} << (n - T);
  const uint32_t blocks = (uint32_t) std::min<uint64_t>(ntiles, uint64_t{kNumSMs} * occ);
  const size_t pdoubles = size_t{blocks} * kMomMaxT * 4;
  int rc = ensure_scratch(ctx, (pdoubles + 4 * size_t{n}) * sizeof(double));
  if (rc) return rc;
  rc = ensure_pinned(ctx, 4 * size_t{n} * sizeof(double));
  if (rc) return rc;
  double* partials = (double*) ctx->scratch;
  double* dres = partials + pdoubles;

  unsigned next = 0;  // first qubit not measured yet
  while (next < n) {
    MomGeom g;
    g.T = T;
    g.ntiles = ntiles;
    bool in[kMaxQubits + 1] = {};
    for (unsigned b = 0; b < L; ++b) in[b] = true;
    unsigned have = L, fresh[kMomMaxT], nfresh = 0;
    if (next == 0) {
      for (unsigned b = 0; b < T; ++b) { in[b] = true; fresh[nfresh++] = b; }
      have = T;
      next = T;
    } else {
      while (have < T && next < n) { in[next] = true; fresh[nfresh++] = next++; ++have; }
      for (unsigned b = L; have < T; ++b)  // fill the tile with already measured low bits
        if (!in[b]) { in[b] = true; ++have; }
    }
    unsigned k = 0;
    for (unsigned b = 0; b < n; ++b)
      if (in[b]) g.pos[k++] = b;
    g.nm = nfresh;
    for (unsigned i = 0; i < nfresh; ++i) {
      g.qubit[i] = fresh[i];
      for (unsigned kk = 0; kk < T; ++kk)
        if (g.pos[kk] == fresh[i]) g.mk[i] = kk;
    }
    kern<<<blocks, kMomNT, 0, ctx->stream>>>(st, g, partials);
    QB_LAUNCHED(ctx);
    k_moments_finish<<<1, 64, 0, ctx->stream>>>(partials, blocks, g, dres);
    QB_LAUNCHED(ctx);
  }
  QB_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, dres, 4 * size_t{n} * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::copy((const double*) ctx->pinned, (const double*) ctx->pinned + 4 * size_t{n}, out);
  return QB200_OK;
}

}  // namespace qb200

using namespace qb200;

extern "C" int qb200_one_qubit_moments(qb200_ctx* ctx, int dtype, const void* state, unsigned num_qubits,
                                       double* out) {
  if (!ctx || !state || !out || num_qubits > kMaxQubits) return QB200_ERR_INVALID;
  DeviceGuard guard(ctx);
  if (dtype == QB200_F32) return one_qubit_moments<float, 12>(ctx, (const float*) state, num_qubits, out);
  if (dtype == QB200_F64) return one_qubit_moments<double, 11>(ctx, (const double*) state, num_qubits, out);
  return QB200_ERR_INVALID;
}
