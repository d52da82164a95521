// moments.cu -- qb200_one_qubit_moments: the reduced density matrix of EVERY qubit from a handful of
// read-only passes (the observable loop of lib/expect.h:106-151 / apps/qsim_qtrajectory_cuda.cu
// costs one full pass per single-qubit operator; 52 of the 56 passes per trajectory of BASELINE config 5).
//
// For qubit q:  S00 = sum |a_i|^2 over bit_q(i) = 0,  S11 = the same over bit_q(i) = 1,
//               S01 = sum conj(a_i0) * a_i1 over the pairs (i0, i1 = i0 | 1 << q),
// and any single-qubit operator M has  <psi|M|psi> = m00 S00 + m11 S11 + m01 S01 + m10 conj(S01).
//
// One pass = one set B of T tile bits (T = 12 fp32 / 11 fp64: a 32 KB tile): the low 4 index bits (128-byte
// runs -> fully coalesced 128-bit loads) plus a window of T - 4 consecutive higher bit positions.  A block stages the 2^T
// amplitudes whose other index bits equal the tile number in shared memory (next tile prefetched into
// registers meanwhile); the measured tile bits are taken two at a time: a warp reads the four amplitudes of a
// (k1, k2) quad once and feeds both qubits' pairs from registers (fp32: packed FFMA2 on the (re, im) pairs as
// they lie in shared memory, 4 instructions per amplitude pair); a low bit is grouped with a high one so that
// the lanes of a warp spread over the banks.  Products in FP, per-tile sums in FP, running sums in double (the
// reference's precision contract, lib/simulator_basic.h:323-324), fixed reduction tree -> deterministic.
// Pass 0 measures bits 0..T-1, every later pass T - 4 new qubits: 3 passes at 26 qubits, 4 at 30 -- against
// one pass per operator.  Measured (profiles/r01_ncu_moments_summary.txt): 0.63 ms at 26 qubits, 10.6 ms at 30
// = 8 read-pass equivalents; pass 0 is bound by issue slots + the tile barrier (6 groups on 8 warps), not HBM.
#include <algorithm>

#include "gate_kernels.cuh"

namespace qb200 {

constexpr int kMomNT = 256;
constexpr int kMomMaxT = 12;

struct MomGeom {
  uint32_t T = 0;                // tile bits: index bits [0, L) and the window [hs, hs + T - L)
  uint32_t L = 0, hs = 0;
  uint32_t pos[kMomMaxT] = {};   // their positions in the amplitude index, ascending
  uint32_t nm = 0;               // measured tile bits in this pass
  uint32_t mk[kMomMaxT] = {};    // tile-bit number of each
  uint32_t qubit[kMomMaxT] = {};  // = pos[mk[.]]
  uint64_t ntiles = 0;
  // measured tile bits two at a time: a warp reads the four amplitudes of a (k1, k2) quad once and feeds both
  // qubits' pairs from registers -- half the shared-memory traffic and index arithmetic of a walk per qubit
  uint32_t ngroups = 0;          // 0: tile too small, walk the pairs of one measured bit per warp
  uint32_t sf = 1;               // warps per group (each takes 1/sf of the quads); ngroups * sf <= 8
  uint32_t gk[6][2] = {};        // tile-bit numbers of the group, ascending
  uint32_t gm[6][2] = {};        // index into mk[] / partials, 0xffffffff = filler bit
};

template <typename FP, int TMAX, int MINB>
__global__ void __launch_bounds__(kMomNT, MINB)
k_moments(const FP* __restrict__ st, const __grid_constant__ MomGeom g, double* __restrict__ partials) {
  using V2 = typename Vec2<FP>::type;
  constexpr int V = 16 / (int) sizeof(V2);                // amplitudes per 128-bit access
  constexpr int NLD = (1 << TMAX) / V / kMomNT;            // 128-bit loads per thread per tile
  __shared__ __align__(16) V2 s[1 << TMAX];
  const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t tile_amps = 1u << g.T;
  const uint32_t nvec = tile_amps / V;

  // tile bit k <-> index bit k (k < L) or hs + k - L: two shifts instead of a loop over bit positions
  const uint32_t lmask = (1u << g.L) - 1, d = g.hs - g.L;
  auto deposit = [&](uint32_t j) { return uint64_t(j & lmask) | (uint64_t(j >> g.L) << g.hs); };
  auto tile_base = [&](uint64_t t) {
    return ((t >> d) << (g.hs + g.T - g.L)) | ((t & ((uint64_t{1} << d) - 1)) << g.L);
  };

  uint4 buf[NLD];
  auto load = [&](uint64_t t) {
    const FP* base = st + 2 * tile_base(t);
#pragma unroll
    for (int r = 0; r < NLD; ++r)
      if (tid + kMomNT * r < nvec)
        buf[r] = *reinterpret_cast<const uint4*>(base + 2 * deposit((tid + kMomNT * r) * V));
  };

  double acc[2][4] = {};
  const uint32_t grp = w / g.sf, part = w - grp * g.sf;
  const uint32_t gi = grp < g.ngroups ? grp : 0;
  const uint32_t b1 = 1u << g.gk[gi][0], b2 = 1u << g.gk[gi][1], lo1 = b1 - 1, lo2 = b2 - 1;
  // quad number -> tile index with zeros at the two measured bits; stepping by 32 quads = adding expand2(32) with
  // the held bits set so that carries run through them
  auto expand2 = [&](uint32_t o) {
    uint32_t j = ((o & ~lo1) << 1) | (o & lo1);
    return ((j & ~lo2) << 1) | (j & lo2);
  };
  const uint32_t held = b1 | b2, step = expand2(32);
  uint64_t t = blockIdx.x;
  if (t < g.ntiles) load(t);
  for (; t < g.ntiles; t += gridDim.x) {
#pragma unroll
    for (int r = 0; r < NLD; ++r)
      if (tid + kMomNT * r < nvec) reinterpret_cast<uint4*>(s)[tid + kMomNT * r] = buf[r];
    __syncthreads();
    if (t + gridDim.x < g.ntiles) load(t + gridDim.x);
    if (g.ngroups == 0) {
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
    } else if (grp < g.ngroups) {  // warp-uniform
      const uint32_t per = (tile_amps >> 2) / g.sf;
      if constexpr (sizeof(FP) == 4) {
        // packed FFMA2: (x, y) pairs as they lie in shared memory; 4 instructions per amplitude pair
        const uint64_t* s64 = reinterpret_cast<const uint64_t*>(s);
        uint64_t pa[2][4] = {};
        auto pair_acc = [&](int x, uint64_t a0, uint64_t a1) {
          float a1x, a1y;
          unpack2(a1, a1x, a1y);
          pa[x][0] = fma2(a0, a0, pa[x][0]);                 // (sum x0^2, sum y0^2)
          pa[x][1] = fma2(a1, a1, pa[x][1]);
          pa[x][2] = fma2(a0, a1, pa[x][2]);                 // (sum x0 x1, sum y0 y1)
          pa[x][3] = fma2(a0, pack2(a1y, a1x), pa[x][3]);    // (sum x0 y1, sum y0 x1)
        };
        uint32_t j = expand2(part * per + lane);
        for (uint32_t o = part * per + lane; o < (part + 1) * per; o += 32, j = ((j | held) + step) & ~held) {
          const uint64_t a00 = s64[j], a01 = s64[j | b1], a10 = s64[j | b2], a11 = s64[j | b1 | b2];
          pair_acc(0, a00, a01);
          pair_acc(0, a10, a11);
          pair_acc(1, a00, a10);
          pair_acc(1, a01, a11);
        }
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          float l, h;
          unpack2(pa[x][0], l, h); acc[x][0] += l + h;
          unpack2(pa[x][1], l, h); acc[x][1] += l + h;
          unpack2(pa[x][2], l, h); acc[x][2] += l + h;
          unpack2(pa[x][3], l, h); acc[x][3] += l - h;
        }
      } else {
        FP pa[2][4] = {};
        auto pair_acc = [&](int x, V2 a0, V2 a1) {
          pa[x][0] = fma(a0.x, a0.x, fma(a0.y, a0.y, pa[x][0]));
          pa[x][1] = fma(a1.x, a1.x, fma(a1.y, a1.y, pa[x][1]));
          pa[x][2] = fma(a0.x, a1.x, fma(a0.y, a1.y, pa[x][2]));
          pa[x][3] = fma(a0.x, a1.y, fma(-a0.y, a1.x, pa[x][3]));
        };
        uint32_t j = expand2(part * per + lane);
        for (uint32_t o = part * per + lane; o < (part + 1) * per; o += 32, j = ((j | held) + step) & ~held) {
          const V2 a00 = s[j], a01 = s[j | b1], a10 = s[j | b2], a11 = s[j | b1 | b2];
          pair_acc(0, a00, a01);
          pair_acc(0, a10, a11);
          pair_acc(1, a00, a10);
          pair_acc(1, a01, a11);
        }
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[x][c] += pa[x][c];
      }
    }
    __syncthreads();
  }

  if (g.ngroups == 0) {
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
    return;
  }
  // the sf warps of a group -> one partial per (block, measured bit), fixed order
  __shared__ double red[kMomNT / 32][8];
  if (grp < g.ngroups) {
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double v = warp_sum(acc[x][c]);
        if (lane == 0) red[w][4 * x + c] = v;
      }
  }
  __syncthreads();
  if (tid < g.ngroups * 8) {
    const uint32_t gg = tid >> 3, xc = tid & 7, m = g.gm[gg][xc >> 2];
    if (m != 0xffffffffu) {
      double v = 0;
      for (uint32_t p = 0; p < g.sf; ++p) v += red[gg * g.sf + p][xc];
      partials[(size_t{blockIdx.x} * kMomMaxT + m) * 4 + (xc & 3)] = v;
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

template <typename FP, int TMAX, int MINB>
int one_qubit_moments(qb200_ctx* ctx, const FP* st, unsigned n, double* out) {
  if (n == 0) return QB200_OK;
  if (reinterpret_cast<uintptr_t>(st) & 15) return QB200_ERR_UNSUPPORTED;  // 128-bit loads
  const unsigned T = std::min<unsigned>(TMAX, n), L = std::min<unsigned>(4, T);
  auto kern = k_moments<FP, TMAX, MINB>;
  static PerDevice occ_cache;
  const int occ = occ_cache.get(ctx, [&] {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kMomNT, 0) != cudaSuccess || nb < 1) {
      (void) cudaGetLastError();
      nb = 1;
    }
    return nb;
  });
  const uint64_t ntiles = uint64_t{1} << (n - T);
  const uint32_t blocks = (uint32_t) std::min<uint64_t>(ntiles, uint64_t(grid_sms(ctx)) * grid_occ(ctx, occ));
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
    const unsigned W = T - L;
    unsigned fresh[kMomMaxT], nfresh = 0;
    g.L = L;
    if (next == 0) {
      g.hs = L;
      for (unsigned b = 0; b < T; ++b) fresh[nfresh++] = b;
      next = T;
    } else {
      g.hs = std::min(next, n - W);  // the window slides down over measured bits when fewer than W qubits are left
      while (nfresh < W && next < n) fresh[nfresh++] = next++;
    }
    for (unsigned k = 0; k < T; ++k) g.pos[k] = k < L ? k : g.hs + (k - L);
    g.nm = nfresh;
    for (unsigned i = 0; i < nfresh; ++i) {
      g.qubit[i] = fresh[i];
      for (unsigned kk = 0; kk < T; ++kk)
        if (g.pos[kk] == fresh[i]) g.mk[i] = kk;
    }
    if (T >= 2) {
      const unsigned h = (nfresh + 1) / 2;  // bit i with bit i + h: the low (bank-conflicting) bits get a high partner
      for (unsigned i = 0; i < h; ++i) {
        unsigned k1 = g.mk[i], m1 = i, k2, m2;
        if (i + h < nfresh) { k2 = g.mk[i + h]; m2 = i + h; }
        else { k2 = k1 == T - 1 ? 0 : T - 1; m2 = 0xffffffffu; }  // filler partner, result dropped
        if (k1 > k2) { std::swap(k1, k2); std::swap(m1, m2); }
        g.gk[i][0] = k1; g.gk[i][1] = k2;
        g.gm[i][0] = m1; g.gm[i][1] = m2;
      }
      g.ngroups = h;
      g.sf = 1;
      while (g.ngroups * g.sf * 2 <= 8 && g.sf * 2 <= (1u << (T - 2))) g.sf *= 2;
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
  // two resident blocks per SM (128 registers): three (80 registers, 128-184 B of spills) measured 7 % slower
  if (dtype == QB200_F32) return one_qubit_moments<float, 12, 2>(ctx, (const float*) state, num_qubits, out);
  if (dtype == QB200_F64) return one_qubit_moments<double, 11, 2>(ctx, (const double*) state, num_qubits, out);
  return QB200_ERR_INVALID;
}
