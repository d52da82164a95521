// expect_monomial.cu -- expectation values of XOR-monomial operators without a mat-vec.
//
// A tensor product of Paulis / phase gates -- what lib/expect.h:106-151 fuses an operator string into, and what
// qsimcirq's simulate_expectation_values feeds it -- is a matrix with ONE non-zero per row, in column r ^ xm for
// a fixed mask xm.  Then  <psi|M|psi> = sum_i conj(a_i) v[r(i)] a_(i ^ XM)  (r(i) = the target bits of i, XM = xm
// deposited at the target positions): no 2^G x 2^G mat-vec, no tensor cores, every amplitude read once -- a pure
// HBM-bound read pass (8 * 2^n bytes in fp32) for any G <= 6, where the dense kernels need 2.1-3.5 ms (fp32,
// G = 4..6, n = 30) or 5-65 ms (fp64) against 1.3 / 2.6 ms of HBM time.  Detected on the host from exact zeros
// (products of matrices with exact zeros keep them exact); anything else takes the dense kernels.
// Arithmetic contract as everywhere: products in FP, accumulation in double (lib/simulator_basic.h:323-324).
#include "gate_kernels.cuh"

namespace qb200 {

int finish_expectation(qb200_ctx* ctx, double* partials, uint32_t blocks, double out[2]);

template <typename FP>
struct MonoParam {
  FP v[2 << kMaxTargets];     // (re, im) of the non-zero of row r
  uint64_t XM = 0;            // xm deposited at the target positions
  uint32_t xm = 0;
  uint32_t nq = 0;
  uint32_t tpos[kMaxTargets] = {};
  uint32_t pivot = 0;         // highest set bit of XM (above bit 0): a thread starts from the pairs with that bit 0
};

constexpr int kMonoNT = 256;
constexpr int kMonoUnroll = 2;

// A thread takes the aligned amplitude pair A = (i, i + 1), i even, with one 128-bit access (two for fp64) and --
// MODE 2 -- the pair B = A ^ XM' (XM' = XM without bit 0) of its partners; the partner of A[e] is B[e ^ f] with
// f = bit 0 of XM.  MODE 0: diagonal operator (no partner).  MODE 1: XM = 1, the partners sit inside A.
// MODE 2 walks only the pairs whose `pivot` bit (highest bit of XM') is 0 and adds both directions, so every
// amplitude is read exactly once in every mode.  NQ compile-time: the row extraction is unrolled.
template <typename FP, int NQ, int MODE>
__global__ void __launch_bounds__(kMonoNT)
k_expect_mono(const FP* __restrict__ st, const uint64_t items, const __grid_constant__ MonoParam<FP> mp,
              double* __restrict__ partials) {
  __shared__ FP sv[2 << NQ];
  for (uint32_t k = threadIdx.x; k < (2u << NQ); k += kMonoNT) sv[k] = mp.v[k];
  __syncthreads();
  double ere = 0, eim = 0;
  const uint64_t stride = uint64_t{gridDim.x} * kMonoNT;
  const uint64_t low = (uint64_t{1} << (mp.pivot - 1)) - 1;  // MODE 2: pivot >= 1, in units of pairs
  const uint64_t XMp = mp.XM & ~uint64_t{1};
  const uint32_t f = (uint32_t) (mp.XM & 1u);
  const uint32_t t0 = mp.tpos[0] == 0 ? 1u : 0u;  // bit 0 is a target: rows of A[0] and A[1] differ in bit 0
  auto row_of = [&](uint64_t i) {
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < NQ; ++k) r |= (uint32_t) ((i >> mp.tpos[k]) & 1u) << k;
    return r;
  };
  // conj(x) * (v[r] * y)
  auto term = [&](uint32_t r, FP xr, FP xi, FP yr, FP yi, FP& re, FP& im) {
    const FP vr = sv[2 * r], vi = sv[2 * r + 1];
    const FP tr = vr * yr - vi * yi, ti = vr * yi + vi * yr;
    re += xr * tr + xi * ti;
    im += xr * ti - xi * tr;
  };
  for (uint64_t w0 = blockIdx.x * uint64_t{kMonoNT} + threadIdx.x; w0 < items; w0 += kMonoUnroll * stride) {
    uint64_t idx[kMonoUnroll];
    FP a[kMonoUnroll][4], b[kMonoUnroll][4];
    bool ok[kMonoUnroll];
#pragma unroll
    for (int u = 0; u < kMonoUnroll; ++u) {
      const uint64_t w = w0 + u * stride;
      ok[u] = w < items;
      const uint64_t ww = ok[u] ? w : w0;
      idx[u] = 2 * (MODE == 2 ? (((ww & ~low) << 1) | (ww & low)) : ww);
      ld2(st + 2 * idx[u], a[u][0], a[u][1], a[u][2], a[u][3]);
      if constexpr (MODE == 2) ld2(st + 2 * (idx[u] ^ XMp), b[u][0], b[u][1], b[u][2], b[u][3]);
    }
    FP re = 0, im = 0;
#pragma unroll
    for (int u = 0; u < kMonoUnroll; ++u) {
      if (!ok[u]) continue;
      const uint32_t r0 = row_of(idx[u]);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const uint32_t r = r0 | (e ? t0 : 0u);
        const FP xr = a[u][2 * e], xi = a[u][2 * e + 1];
        if constexpr (MODE == 0) {
          const FP n2 = xr * xr + xi * xi;
          re += sv[2 * r] * n2;
          im += sv[2 * r + 1] * n2;
        } else if constexpr (MODE == 1) {
          term(r, xr, xi, a[u][2 * (e ^ 1)], a[u][2 * (e ^ 1) + 1], re, im);
        } else {
          // f is warp-uniform: partner of A[e] is B[e ^ f]
          const FP yr = f ? b[u][2 * (e ^ 1)] : b[u][2 * e], yi = f ? b[u][2 * (e ^ 1) + 1] : b[u][2 * e + 1];
          term(r, xr, xi, yr, yi, re, im);
          term(r ^ mp.xm, yr, yi, xr, xi, re, im);
        }
      }
    }
    ere += re;  // products and the <= 8 terms of an iteration in FP, running sum in double
    eim += im;
  }
  block_sum2<kMonoNT>(ere, eim);
  if (threadIdx.x == 0) {
    partials[2 * blockIdx.x] = ere;
    partials[2 * blockIdx.x + 1] = eim;
  }
}

template <typename FP, int NQ>
void launch_mono(int mode, uint32_t blocks, cudaStream_t stream, const FP* st, uint64_t items, const MonoParam<FP>& mp,
                 double* partials) {
  if (mode == 0) k_expect_mono<FP, NQ, 0><<<blocks, kMonoNT, 0, stream>>>(st, items, mp, partials);
  else if (mode == 1) k_expect_mono<FP, NQ, 1><<<blocks, kMonoNT, 0, stream>>>(st, items, mp, partials);
  else k_expect_mono<FP, NQ, 2><<<blocks, kMonoNT, 0, stream>>>(st, items, mp, partials);
}

// QB200_ERR_UNSUPPORTED: not an XOR-monomial matrix (or arguments the dense path should judge) -> caller falls through.
template <typename FP>
int expect_monomial(qb200_ctx* ctx, const FP* st, unsigned n, const unsigned* qs, unsigned nq, const FP* m,
                    double* out) {
  if (!ctx || !st || !qs || !m || nq == 0 || nq > kMaxTargets || n > kMaxQubits || nq > n) return QB200_ERR_UNSUPPORTED;
  for (unsigned k = 0; k < nq; ++k)
    if (qs[k] >= n || (k > 0 && qs[k] <= qs[k - 1])) return QB200_ERR_UNSUPPORTED;
  if (nq < 3 || n < 2 || (reinterpret_cast<uintptr_t>(st) & 15)) return QB200_ERR_UNSUPPORTED;
  const unsigned N = 1u << nq;
  MonoParam<FP> mp;
  bool have = false;
  for (unsigned r = 0; r < N; ++r) {
    int col = -1;
    for (unsigned c = 0; c < N; ++c) {
      const FP re = m[2 * (size_t{r} * N + c)], im = m[2 * (size_t{r} * N + c) + 1];
      if (re != 0 || im != 0) {
        if (col >= 0) return QB200_ERR_UNSUPPORTED;  // two non-zeros in a row (NaN counts as non-zero)
        col = (int) c;
      }
    }
    mp.v[2 * r] = mp.v[2 * r + 1] = 0;
    if (col < 0) continue;  // zero row
    const unsigned x = r ^ (unsigned) col;
    if (have && x != mp.xm) return QB200_ERR_UNSUPPORTED;
    have = true;
    mp.xm = x;
    mp.v[2 * r] = m[2 * (size_t{r} * N + col)];
    mp.v[2 * r + 1] = m[2 * (size_t{r} * N + col) + 1];
  }
  mp.nq = nq;
  for (unsigned k = 0; k < nq; ++k) {
    mp.tpos[k] = qs[k];
    if ((mp.xm >> k) & 1u) {
      mp.XM |= uint64_t{1} << qs[k];
      if (qs[k] > 0) mp.pivot = qs[k];  // highest set bit of XM above bit 0
    }
  }
  DeviceGuard guard(ctx);
  const int mode = mp.XM == 0 ? 0 : (mp.XM == 1 ? 1 : 2);
  const uint64_t items = mode == 2 ? uint64_t{1} << (n - 2) : uint64_t{1} << (n - 1);  // amplitude pairs a thread starts from
  uint64_t need = (items + uint64_t{kMonoNT} * kMonoUnroll - 1) / (uint64_t{kMonoNT} * kMonoUnroll);
  const uint32_t blocks = (uint32_t) (need < kNumSMs * 8 ? need : kNumSMs * 8);
  int rc = ensure_scratch(ctx, (2 * size_t{blocks} + 2) * sizeof(double));
  if (rc) return rc;
  double* partials = (double*) ctx->scratch;
  switch (nq) {
    case 3: launch_mono<FP, 3>(mode, blocks, ctx->stream, st, items, mp, partials); break;
    case 4: launch_mono<FP, 4>(mode, blocks, ctx->stream, st, items, mp, partials); break;
    case 5: launch_mono<FP, 5>(mode, blocks, ctx->stream, st, items, mp, partials); break;
    default: launch_mono<FP, 6>(mode, blocks, ctx->stream, st, items, mp, partials); break;
  }
  QB_LAUNCHED(ctx);
  return finish_expectation(ctx, partials, blocks, out);
}

int expect_monomial_f32(qb200_ctx* ctx, const float* st, unsigned n, const unsigned* qs, unsigned nq, const float* m,
                        double* out) {
  return expect_monomial<float>(ctx, st, n, qs, nq, m, out);
}
int expect_monomial_f64(qb200_ctx* ctx, const double* st, unsigned n, const unsigned* qs, unsigned nq, const double* m,
                        double* out) {
  return expect_monomial<double>(ctx, st, n, qs, nq, m, out);
}

}  // namespace qb200
