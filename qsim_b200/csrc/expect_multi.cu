// expect_multi.cu -- expectation values of SEVERAL operators on the same one or two qubits in ONE read pass.
//
// What for: sampling a Kraus operator of a non-unitary channel (lib/qtrajectory.h:344-352) asks for
// p_i = <psi| K_i^dagger K_i |psi> operator by operator, one full read of the state and one host
// synchronisation each.  All K_i of a channel act on the same qubits of the same state, so one pass that loads
// every group of 2^G amplitudes once and applies all matrices to it yields every p_i ("batched
// Kraus-probability ExpectationValues").  Arithmetic per operator as in SimulatorCUDA::ExpectationValue
// (lib/simulator_cuda.h:216-260): products in the state's precision, accumulation in double.
#include "common.cuh"
#include "gate_geom.cuh"

namespace qb200 {

int ensure_results(qb200_ctx* ctx, uint32_t slots);
int finish_expectations(qb200_ctx* ctx, double* partials, uint32_t blocks, uint32_t count, double* out);

namespace {

constexpr int kMaxOps = 8;
constexpr int kMultiNT = 256;

template <typename FP, int G>
struct MultiMat {
  FP m[kMaxOps][2 << (2 * G)];   // row-major, interleaved (re, im)
  uint32_t count;
};

template <typename FP> struct Amp { FP re, im; };
template <typename FP> __device__ __forceinline__ Amp<FP> load_amp(const FP* p);
template <> __device__ __forceinline__ Amp<float> load_amp<float>(const float* p) {
  const float2 v = __ldg(reinterpret_cast<const float2*>(p));
  return {v.x, v.y};
}
template <> __device__ __forceinline__ Amp<double> load_amp<double>(const double* p) {
  const double2 v = __ldg(reinterpret_cast<const double2*>(p));
  return {v.x, v.y};
}

// one group per thread and iteration, two iterations in flight; partials[(op * gridDim.x + block) * 2 + {0,1}]
template <typename FP, int G>
__global__ void __launch_bounds__(kMultiNT)
k_expect_multi(const FP* __restrict__ st, const __grid_constant__ Geom g, const __grid_constant__ MultiMat<FP, G> mats,
               double* __restrict__ partials) {
  constexpr int N = 1 << G;
  double acc[kMaxOps][2];
#pragma unroll
  for (int o = 0; o < kMaxOps; ++o) acc[o][0] = acc[o][1] = 0;
  const uint64_t stride = uint64_t{gridDim.x} * kMultiNT;
  auto load = [&](uint64_t i, Amp<FP> (&x)[N]) {
    const FP* p = st + 2 * expand_index(i, g);
#pragma unroll
    for (int k = 0; k < N; ++k) x[k] = load_amp<FP>(p + 2 * elem_offset<G>(k, g));
  };
  uint64_t i = blockIdx.x * uint64_t{kMultiNT} + threadIdx.x;
  Amp<FP> xn[N];
  if (i < g.work) load(i, xn);
  for (; i < g.work; i += stride) {
    Amp<FP> x[N];
#pragma unroll
    for (int k = 0; k < N; ++k) x[k] = xn[k];
    if (i + stride < g.work) load(i + stride, xn);
#pragma unroll
    for (int o = 0; o < kMaxOps; ++o) {
      if (o < (int) mats.count) {
#pragma unroll
        for (int r = 0; r < N; ++r) {
          FP re = 0, im = 0;
#pragma unroll
          for (int c = 0; c < N; ++c) {
            const FP mr = mats.m[o][2 * (r * N + c)], mi = mats.m[o][2 * (r * N + c) + 1];
            re += mr * x[c].re - mi * x[c].im;
            im += mr * x[c].im + mi * x[c].re;
          }
          acc[o][0] += x[r].re * re + x[r].im * im;
          acc[o][1] += x[r].re * im - x[r].im * re;
        }
      }
    }
  }
#pragma unroll
  for (int o = 0; o < kMaxOps; ++o) {
    if (o < (int) mats.count) {
      block_sum2<kMultiNT>(acc[o][0], acc[o][1]);
      if (threadIdx.x == 0) {
        partials[(size_t{(unsigned) o} * gridDim.x + blockIdx.x) * 2] = acc[o][0];
        partials[(size_t{(unsigned) o} * gridDim.x + blockIdx.x) * 2 + 1] = acc[o][1];
      }
    }
  }
}

template <typename FP, int G>
int run_multi(qb200_ctx* ctx, const FP* st, unsigned n, const unsigned* qs, const FP* matrices, unsigned count,
              double* out) {
  Geom g;
  int rc = make_geom(n, qs, G, nullptr, 0, 0, false, &g);
  if (rc) return rc;
  MultiMat<FP, G> mats;
  mats.count = count;
  constexpr size_t per = size_t{2} << (2 * G);
  for (unsigned o = 0; o < count; ++o)
    for (size_t k = 0; k < per; ++k) mats.m[o][k] = matrices[o * per + k];
  const uint64_t need = (g.work + kMultiNT - 1) / kMultiNT;
  const uint32_t blocks = (uint32_t) std::min<uint64_t>(need, uint64_t{kNumSMs} * 8);
  rc = ensure_scratch(ctx, size_t{count} * blocks * 2 * sizeof(double));
  if (rc) return rc;
  double* partials = (double*) ctx->scratch;
  k_expect_multi<FP, G><<<blocks, kMultiNT, 0, ctx->stream>>>(st, g, mats, partials);
  QB_LAUNCHED(ctx);
  return finish_expectations(ctx, partials, blocks, count, out);
}

template <typename FP>
int expectation_values_multi(qb200_ctx* ctx, const FP* st, unsigned n, const unsigned* qs, unsigned nq,
                             const FP* matrices, unsigned count, double* out) {
  if (!ctx || !st || !qs || !matrices || !out || count == 0) return QB200_ERR_INVALID;
  if (nq < 1 || nq > 2 || count > (unsigned) kMaxOps) return QB200_ERR_UNSUPPORTED;
  DeviceGuard guard(ctx);
  if (int drc = check_state_device(ctx, st)) return drc;
  return nq == 1 ? run_multi<FP, 1>(ctx, st, n, qs, matrices, count, out)
                 : run_multi<FP, 2>(ctx, st, n, qs, matrices, count, out);
}

}  // namespace
}  // namespace qb200

using namespace qb200;

extern "C" int qb200_expectation_values_multi(qb200_ctx* ctx, int dtype, const void* state, unsigned num_qubits,
                                              const unsigned* qs, unsigned num_targets, const void* matrices,
                                              unsigned count, double* out_re_im) {
  if (out_re_im)
    for (unsigned i = 0; i < 2 * count; ++i) out_re_im[i] = 0;
  if (dtype == QB200_F32)
    return expectation_values_multi<float>(ctx, (const float*) state, num_qubits, qs, num_targets,
                                           (const float*) matrices, count, out_re_im);
  if (dtype == QB200_F64)
    return expectation_values_multi<double>(ctx, (const double*) state, num_qubits, qs, num_targets,
                                            (const double*) matrices, count, out_re_im);
  return QB200_ERR_INVALID;
}
