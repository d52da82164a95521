// gates_f32_apply.cu -- instantiates the float gate-application kernels.
#include "gate_launch.cuh"

namespace qb200 {
int gate_apply_f32(qb200_ctx* ctx, float* st, unsigned n, const unsigned* qs, unsigned nq,
                      const unsigned* cqs, unsigned nc, uint64_t cvals, const float* m, double* out) {
  return gate_pass<float, false>(ctx, st, n, qs, nq, cqs, nc, cvals, m, out);
}
}  // namespace qb200
