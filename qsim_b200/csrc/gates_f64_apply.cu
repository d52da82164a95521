// gates_f64_apply.cu -- instantiates the double gate-application kernels.
#include "gate_launch.cuh"

namespace qb200 {
int gate_apply_f64(qb200_ctx* ctx, double* st, unsigned n, const unsigned* qs, unsigned nq,
                      const unsigned* cqs, unsigned nc, uint64_t cvals, const double* m, double* out) {
  return gate_pass<double, false>(ctx, st, n, qs, nq, cqs, nc, cvals, m, out);
}
}  // namespace qb200
