// gates_f64_expect.cu -- instantiates the double expectation-value kernels.
#include "gate_launch.cuh"

namespace qb200 {
int gate_expect_f64(qb200_ctx* ctx, double* st, unsigned n, const unsigned* qs, unsigned nq,
                      const unsigned* cqs, unsigned nc, uint64_t cvals, const double* m, double* out) {
  return gate_pass<double, true>(ctx, st, n, qs, nq, cqs, nc, cvals, m, out);
}
}  // namespace qb200
