// simulator_b200.h -- drop-in for qsim's SimulatorCUDA (lib/simulator_cuda.h:35-919):
// same member set, so it compiles behind lib/run_qsim.h, lib/qtrajectory.h,
// lib/expect.h, lib/hybrid.h, the pybind layer and every tests/*_testfixture.h.
#ifndef QSIM_B200_SIMULATOR_B200_H_
#define QSIM_B200_SIMULATOR_B200_H_

#include <complex>
#include <cstdint>
#include <vector>

#include "statespace_b200.h"

namespace qsim {

template <typename FP = float>
class SimulatorB200 final {
 public:
  using StateSpace = StateSpaceB200<FP>;
  using State = typename StateSpace::State;
  using fp_type = typename StateSpace::fp_type;

  SimulatorB200() : ctx_(b200::MakeContext()) {}

  // lib/simulator_cuda.h:70-125; qs sorted ascending, up to 6 qubits (more: ignored).
  void ApplyGate(const std::vector<unsigned>& qs, const fp_type* matrix, State& state) const {
    QB200_CHECK(ctx(), qb200_apply_gate(ctx(), b200::DType<FP>::value, state.get(), state.num_qubits(),
                                        qs.data(), (unsigned) qs.size(), matrix));
  }

  // lib/simulator_cuda.h:135-207
  void ApplyControlledGate(const std::vector<unsigned>& qs, const std::vector<unsigned>& cqs,
                           uint64_t cvals, const fp_type* matrix, State& state) const {
    QB200_CHECK(ctx(), qb200_apply_controlled_gate(ctx(), b200::DType<FP>::value, state.get(),
                                                   state.num_qubits(), qs.data(), (unsigned) qs.size(),
                                                   cqs.data(), (unsigned) cqs.size(), cvals, matrix));
  }

  // lib/simulator_cuda.h:216-260
  std::complex<double> ExpectationValue(const std::vector<unsigned>& qs, const fp_type* matrix,
                                        const State& state) const {
    double out[2] = {0, 0};
    QB200_CHECK(ctx(), qb200_expectation_value(ctx(), b200::DType<FP>::value, state.get(), state.num_qubits(),
                                               qs.data(), (unsigned) qs.size(), matrix, out));
    return {out[0], out[1]};
  }

  // Batched expectation values (no reference counterpart; used by expect_b200.h): between Begin and End
  // ExpectationValue enqueues its pass and returns NaN at once; End synchronises the stream once and
  // returns the values in call order.
  void BeginExpectationBatch(unsigned expected = 0) const {
    QB200_CHECK(ctx(), qb200_reduce_batch_begin(ctx(), expected));
  }
  std::vector<std::complex<double>> EndExpectationBatch(unsigned max_count) const {
    std::vector<double> buf(2 * std::size_t{max_count} + 2);
    uint32_t count = 0;
    QB200_CHECK(ctx(), qb200_reduce_batch_end(ctx(), buf.data(), max_count, &count));
    std::vector<std::complex<double>> out(count);
    for (uint32_t i = 0; i < count; ++i) out[i] = {buf[2 * i], buf[2 * i + 1]};
    return out;
  }

  // Reduced density matrices of all qubits from 3-4 read-only passes (qb200_one_qubit_moments): 4 doubles per
  // qubit, S00, S11, Re S01, Im S01.  Empty when the state cannot take the path (unaligned wrapped memory).
  std::vector<double> OneQubitMoments(const State& state) const {
    std::vector<double> out(4 * std::size_t{state.num_qubits()} + 4);
    int rc = qb200_one_qubit_moments(ctx(), b200::DType<FP>::value, state.get(), state.num_qubits(), out.data());
    if (rc == QB200_ERR_UNSUPPORTED) return {};
    QB200_CHECK(ctx(), rc);
    out.resize(4 * std::size_t{state.num_qubits()});
    return out;
  }

  // lib/simulator_cuda.h:265-267 (the reference's tests size their sweeps from it)
  static unsigned SIMDRegisterSize() { return 32; }

  void SetStream(void* cuda_stream) const { QB200_CHECK(ctx(), qb200_ctx_set_stream(ctx(), cuda_stream)); }

 private:
  qb200_ctx* ctx() const { return ctx_.get(); }
  std::shared_ptr<qb200_ctx> ctx_;
};

}  // namespace qsim

#endif  // QSIM_B200_SIMULATOR_B200_H_
