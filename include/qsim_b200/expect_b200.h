// expect_b200.h -- batched sibling of lib/expect.h:106-151 (ExpectationValue<IO, Fuser> over
// operator strings) for SimulatorB200: the same host logic per string (weight-only strings,
// single operators as they are, longer strings fused with max_fused_size = 6 into one gate), but
// the read-only passes of ALL observables are enqueued back to back and read after ONE stream
// synchronisation instead of one per operator string, and -- when there are at
// least kMomentsThreshold of them -- all single-qubit operators are evaluated from the reduced density
// matrices of the qubits (SimulatorB200::OneQubitMoments: 3-4 passes for every qubit together instead
// of one pass per operator).  With `use_moments` = false the values are identical to calling
// qsim::ExpectationValue<IO, Fuser> per observable (the same kernels in the same order); with the
// moments they agree to the precision contract (products in FP, sums in double; |delta| ~ 1e-7 in fp32).
#ifndef QSIM_B200_EXPECT_B200_H_
#define QSIM_B200_EXPECT_B200_H_

#include <complex>
#include <cstddef>
#include <vector>

#include "expect.h"  // the reference's OpString / fused-gate types, consumed in place

namespace qsim {

constexpr std::size_t kMomentsThreshold = 6;

/**
 * Expectation values of several observables, each a sum of weighted operator strings
 * (argument meaning as in lib/expect.h:95-104).  An observable whose strings cannot be
 * fused into one gate of at most six qubits reports 0, like the reference.
 */
template <typename IO, typename Fuser, typename FP, typename Simulator>
std::vector<std::complex<double>> ExpectationValues(
    const std::vector<std::vector<OpString<FP>>>& observables,
    const Simulator& simulator, const typename Simulator::State& state, bool use_moments = true) {
  struct Term {
    std::size_t observable;
    std::complex<double> weight;
  };
  std::vector<std::complex<double>> evals(observables.size(), 0);
  std::vector<char> failed(observables.size(), 0);
  std::vector<Term> terms;

  typename Fuser::Parameter param;
  param.max_fused_size = 6;

  std::size_t expected = 0, single = 0;
  for (const auto& strings : observables) {
    expected += strings.size();
    for (const auto& str : strings) single += str.ops.size() == 1 && str.ops[0].qubits.size() == 1;
  }
  std::vector<double> moments;  // S00, S11, Re S01, Im S01 per qubit
  if (use_moments && single >= kMomentsThreshold) moments = simulator.OneQubitMoments(state);

  simulator.BeginExpectationBatch((unsigned) expected);

  for (std::size_t k = 0; k < observables.size(); ++k) {
    for (const auto& str : observables[k]) {
      if (str.ops.size() == 0) {
        evals[k] += str.weight;
      } else if (str.ops.size() == 1) {
        const auto& op = str.ops[0];
        if (!moments.empty() && op.qubits.size() == 1) {
          // <M> = m00 S00 + m11 S11 + m01 S01 + m10 conj(S01); matrix row-major, interleaved (lib/matrix.h:26-33)
          const double* s = &moments[4 * std::size_t{op.qubits[0]}];
          const std::complex<double> s01(s[2], s[3]);
          const auto m = [&](int i) { return std::complex<double>(op.matrix[2 * i], op.matrix[2 * i + 1]); };
          evals[k] += str.weight * (m(0) * s[0] + m(3) * s[1] + m(1) * s01 + m(2) * std::conj(s01));
          continue;
        }
        (void) simulator.ExpectationValue(op.qubits, op.matrix.data(), state);
        terms.push_back({k, str.weight});
      } else {
        auto fused_gates = Fuser::FuseGates(param, state.num_qubits(), str.ops);
        if (fused_gates.size() != 1) {
          IO::errorf("too many fused gates; cannot compute the expectation value.\n");
          failed[k] = 1;
          break;
        }
        const auto* pg = OpGetAlternative<FusedGate<FP>>(fused_gates[0]);
        if (pg == nullptr) {
          IO::errorf("gate fusion error; cannot compute the expectation value.\n");
          failed[k] = 1;
          break;
        }
        if (pg->qubits.size() > 6) {
          IO::errorf("operator string acts on too many qubits; cannot compute the expectation value.\n");
          failed[k] = 1;
          break;
        }
        // the matrix is copied by the C ABI before the call returns (ownership)
        (void) simulator.ExpectationValue(pg->qubits, pg->matrix.data(), state);
        terms.push_back({k, str.weight});
      }
    }
  }

  const auto values = simulator.EndExpectationBatch((unsigned) terms.size());
  for (std::size_t i = 0; i < terms.size() && i < values.size(); ++i) {
    evals[terms[i].observable] += terms[i].weight * values[i];
  }
  for (std::size_t k = 0; k < observables.size(); ++k) {
    if (failed[k]) evals[k] = 0;
  }
  return evals;
}

}  // namespace qsim

#endif  // QSIM_B200_EXPECT_B200_H_
