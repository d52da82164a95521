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
 * The host half of lib/expect.h:106-151 done ONCE for a set of observables: every operator string reduced to
 * (qubits, matrix, weight) -- single operators as they are, longer strings fused with max_fused_size = 6.
 * A driver that evaluates the same observables on many states (Monte-Carlo trajectories) builds the plan once
 * instead of re-fusing every string per state.
 */
template <typename FP>
struct ObservablePlan {
  struct Term {
    std::size_t observable;
    std::complex<double> weight;
    std::vector<unsigned> qubits;
    std::vector<FP> matrix;  // row-major, interleaved (lib/matrix.h:26-33)
  };
  std::size_t num_observables = 0;
  std::vector<std::complex<double>> constant;  // sum of the weights of empty strings, per observable
  std::vector<char> failed;                    // strings that do not fuse into one gate of <= 6 qubits: value 0
  std::vector<Term> terms;
  std::size_t single_qubit_terms = 0;
};

template <typename IO, typename Fuser, typename FP>
ObservablePlan<FP> MakeObservablePlan(const std::vector<std::vector<OpString<FP>>>& observables,
                                      unsigned num_qubits) {
  ObservablePlan<FP> plan;
  plan.num_observables = observables.size();
  plan.constant.assign(observables.size(), 0);
  plan.failed.assign(observables.size(), 0);

  typename Fuser::Parameter param;
  param.max_fused_size = 6;

  for (std::size_t k = 0; k < observables.size(); ++k) {
    for (const auto& str : observables[k]) {
      if (str.ops.size() == 0) {
        plan.constant[k] += str.weight;
      } else if (str.ops.size() == 1) {
        const auto& op = str.ops[0];
        plan.terms.push_back({k, str.weight, op.qubits, std::vector<FP>(op.matrix.begin(), op.matrix.end())});
        plan.single_qubit_terms += op.qubits.size() == 1;
      } else {
        auto fused_gates = Fuser::FuseGates(param, num_qubits, str.ops);
        if (fused_gates.size() != 1) {
          IO::errorf("too many fused gates; cannot compute the expectation value.\n");
          plan.failed[k] = 1;
          break;
        }
        const auto* pg = OpGetAlternative<FusedGate<FP>>(fused_gates[0]);
        if (pg == nullptr) {
          IO::errorf("gate fusion error; cannot compute the expectation value.\n");
          plan.failed[k] = 1;
          break;
        }
        if (pg->qubits.size() > 6) {
          IO::errorf("operator string acts on too many qubits; cannot compute the expectation value.\n");
          plan.failed[k] = 1;
          break;
        }
        plan.terms.push_back({k, str.weight, pg->qubits, std::vector<FP>(pg->matrix.begin(), pg->matrix.end())});
        plan.single_qubit_terms += pg->qubits.size() == 1;
      }
    }
  }
  return plan;
}

/**
 * Evaluates a plan on a state: single-qubit operators from the reduced density matrices when there are at
 * least kMomentsThreshold of them (and `use_moments`), everything else as one batch of read-only passes read
 * after one stream synchronisation.
 */
template <typename FP, typename Simulator>
std::vector<std::complex<double>> ExpectationValues(
    const ObservablePlan<FP>& plan, const Simulator& simulator, const typename Simulator::State& state,
    bool use_moments = true) {
  std::vector<std::complex<double>> evals(plan.constant);
  std::vector<double> moments;  // S00, S11, Re S01, Im S01 per qubit
  if (use_moments && plan.single_qubit_terms >= kMomentsThreshold) moments = simulator.OneQubitMoments(state);

  std::vector<std::size_t> queued;
  simulator.BeginExpectationBatch((unsigned) plan.terms.size());
  for (std::size_t i = 0; i < plan.terms.size(); ++i) {
    const auto& t = plan.terms[i];
    if (!moments.empty() && t.qubits.size() == 1) {
      // <M> = m00 S00 + m11 S11 + m01 S01 + m10 conj(S01)
      const double* s = &moments[4 * std::size_t{t.qubits[0]}];
      const std::complex<double> s01(s[2], s[3]);
      const auto m = [&](int j) { return std::complex<double>(t.matrix[2 * j], t.matrix[2 * j + 1]); };
      evals[t.observable] += t.weight * (m(0) * s[0] + m(3) * s[1] + m(1) * s01 + m(2) * std::conj(s01));
      continue;
    }
    // the matrix is copied by the C ABI before the call returns (ownership)
    (void) simulator.ExpectationValue(t.qubits, t.matrix.data(), state);
    queued.push_back(i);
  }
  const auto values = simulator.EndExpectationBatch((unsigned) queued.size());
  for (std::size_t i = 0; i < queued.size() && i < values.size(); ++i) {
    const auto& t = plan.terms[queued[i]];
    evals[t.observable] += t.weight * values[i];
  }
  for (std::size_t k = 0; k < plan.num_observables; ++k) {
    if (plan.failed[k]) evals[k] = 0;
  }
  return evals;
}

/**
 * Expectation values of several observables, each a sum of weighted operator strings
 * (argument meaning as in lib/expect.h:95-104).  An observable whose strings cannot be
 * fused into one gate of at most six qubits reports 0, like the reference.
 */
template <typename IO, typename Fuser, typename FP, typename Simulator>
std::vector<std::complex<double>> ExpectationValues(
    const std::vector<std::vector<OpString<FP>>>& observables,
    const Simulator& simulator, const typename Simulator::State& state, bool use_moments = true) {
  return ExpectationValues(MakeObservablePlan<IO, Fuser>(observables, state.num_qubits()), simulator, state,
                           use_moments);
}

}  // namespace qsim

#endif  // QSIM_B200_EXPECT_B200_H_
