// simulator_b200_sharded.h -- Simulator for a state sharded over several B200s: the counterpart of
// lib/simulator_custatevecex.h:35-200 over the qb200_sv_* C ABI.  Gates may touch any qubits as long as the
// targets fit one shard (:67-71); a target that currently lives on a rank bit is swapped in first.
#ifndef QSIM_B200_SIMULATOR_B200_SHARDED_H_
#define QSIM_B200_SIMULATOR_B200_SHARDED_H_

#include <complex>
#include <cstdint>
#include <vector>

#include "statespace_b200_sharded.h"

namespace qsim {

template <typename FP = float>
class SimulatorB200Sharded final {
 public:
  using StateSpace = StateSpaceB200Sharded<FP>;
  using State = typename StateSpace::State;
  using fp_type = typename StateSpace::fp_type;

  SimulatorB200Sharded() {}

  void ApplyGate(const std::vector<unsigned>& qs, const fp_type* matrix, State& state) const {
    QB200_SV_CHECK(state.get(), qb200_sv_apply_gate(state.get(), qs.data(), (unsigned) qs.size(), matrix));
  }

  void ApplyControlledGate(const std::vector<unsigned>& qs, const std::vector<unsigned>& cqs, uint64_t cvals,
                           const fp_type* matrix, State& state) const {
    QB200_SV_CHECK(state.get(), qb200_sv_apply_controlled_gate(state.get(), qs.data(), (unsigned) qs.size(),
                                                               cqs.data(), (unsigned) cqs.size(), cvals, matrix));
  }

  // the operator's qubits are made local first (lib/simulator_custatevecex.h:147-196)
  std::complex<double> ExpectationValue(const std::vector<unsigned>& qs, const fp_type* matrix,
                                        const State& state) const {
    double out[2] = {0, 0};
    QB200_SV_CHECK(state.get(), qb200_sv_expectation_value(state.get(), qs.data(), (unsigned) qs.size(), matrix, out));
    return {out[0], out[1]};
  }

  // A whole list of fused gates at once: the library plans the exchanges over the list (look-ahead, commuting
  // gates reordered) instead of reacting gate by gate.  Used by B200Runner (run_b200.h).
  static constexpr bool kHasRunGates = true;
  void RunGates(const std::vector<qb200_gate>& gates, State& state) const {
    QB200_SV_CHECK(state.get(), qb200_sv_run(state.get(), gates.data(), gates.size()));
  }

  static unsigned SIMDRegisterSize() { return 32; }
};

}  // namespace qsim

#endif  // QSIM_B200_SIMULATOR_B200_SHARDED_H_
