// run_b200.h -- B200Runner: QSimRunner's Run overloads (lib/run_qsim.h:33-330) for the B200 backends.
// Precedent for a backend-specific runner: CuStateVecExRunner (lib/run_custatevecex.h:37), wired through
// Factory::Runner by pybind_interface/custatevecex/pybind_main_custatevecex.cpp:77-86.
//
// What changes against QSimRunner: the fused circuit is not fed one gate at a time.  Stretches of fused gates
// between measurement gates / measurement times are handed to the simulator as ONE list
// (Simulator::RunGates), so a sharded state plans its local<->global exchanges over the whole stretch
// (csrc/sv_plan.h) -- the role custatevecExSVUpdaterEnqueueMatrix + Apply play in the reference runner
// (lib/run_custatevecex.h:243-305).  A simulator without RunGates gets the plain per-gate loop.
#ifndef QSIM_B200_RUN_B200_H_
#define QSIM_B200_RUN_B200_H_

#include <cstdint>
#include <random>
#include <type_traits>
#include <vector>

#include "../qsim_b200.h"
#include "circuit.h"         // reference
#include "gate.h"            // reference
#include "gate_appl.h"       // reference
#include "operation_base.h"  // reference
#include "util.h"            // reference

namespace qsim {

namespace b200 {
template <typename S, typename = void> struct HasRunGates : std::false_type {};
template <typename S> struct HasRunGates<S, std::enable_if_t<S::kHasRunGates>> : std::true_type {};
}  // namespace b200

template <typename IO, typename Fuser, typename Factory, typename RGen = std::mt19937>
struct B200Runner final {
 public:
  using Simulator = typename Factory::Simulator;
  using StateSpace = typename Simulator::StateSpace;
  using State = typename StateSpace::State;
  using MeasurementResult = typename StateSpace::MeasurementResult;
  using fp_type = typename Simulator::fp_type;

  struct Parameter : public Fuser::Parameter {
    uint64_t seed;
  };

  // lib/run_qsim.h:61-66
  template <typename Circuit, typename MeasurementFunc>
  static bool Run(const Parameter& param, const Factory& factory, const Circuit& circuit, MeasurementFunc measure) {
    unsigned time = OpTime(circuit.ops.back());
    return Run(param, factory, {time}, circuit, measure);
  }

  // lib/run_qsim.h:78-160
  template <typename Circuit, typename MeasurementFunc>
  static bool Run(const Parameter& param, const Factory& factory, const std::vector<unsigned>& times_to_measure_at,
                  const Circuit& circuit, MeasurementFunc measure) {
    double t0 = 0.0, t1 = 0.0;
    if (param.verbosity > 1) t0 = GetTime();
    RGen rgen(param.seed);
    StateSpace state_space = factory.CreateStateSpace();
    auto state = state_space.Create(circuit.num_qubits);
    if (state_space.IsNull(state)) {
      IO::errorf("not enough memory: is the number of qubits too large?\n");
      return false;
    }
    state_space.SetStateZero(state);
    Simulator simulator = factory.CreateSimulator();
    if (param.verbosity > 1) {
      t1 = GetTime();
      IO::messagef("init time is %g seconds.\n", t1 - t0);
      t0 = GetTime();
    }
    const auto& ops = Operations<Circuit>::get(circuit);
    auto fused_ops = Fuser::FuseGates(param, circuit.num_qubits, ops, times_to_measure_at);
    if (fused_ops.size() == 0 && circuit.ops.size() > 0) return false;
    if (param.verbosity > 1) {
      t1 = GetTime();
      IO::messagef("fuse time is %g seconds.\n", t1 - t0);
    }
    if (param.verbosity > 0) t0 = GetTime();

    unsigned cur_time_index = 0;
    std::vector<MeasurementResult> discarded;
    std::size_t begin = 0;
    for (std::size_t i = 0; i < fused_ops.size(); ++i) {
      unsigned t = times_to_measure_at[cur_time_index];
      if (i == fused_ops.size() - 1 || t < OpTime(fused_ops[i + 1])) {
        if (!ApplyRange(state_space, simulator, fused_ops, begin, i + 1, rgen, state, discarded)) {
          IO::errorf("measurement failed.\n");
          return false;
        }
        begin = i + 1;
        measure(cur_time_index, state_space, state);
        ++cur_time_index;
      }
    }
    if (param.verbosity > 0) {
      state_space.DeviceSync();
      double t2 = GetTime();
      IO::messagef("time is %g seconds.\n", t2 - t0);
    }
    return true;
  }

  // lib/run_qsim.h:178-186
  template <typename Circuit>
  static bool Run(const Parameter& param, const Factory& factory, const Circuit& circuit, State& state,
                  std::vector<MeasurementResult>& measure_results) {
    StateSpace state_space = factory.CreateStateSpace();
    Simulator simulator = factory.CreateSimulator();
    return Run(param, circuit, state_space, simulator, state, measure_results);
  }

  // lib/run_qsim.h:199-210
  template <typename Circuit>
  static bool Run(const Parameter& param, const Factory& factory, const Circuit& circuit, State& state) {
    StateSpace state_space = factory.CreateStateSpace();
    Simulator simulator = factory.CreateSimulator();
    std::vector<MeasurementResult> discarded_results;
    return Run(param, circuit, state_space, simulator, state, discarded_results);
  }

  // lib/run_qsim.h:226-289
  template <typename Circuit>
  static bool Run(const Parameter& param, const Circuit& circuit, const StateSpace& state_space,
                  const Simulator& simulator, State& state, std::vector<MeasurementResult>& measure_results) {
    double t0 = 0.0, t1 = 0.0;
    if (param.verbosity > 1) t0 = GetTime();
    RGen rgen(param.seed);
    if (param.verbosity > 1) {
      t1 = GetTime();
      IO::messagef("init time is %g seconds.\n", t1 - t0);
      t0 = GetTime();
    }
    const auto& ops = Operations<Circuit>::get(circuit);
    auto fused_ops = Fuser::FuseGates(param, state.num_qubits(), ops);
    if (fused_ops.size() == 0 && ops.size() > 0) return false;
    measure_results.reserve(fused_ops.size());
    if (param.verbosity > 1) {
      t1 = GetTime();
      IO::messagef("fuse time is %g seconds.\n", t1 - t0);
    }
    if (param.verbosity > 0) t0 = GetTime();
    if (!ApplyRange(state_space, simulator, fused_ops, 0, fused_ops.size(), rgen, state, measure_results)) {
      IO::errorf("measurement failed.\n");
      return false;
    }
    if (param.verbosity > 0) {
      state_space.DeviceSync();
      double t2 = GetTime();
      IO::messagef("simu time is %g seconds.\n", t2 - t0);
    }
    return true;
  }

  // lib/run_qsim.h:305-315
  template <typename Circuit>
  static bool Run(const Parameter& param, const Circuit& circuit, const StateSpace& state_space,
                  const Simulator& simulator, State& state) {
    std::vector<MeasurementResult> discarded_results;
    return Run(param, circuit, state_space, simulator, state, discarded_results);
  }

 private:
  // fused_ops[begin, end): gates go to the simulator in stretches, measurement gates one by one in between
  template <typename FusedOps>
  static bool ApplyRange(const StateSpace& state_space, const Simulator& simulator, const FusedOps& fused_ops,
                         std::size_t begin, std::size_t end, RGen& rgen, State& state,
                         std::vector<MeasurementResult>& mresults) {
    if constexpr (!b200::HasRunGates<Simulator>::value) {
      for (std::size_t i = begin; i < end; ++i)
        if (!ApplyGate(state_space, simulator, fused_ops[i], rgen, state, mresults)) return false;
      return true;
    } else {
      std::vector<qb200_gate> batch;
      auto flush = [&]() {
        if (!batch.empty()) simulator.RunGates(batch, state);
        batch.clear();
      };
      for (std::size_t i = begin; i < end; ++i) {
        const auto& op = fused_ops[i];
        if (OpGetAlternative<Measurement>(op)) {
          flush();
          if (!ApplyGate(state_space, simulator, op, rgen, state, mresults)) return false;
        } else if (const auto* pg = OpGetAlternative<Gate<fp_type>>(op)) {
          batch.push_back({(unsigned) pg->qubits.size(), pg->qubits.data(), 0, nullptr, 0, pg->matrix.data()});
        } else if (const auto* pg = OpGetAlternative<FusedGate<fp_type>>(op)) {
          batch.push_back({(unsigned) pg->qubits.size(), pg->qubits.data(), 0, nullptr, 0, pg->matrix.data()});
        } else if (const auto* pg = OpGetAlternative<ControlledGate<fp_type>>(op)) {
          batch.push_back({(unsigned) pg->qubits.size(), pg->qubits.data(), (unsigned) pg->controlled_by.size(),
                           pg->controlled_by.data(), pg->cmask, pg->matrix.data()});
        } else {
          flush();
          ApplyGate(simulator, op, state);
        }
      }
      flush();
      return true;
    }
  }
};

}  // namespace qsim

#endif  // QSIM_B200_RUN_B200_H_
