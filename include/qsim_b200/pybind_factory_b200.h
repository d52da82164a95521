// pybind_factory_b200.h -- the type bundle the reference's pybind layer expects under the name
// qsim::Factory (pybind_interface/pybind_main.cpp uses Factory::{Simulator, StateSpace, Gate,
// Operation, Runner, RunnerParameter, NoisyRunner, NoisyRunnerParameter}, constructs it from the
// Python options dict and asks it for state spaces and simulators).  Include after the reference's
// fuser_mqubit.h, gates_cirq.h, io.h, run_qsim.h, qtrajectory.h and pybind11.
#ifndef QSIM_B200_PYBIND_FACTORY_B200_H_
#define QSIM_B200_PYBIND_FACTORY_B200_H_

#include "simulator_b200.h"

namespace qsim {
namespace b200 {

template <typename FP, typename OptionsDict>
class PybindFactory {
 public:
  using Simulator = SimulatorB200<FP>;
  using StateSpace = typename Simulator::StateSpace;
  using Gate = Cirq::GateCirq<FP>;
  using Operation = qsim::Operation<FP>;
  using Runner = QSimRunner<IO, MultiQubitGateFuser<IO>, PybindFactory>;
  using RunnerParameter = typename Runner::Parameter;
  using NoisyRunner = QuantumTrajectorySimulator<IO, Runner>;
  using NoisyRunnerParameter = typename NoisyRunner::Parameter;

  // The CUDA module reads "gsst" / "gdb" (threads and dblocks of its state-space kernels) here;
  // launch shapes are chosen by libqsim_b200, so the options are accepted and not consulted.
  explicit PybindFactory(const OptionsDict&) {}

  Simulator CreateSimulator() const { return Simulator(); }
  StateSpace CreateStateSpace() const { return StateSpace(); }
};

#ifdef QSIM_B200_RUN_B200_H_
// The same bundle on a state sharded over several GPUs (include run_b200.h and simulator_b200_sharded.h first):
// Runner = B200Runner, the hook pybind_interface/custatevecex/pybind_main_custatevecex.cpp:77-86 uses for its own
// backend-specific runner.  Options: "gnd" = number of GPUs (a power of two; default: all visible ones),
// "gswap" = exchange mode (qb200_sv_set_option "swap_mode").
template <typename FP, typename OptionsDict>
class PybindShardedFactory {
 public:
  using Simulator = SimulatorB200Sharded<FP>;
  using StateSpace = typename Simulator::StateSpace;
  using Gate = Cirq::GateCirq<FP>;
  using Operation = qsim::Operation<FP>;
  using Runner = B200Runner<IO, MultiQubitGateFuser<IO>, PybindShardedFactory>;
  using RunnerParameter = typename Runner::Parameter;
  using NoisyRunner = QuantumTrajectorySimulator<IO, Runner>;
  using NoisyRunnerParameter = typename NoisyRunner::Parameter;

  explicit PybindShardedFactory(const OptionsDict& options) {
    int count = 1;
    qb200_device_count(&count);
    unsigned want = 0;
    if (options.contains("gnd")) want = options["gnd"].template cast<unsigned>();
    if (want == 0) want = (unsigned) (count < 1 ? 1 : count);
    unsigned p = 1;
    while (2 * p <= want) p *= 2;
    for (unsigned r = 0; r < p; ++r) param_.devices.push_back((int) (r % (unsigned) (count < 1 ? 1 : count)));
    if (options.contains("gswap")) param_.swap_mode = options["gswap"].template cast<int>();
  }

  Simulator CreateSimulator() const { return Simulator(); }
  StateSpace CreateStateSpace() const { return StateSpace(param_); }

 private:
  ShardedParameter param_;
};
#endif  // QSIM_B200_RUN_B200_H_

}  // namespace b200
}  // namespace qsim

#endif  // QSIM_B200_PYBIND_FACTORY_B200_H_
