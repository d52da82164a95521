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

}  // namespace b200
}  // namespace qsim

#endif  // QSIM_B200_PYBIND_FACTORY_B200_H_
