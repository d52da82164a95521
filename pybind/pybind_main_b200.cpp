// pybind_main_b200.cpp -- the `qsim_b200` Python extension: the reference's whole pybind layer
// (pybind_interface/pybind_main.{h,cpp}: circuit building, qsim_simulate*, qsim_sample*,
// qtrajectory_*, expectation values, qsimh_simulate) compiled unchanged, in place, on top of
// the B200 backend.  Sibling of pybind_interface/cuda/pybind_main_cuda.cpp:23-57; qsimcirq
// would select this module exactly like `qsim_cuda` (qsimcirq/qsim_simulator.py:186-217).
// The GPU options "gsst" / "gdb" (state-space threads / dblocks of the reference CUDA backend)
// are accepted and ignored: launch shapes are chosen by the library.
#include "pybind_main.h"  // reference: pybind_interface/pybind_main.h

// MODULE_BINDINGS = GPU_MODULE_BINDINGS + the circuit-building entry points (Circuit, OpString, add_gate, ...),
// so the module is usable on its own; qsimcirq builds circuits with its CPU module and could use either.
PYBIND11_MODULE(qsim_b200_py, m) { MODULE_BINDINGS }

#include "fuser_mqubit.h"
#include "gates_cirq.h"
#include "io.h"
#include "run_qsim.h"

#include "qsim_b200/simulator_b200.h"

namespace qsim {
using Simulator = SimulatorB200<float>;

struct Factory {
  explicit Factory(const py::dict&) {}

  using Simulator = qsim::Simulator;
  using StateSpace = Simulator::StateSpace;

  using Gate = Cirq::GateCirq<float>;
  using Operation = qsim::Operation<float>;
  using Runner = QSimRunner<IO, MultiQubitGateFuser<IO>, Factory>;
  using RunnerParameter = Runner::Parameter;
  using NoisyRunner = qsim::QuantumTrajectorySimulator<IO, Runner>;
  using NoisyRunnerParameter = NoisyRunner::Parameter;

  StateSpace CreateStateSpace() const { return StateSpace(); }
  Simulator CreateSimulator() const { return Simulator(); }
};

inline void SetFlushToZeroAndDenormalsAreZeros() {}
inline void ClearFlushToZeroAndDenormalsAreZeros() {}
}  // namespace qsim

#include "pybind_main.cpp"  // reference: pybind_interface/pybind_main.cpp
