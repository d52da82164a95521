// pybind_main_b200.cpp -- the `qsim_b200_py` Python extension: the reference's whole pybind layer
// (pybind_interface/pybind_main.{h,cpp}: circuit building, qsim_simulate*, qsim_sample*,
// qtrajectory_*, expectation values, qsimh_simulate) compiled unchanged, in place, on top of the
// B200 backend.  It plays the role of pybind_interface/cuda/pybind_main_cuda.cpp; qsimcirq would
// select this module exactly like `qsim_cuda` (qsimcirq/qsim_simulator.py:186-217).
#include "pybind_main.h"  // reference: pybind_interface/pybind_main.h (declarations + binding macros)

// MODULE_BINDINGS = GPU_MODULE_BINDINGS + the circuit-building entry points (Circuit, OpString,
// add_gate, ...), so the module is usable on its own; qsimcirq builds circuits with its CPU module
// and could use either.
PYBIND11_MODULE(qsim_b200_py, m) { MODULE_BINDINGS }

#include "fuser_mqubit.h"
#include "gates_cirq.h"
#include "io.h"
#include "qtrajectory.h"
#include "run_qsim.h"

#include "qsim_b200/pybind_factory_b200.h"

namespace qsim {

using Factory = b200::PybindFactory<float, py::dict>;
using Simulator = Factory::Simulator;

// the CPU modules toggle FTZ/DAZ in the MXCSR around a run; nothing to do for a device backend
inline void SetFlushToZeroAndDenormalsAreZeros() {}
inline void ClearFlushToZeroAndDenormalsAreZeros() {}

}  // namespace qsim

#include "pybind_main.cpp"  // reference: pybind_interface/pybind_main.cpp (definitions)
