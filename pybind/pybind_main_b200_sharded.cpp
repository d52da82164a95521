// pybind_main_b200_sharded.cpp -- the `qsim_b200_sharded_py` Python extension: the reference's pybind layer
// (pybind_interface/pybind_main.{h,cpp}, compiled unchanged, in place) on a state SHARDED over several B200s.
// It plays the role of pybind_interface/custatevecex/pybind_main_custatevecex.cpp: a multi-device backend with
// its own runner (Factory::Runner = B200Runner, which hands whole stretches of fused gates to the library so the
// local<->global exchanges are planned over them).  Options: the usual ones plus "gnd" (number of GPUs).
#include "pybind_main.h"  // reference: pybind_interface/pybind_main.h

// GPU_MODULE_BINDINGS: the simulation entry points only, like qsim_cuda / qsim_custatevecex -- circuits are built
// with a module that registers the circuit classes (qsim_b200_py or the reference's CPU modules), exactly as
// qsimcirq does (two modules cannot both register Circuit / OpString in one interpreter).
PYBIND11_MODULE(qsim_b200_sharded_py, m) { GPU_MODULE_BINDINGS }

#include "fuser_mqubit.h"
#include "gates_cirq.h"
#include "io.h"
#include "qtrajectory.h"
#include "run_qsim.h"

#include "qsim_b200/run_b200.h"
#include "qsim_b200/simulator_b200_sharded.h"
#include "qsim_b200/pybind_factory_b200.h"

namespace qsim {

using Factory = b200::PybindShardedFactory<float, py::dict>;
using Simulator = Factory::Simulator;

inline void SetFlushToZeroAndDenormalsAreZeros() {}
inline void ClearFlushToZeroAndDenormalsAreZeros() {}

}  // namespace qsim

#include "pybind_main.cpp"  // reference: pybind_interface/pybind_main.cpp
