"""qsim_b200 -- B200-native state-vector engine behind qsim's Simulator/StateSpace API.

Python mirror of the reference's duck-typed backend interface
(lib/simulator_cuda.h, lib/statespace_cuda.h, lib/vectorspace_cuda.h) over the
C ABI in include/qsim_b200.h.  The C++ mirror lives in include/qsim_b200/.
"""
from ._lib import F32, F64, QB200Error, load  # noqa: F401
from .backend import MeasurementResult, SimulatorB200, State, StateSpaceB200  # noqa: F401
from .trace import read_trace  # noqa: F401
