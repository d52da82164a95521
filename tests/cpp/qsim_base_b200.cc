// qsim_base_b200 -- the reference's apps/qsim_base_cuda.cu flow on the B200 backend:
// reference parser + MultiQubitGateFuser + QSimRunner, unchanged, with
// Factory = {SimulatorB200, StateSpaceB200}.  Prints the first 8 amplitudes in the
// same format as apps/qsim_base.cc:93-107.
//   usage: qsim_base_b200 -c circuit -d maxtime -s seed -f max_fused_size -v verbosity
#include <unistd.h>

#include <algorithm>
#include <complex>
#include <cstdlib>
#include <limits>
#include <string>

#include "circuit_qsim_parser.h"
#include "fuser_mqubit.h"
#include "gates_qsim.h"
#include "io_file.h"
#include "run_qsim.h"

#include "factory_b200.h"

int main(int argc, char* argv[]) {
  using namespace qsim;
  std::string circuit_file;
  unsigned maxtime = std::numeric_limits<unsigned>::max();
  unsigned seed = 1, max_fused_size = 2, verbosity = 0;
  int k;
  while ((k = getopt(argc, argv, "c:d:s:f:v:")) != -1) {
    switch (k) {
      case 'c': circuit_file = optarg; break;
      case 'd': maxtime = std::atoi(optarg); break;
      case 's': seed = std::atoi(optarg); break;
      case 'f': max_fused_size = std::atoi(optarg); break;
      case 'v': verbosity = std::atoi(optarg); break;
      default: IO::errorf("usage: qsim_base_b200 -c circuit -d maxtime -s seed -f max_fused_size -v verbosity\n"); return 1;
    }
  }
  if (circuit_file.empty()) {
    IO::errorf("circuit file is not provided.\n");
    return 1;
  }

  Circuit<Operation<float>> circuit;
  if (!CircuitQsimParser<IOFile>::FromFile(maxtime, circuit_file, circuit)) return 1;

  using Factory = qsim::Factory<float>;
  using StateSpace = Factory::StateSpace;
  using Fuser = MultiQubitGateFuser<IO>;
  using Runner = QSimRunner<IO, Fuser, Factory>;

  Factory factory;
  StateSpace state_space = factory.CreateStateSpace();
  auto state = state_space.Create(circuit.num_qubits);
  if (state_space.IsNull(state)) {
    IO::errorf("not enough memory: is the number of qubits too large?\n");
    return 1;
  }
  state_space.SetStateZero(state);

  Runner::Parameter param;
  param.max_fused_size = max_fused_size;
  param.seed = seed;
  param.verbosity = verbosity;

  if (Runner::Run(param, factory, circuit, state)) {
    static constexpr char const* bits[8] = {"000", "001", "010", "011", "100", "101", "110", "111"};
    uint64_t size = std::min(uint64_t{8}, uint64_t{1} << circuit.num_qubits);
    unsigned s = 3 - std::min(unsigned{3}, circuit.num_qubits);
    for (uint64_t i = 0; i < size; ++i) {
      auto a = state_space.GetAmpl(state, i);
      IO::messagef("%s:%16.8g%16.8g%16.8g\n", bits[i] + s, std::real(a), std::imag(a), std::norm(a));
    }
  }
  return 0;
}
