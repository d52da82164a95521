// qsim_amplitudes_b200 -- the reference's apps/qsim_amplitudes.cc on the B200 backend: runs a circuit, and at
// each requested time writes the amplitudes of the bitstrings listed in an input file (one bitstring per line,
// circuits/bitstrings_q*_s*) to an output file, in the reference's text format (apps/qsim_amplitudes.cc:136-150).
//   usage: qsim_amplitudes_b200 -c circuit_file -d times_to_save_results -i input_files -o output_files
//                               -s seed -f max_fused_size -v verbosity [-g shards]
#include <unistd.h>

#include <algorithm>
#include <complex>
#include <cstdlib>
#include <iomanip>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include "bitstring.h"
#include "circuit_qsim_parser.h"
#include "fuser_mqubit.h"
#include "gates_qsim.h"
#include "io_file.h"
#include "util.h"

#include "qsim_b200/run_b200.h"
#include "qsim_b200/simulator_b200.h"
#include "qsim_b200/simulator_b200_sharded.h"

namespace {

struct Options {
  std::string circuit_file;
  std::vector<unsigned> times;
  std::vector<std::string> input_files, output_files;
  unsigned seed = 1, max_fused_size = 2, verbosity = 0, shards = 1;
};

template <typename FP>
struct SingleFactory {
  using fp_type = FP;
  using Simulator = qsim::SimulatorB200<FP>;
  using StateSpace = typename Simulator::StateSpace;
  StateSpace CreateStateSpace() const { return StateSpace(); }
  Simulator CreateSimulator() const { return Simulator(); }
};

template <typename FP>
struct ShardedFactory {
  using fp_type = FP;
  using Simulator = qsim::SimulatorB200Sharded<FP>;
  using StateSpace = typename Simulator::StateSpace;
  explicit ShardedFactory(unsigned shards) {
    int count = 1;
    qb200_device_count(&count);
    for (unsigned r = 0; r < shards; ++r) param.devices.push_back((int) (r % (unsigned) std::max(count, 1)));
  }
  StateSpace CreateStateSpace() const { return StateSpace(param); }
  Simulator CreateSimulator() const { return Simulator(); }
  qsim::b200::ShardedParameter param;
};

template <typename Factory>
int Main(const Options& opt, const Factory& factory) {
  using namespace qsim;
  using StateSpace = typename Factory::StateSpace;
  using State = typename StateSpace::State;
  using Runner = B200Runner<IO, MultiQubitGateFuser<IO>, Factory>;

  Circuit<Operation<float>> circuit;
  if (!CircuitQsimParser<IOFile>::FromFile(opt.times.back(), opt.circuit_file, circuit)) return 1;

  bool failed = false;
  auto measure = [&](unsigned k, const StateSpace& state_space, const State& state) {
    std::vector<Bitstring> bitstrings;
    BitstringsFromFile<IOFile>(circuit.num_qubits, opt.input_files[k], bitstrings);
    if (bitstrings.empty()) { failed = true; return; }
    std::stringstream ss;
    const unsigned width = 2 * sizeof(float) + 1;
    ss << std::setprecision(width);
    for (const auto& b : bitstrings) {
      auto a = state_space.GetAmpl(state, b);
      ss << std::setw(width + 8) << std::real(a) << std::setw(width + 8) << std::imag(a) << "\n";
    }
    if (!IOFile::WriteToFile(opt.output_files[k], ss.str())) failed = true;
  };

  typename Runner::Parameter param;
  param.max_fused_size = opt.max_fused_size;
  param.seed = opt.seed;
  param.verbosity = opt.verbosity;
  if (!Runner::Run(param, factory, opt.times, circuit, measure) || failed) return 1;
  IO::messagef("all done.\n");
  return 0;
}

}  // namespace

int main(int argc, char* argv[]) {
  constexpr char usage[] = "usage:\n  ./qsim_amplitudes_b200 -c circuit_file -d times_to_save_results -i input_files "
                           "-o output_files -s seed -f max_fused_size -v verbosity [-g shards]\n";
  Options opt;
  auto to_int = [](const std::string& word) -> unsigned { return std::atoi(word.c_str()); };
  int k;
  while ((k = getopt(argc, argv, "c:d:i:s:o:f:v:g:")) != -1) {
    switch (k) {
      case 'c': opt.circuit_file = optarg; break;
      case 'd': qsim::SplitString(optarg, ',', to_int, opt.times); break;
      case 'i': qsim::SplitString(optarg, ',', opt.input_files); break;
      case 'o': qsim::SplitString(optarg, ',', opt.output_files); break;
      case 's': opt.seed = std::atoi(optarg); break;
      case 'f': opt.max_fused_size = std::atoi(optarg); break;
      case 'v': opt.verbosity = std::atoi(optarg); break;
      case 'g': opt.shards = std::atoi(optarg); break;
      default: qsim::IO::errorf(usage); return 1;
    }
  }
  if (opt.times.empty()) opt.times.push_back(std::numeric_limits<unsigned>::max());
  if (opt.circuit_file.empty() || opt.input_files.empty() || opt.output_files.empty()) {
    qsim::IO::errorf("circuit file, input files and output files must be given.\n");
    qsim::IO::errorf(usage);
    return 1;
  }
  if (opt.times.size() != opt.input_files.size() || opt.times.size() != opt.output_files.size()) {
    qsim::IO::errorf("the number of times is not the same as the number of input or output files.\n");
    return 1;
  }
  for (std::size_t i = 1; i < opt.times.size(); ++i) {
    if (opt.times[i - 1] >= opt.times[i]) {
      qsim::IO::errorf("times to save results must be sorted and distinct.\n");
      return 1;
    }
  }
  if (opt.shards < 1 || (opt.shards & (opt.shards - 1))) {
    qsim::IO::errorf("the number of shards must be a power of two.\n");
    return 1;
  }
  if (opt.shards == 1) return Main(opt, SingleFactory<float>());
  return Main(opt, ShardedFactory<float>(opt.shards));
}
