// qsim_base_b200 -- the reference's apps/qsim_base_cuda.cu flow on the B200 backend: reference parser +
// MultiQubitGateFuser, unchanged, with B200Runner (include/qsim_b200/run_b200.h).  Prints the first 8
// amplitudes in the format of apps/qsim_base.cc:93-107.
//   usage: qsim_base_b200 -c circuit -d maxtime -s seed -f max_fused_size -v verbosity
//                         [-g shards]   2^k shards over the visible GPUs (round-robin when there are fewer GPUs)
//                         [-o file]     dump the final state: float32 (re, im) pairs in normal order, the layout
//                                       release_state_to_python hands to Python (pybind_main.cpp:431-460)
//                         [-i file]     start from a dumped state instead of |0...0>
//                         [-m n]        sample n bitstrings from the final state (seed -s) and print the first 8
#include <unistd.h>

#include <algorithm>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <string>
#include <vector>

#include "circuit_qsim_parser.h"
#include "fuser_mqubit.h"
#include "gates_qsim.h"
#include "io_file.h"

#include "qsim_b200/run_b200.h"
#include "qsim_b200/simulator_b200.h"
#include "qsim_b200/simulator_b200_sharded.h"

namespace {

struct Options {
  std::string circuit_file, dump_file, load_file;
  unsigned maxtime = std::numeric_limits<unsigned>::max();
  unsigned seed = 1, max_fused_size = 2, verbosity = 0, shards = 1;
  uint64_t samples = 0;
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

bool ReadFile(const std::string& path, std::vector<float>& buf) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  const size_t got = std::fread(buf.data(), sizeof(float), buf.size(), f);
  std::fclose(f);
  return got == buf.size();
}

bool WriteFile(const std::string& path, const std::vector<float>& buf) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) return false;
  const size_t put = std::fwrite(buf.data(), sizeof(float), buf.size(), f);
  return std::fclose(f) == 0 && put == buf.size();
}

template <typename Factory>
int Main(const Options& opt, const Factory& factory) {
  using namespace qsim;
  using StateSpace = typename Factory::StateSpace;
  using Fuser = MultiQubitGateFuser<IO>;
  using Runner = B200Runner<IO, Fuser, Factory>;

  Circuit<Operation<float>> circuit;
  if (!CircuitQsimParser<IOFile>::FromFile(opt.maxtime, opt.circuit_file, circuit)) return 1;

  StateSpace state_space = factory.CreateStateSpace();
  auto state = state_space.Create(circuit.num_qubits);
  if (state_space.IsNull(state)) {
    IO::errorf("not enough memory: is the number of qubits too large?\n");
    return 1;
  }
  if (opt.load_file.empty()) {
    state_space.SetStateZero(state);
  } else {
    std::vector<float> host(StateSpace::MinSize(circuit.num_qubits));
    if (!ReadFile(opt.load_file, host)) {
      IO::errorf("cannot read %s (expected %lu float32 values).\n", opt.load_file.c_str(), host.size());
      return 1;
    }
    state_space.Copy(host.data(), state);
  }

  typename Runner::Parameter param;
  param.max_fused_size = opt.max_fused_size;
  param.seed = opt.seed;
  param.verbosity = opt.verbosity;
  if (!Runner::Run(param, factory, circuit, state)) return 1;

  static constexpr char const* bits[8] = {"000", "001", "010", "011", "100", "101", "110", "111"};
  const uint64_t size = std::min(uint64_t{8}, uint64_t{1} << circuit.num_qubits);
  const unsigned s = 3 - std::min(unsigned{3}, circuit.num_qubits);
  for (uint64_t i = 0; i < size; ++i) {
    auto a = state_space.GetAmpl(state, i);
    IO::messagef("%s:%16.8g%16.8g%16.8g\n", bits[i] + s, std::real(a), std::imag(a), std::norm(a));
  }
  if (opt.samples > 0) {
    auto bitstrings = state_space.Sample(state, opt.samples, opt.seed);
    for (uint64_t i = 0; i < std::min<uint64_t>(8, bitstrings.size()); ++i)
      IO::messagef("sample %lu: %lu\n", i, bitstrings[i]);
  }
  if (!opt.dump_file.empty()) {
    std::vector<float> host(StateSpace::MinSize(circuit.num_qubits));
    state_space.Copy(state, host.data());
    if (!WriteFile(opt.dump_file, host)) {
      IO::errorf("cannot write %s.\n", opt.dump_file.c_str());
      return 1;
    }
  }
  return 0;
}

}  // namespace

int main(int argc, char* argv[]) {
  constexpr char usage[] = "usage: qsim_base_b200 -c circuit -d maxtime -s seed -f max_fused_size -v verbosity "
                           "[-g shards] [-o dump_file] [-i load_file] [-m num_samples]\n";
  Options opt;
  int k;
  while ((k = getopt(argc, argv, "c:d:s:f:v:g:o:i:m:")) != -1) {
    switch (k) {
      case 'c': opt.circuit_file = optarg; break;
      case 'd': opt.maxtime = std::atoi(optarg); break;
      case 's': opt.seed = std::atoi(optarg); break;
      case 'f': opt.max_fused_size = std::atoi(optarg); break;
      case 'v': opt.verbosity = std::atoi(optarg); break;
      case 'g': opt.shards = std::atoi(optarg); break;
      case 'o': opt.dump_file = optarg; break;
      case 'i': opt.load_file = optarg; break;
      case 'm': opt.samples = std::strtoull(optarg, nullptr, 10); break;
      default: qsim::IO::errorf(usage); return 1;
    }
  }
  if (opt.circuit_file.empty()) {
    qsim::IO::errorf("circuit file is not provided.\n");
    qsim::IO::errorf(usage);
    return 1;
  }
  if (opt.shards < 1 || (opt.shards & (opt.shards - 1))) {
    qsim::IO::errorf("the number of shards must be a power of two.\n");
    return 1;
  }
  if (opt.shards == 1) return Main(opt, SingleFactory<float>());
  return Main(opt, ShardedFactory<float>(opt.shards));
}
