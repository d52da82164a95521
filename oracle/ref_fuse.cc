// ref_fuse.cc -- TEST/BENCH INFRASTRUCTURE.  Runs the UNMODIFIED reference
// parser (lib/circuit_qsim_parser.h) + MultiQubitGateFuser (lib/fuser_mqubit.h)
// + QSimRunner (lib/run_qsim.h) over a circuit file with a *recording* backend:
// every ApplyGate / ApplyControlledGate call the hot path would receive
// (lib/gate_appl.h:36-48) is written to a "fused-gate trace" file instead of
// being executed.  The trace is what bench.py and the parity tests replay
// through the C-ABI, so the host-side fuser stays the reference's own.
//
// usage: ref_fuse <circuit_file> <maxtime> <max_fused_size> <out.trace>
//
// Trace format (little endian):
//   char[8] "QB2TRACE"; u32 version=1; u32 num_qubits; u32 num_ops; u32 fp_bytes(4)
//   per op: u32 num_targets; u32 num_controls; u64 cvals;
//           u32 qs[num_targets]; u32 cqs[num_controls];
//           f32 matrix[2 * 4^num_targets]   (row-major, interleaved re,im)
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <string>
#include <vector>

#include "circuit_qsim_parser.h"
#include "fuser_mqubit.h"
#include "gates_qsim.h"
#include "io_file.h"
#include "run_qsim.h"
#include "statespace.h"

namespace {

struct Trace {
  std::FILE* f = nullptr;
  uint32_t num_ops = 0;
};
Trace g_trace;

template <typename T>
void put(const T& v) { std::fwrite(&v, sizeof(T), 1, g_trace.f); }

struct RecVector {
  explicit RecVector(unsigned n) : n_(n) {}
  unsigned num_qubits() const { return n_; }
  float* get() { return &dummy_; }
  const float* get() const { return &dummy_; }
  unsigned n_;
  float dummy_ = 0;
};

struct RecStateSpace {
  using fp_type = float;
  using State = RecVector;
  struct MeasurementResult {
    uint64_t mask = 0, bits = 0;
    std::vector<unsigned> bitstring;
    bool valid = false;
  };
  State Create(unsigned n) const { return State(n); }
  static bool IsNull(const State&) { return false; }
  void SetStateZero(State&) const {}
  template <typename RGen>
  MeasurementResult Measure(const std::vector<unsigned>&, RGen&, State&) const {
    std::fprintf(stderr, "ref_fuse: measurement gates are not traced\n");
    std::exit(2);
  }
  static void DeviceSync() {}
};

struct RecSimulator {
  using StateSpace = RecStateSpace;
  using State = RecVector;
  using fp_type = float;
  void ApplyGate(const std::vector<unsigned>& qs, const float* m, State&) const {
    Record(qs, {}, 0, m);
  }
  void ApplyControlledGate(const std::vector<unsigned>& qs,
                           const std::vector<unsigned>& cqs, uint64_t cvals,
                           const float* m, State&) const {
    Record(qs, cqs, cvals, m);
  }
  static void Record(const std::vector<unsigned>& qs,
                     const std::vector<unsigned>& cqs, uint64_t cvals,
                     const float* m) {
    put<uint32_t>(qs.size());
    put<uint32_t>(cqs.size());
    put<uint64_t>(cvals);
    for (auto q : qs) put<uint32_t>(q);
    for (auto q : cqs) put<uint32_t>(q);
    std::size_t len = std::size_t{2} << (2 * qs.size());
    std::fwrite(m, sizeof(float), len, g_trace.f);
    ++g_trace.num_ops;
  }
};

struct Factory {
  using Simulator = RecSimulator;
  using StateSpace = RecStateSpace;
  StateSpace CreateStateSpace() const { return StateSpace(); }
  Simulator CreateSimulator() const { return Simulator(); }
};

}  // namespace

int main(int argc, char** argv) {
  using namespace qsim;
  if (argc != 5) {
    std::fprintf(stderr, "usage: ref_fuse circuit maxtime max_fused out.trace\n");
    return 1;
  }
  unsigned maxtime = std::atoi(argv[2]);
  if (maxtime == 0) maxtime = std::numeric_limits<unsigned>::max();

  Circuit<Operation<float>> circuit;
  if (!CircuitQsimParser<IOFile>::FromFile(maxtime, argv[1], circuit)) return 1;

  g_trace.f = std::fopen(argv[4], "wb");
  if (!g_trace.f) return 1;
  std::fwrite("QB2TRACE", 1, 8, g_trace.f);
  put<uint32_t>(1);
  put<uint32_t>(circuit.num_qubits);
  long pos_ops = std::ftell(g_trace.f);
  put<uint32_t>(0);
  put<uint32_t>(4);

  using Fuser = MultiQubitGateFuser<IO>;
  using Runner = QSimRunner<IO, Fuser, Factory>;
  Runner::Parameter param;
  param.max_fused_size = std::atoi(argv[3]);
  param.seed = 1;
  param.verbosity = 0;

  RecStateSpace ss;
  auto state = ss.Create(circuit.num_qubits);
  RecSimulator sim;
  if (!Runner::Run(param, circuit, ss, sim, state)) return 1;

  std::fseek(g_trace.f, pos_ops, SEEK_SET);
  put<uint32_t>(g_trace.num_ops);
  std::fclose(g_trace.f);
  std::printf("%u qubits, %u fused ops -> %s\n", circuit.num_qubits,
              g_trace.num_ops, argv[4]);
  return 0;
}
