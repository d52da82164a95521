// qsim_qtrajectory_b200 -- BASELINE config 5: Monte-Carlo quantum trajectories of a noisy
// circuit on the B200 backend, one process per GPU over a slice [traj0, traj0 + n) of the
// repetition ids (trajectories are independent units: no collective, the observable sums
// printed by each process are added by the launcher, tools/traj_farm.py).
//
// Host side = the reference's own, unchanged, consumed in place from $(QSIM_REF)/lib:
// CircuitQsimParser, MakeNoisy (lib/circuit_noisy.h:34-96) with DepolarizingChannel
// (lib/channels_cirq.h:127-200), QuantumTrajectorySimulator::RunOnce (lib/qtrajectory.h),
// MultiQubitGateFuser, ExpectationValue<IO, Fuser> (lib/expect.h:106-151).
// Backend = include/qsim_b200/*.h (SimulatorB200 / StateSpaceB200), or -- when compiled with
// -DQTRAJ_REFERENCE_CPU for the oracle build under oracle/_ref/ -- the reference's CPU
// simulator selected by lib/simmux.h, so that the same seeds give the same Kraus choices and
// the observable sums can be compared number by number; -DQTRAJ_REFERENCE_CUDA (nvcc) puts the
// same driver on the reference's own CUDA backend (lib/simulator_cuda.h) = the GPU baseline.
// B200 backend only: "-x 1" (default) shares the noiseless prefix of the fused gate list between
// trajectories (include/qsim_b200/qtrajectory_b200.h), bit-identical sums, ~2.5x fewer gate passes at
// p = 0.001; "-C amplitude_damp" replaces the depolarizing channel by amplitude damping with gamma = -p
// (a non-unitary channel: Kraus operators sampled from expectation values, lib/qtrajectory.h:335-372).
//
// Role model: apps/qsim_qtrajectory_cuda.cu (amplitude/phase damping, X observables);
// here: depolarizing noise after every gate qubit; observables = X_q and Z_q on every qubit
// plus one 6-qubit Pauli string per window of 6 qubits.
//
//   [CUDA_VISIBLE_DEVICES=r] qsim_qtrajectory_b200 -c circuit -d maxtime -p 0.001 -0 traj0 -n num -f 4 [-v 0] [-b 2] [-j 1]
// prints one JSON line: {"n":..,"traj0":..,"num":..,"seconds":..,"sums":[re,im,...]}
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "circuit.h"
#include "channels_cirq.h"
#include "circuit_noisy.h"
#include "circuit_qsim_parser.h"
#include "expect.h"
#include "fuser_mqubit.h"
#include "gates_qsim.h"
#include "io_file.h"
#include "operation.h"
#include "qtrajectory.h"
#include "run_qsim.h"

#if defined(QTRAJ_REFERENCE_CPU)
#include "formux.h"
#include "simmux.h"
#elif defined(QTRAJ_REFERENCE_CUDA)
#include "simulator_cuda.h"
#define QTRAJ_REFERENCE_CPU  // same code path as the CPU checker: the reference's runner and expect.h
#define QTRAJ_REFERENCE_IS_CUDA
#else
#include "qsim_b200/expect_b200.h"
#include "qsim_b200/qtrajectory_b200.h"
#include "qsim_b200/simulator_b200.h"
#endif

namespace {

struct Options {
  std::string circuit_file;
  unsigned maxtime = std::numeric_limits<unsigned>::max();
  double p = 0.001;
  unsigned traj0 = 0, num = 8, max_fused_size = 4, verbosity = 0, threads = 0;
  // 2 (default): all observables of a trajectory through qsim::ExpectationValues (expect_b200.h: single-qubit
  // operators from the reduced density matrices, the rest batched, one stream synchronisation); 1: batched
  // only (same kernels and values as 0); 0: the reference's lib/expect.h, one synchronisation per string
  unsigned batch = 2;
  // worker threads of this process, each with its own state, simulator and CUDA per-thread stream, over
  // contiguous sub-slices of the repetition ids: one worker's host phases (fusing, Kraus sampling, reading
  // results) overlap the other's kernels.  B200 backend only.
  unsigned workers = 1;
  unsigned prefix = 1;            // 1: share the noiseless prefix between trajectories (B200 backend)
  std::string channel = "depolarize";  // or "amplitude_damp" (gamma = p)
  int kraus_groups = 1;                // -k: all Kraus probabilities of a non-unitary channel from one read pass
};

Options Parse(int argc, char* argv[]) {
  Options o;
  int k;
  while ((k = getopt(argc, argv, "c:d:p:0:n:f:t:v:b:j:x:C:k:")) != -1) {
    switch (k) {
      case 'c': o.circuit_file = optarg; break;
      case 'd': o.maxtime = std::atoi(optarg); break;
      case 'p': o.p = std::atof(optarg); break;
      case '0': o.traj0 = std::atoi(optarg); break;
      case 'n': o.num = std::atoi(optarg); break;
      case 'f': o.max_fused_size = std::atoi(optarg); break;
      case 't': o.threads = std::atoi(optarg); break;
      case 'v': o.verbosity = std::atoi(optarg); break;
      case 'b': o.batch = std::atoi(optarg); break;
      case 'j': o.workers = std::max(1, std::atoi(optarg)); break;
      case 'x': o.prefix = std::atoi(optarg); break;
      case 'C': o.channel = optarg; break;
      case 'k': o.kraus_groups = std::atoi(optarg); break;
      default:
        std::fprintf(stderr, "usage: %s -c circuit [-d maxtime] [-p prob] [-0 traj0] [-n num] "
                             "[-f max_fused_size] [-t threads] [-v verbosity] [-b batch] [-j workers] [-x prefix_sharing] "
                             "[-C depolarize|amplitude_damp]\n", argv[0]);
        std::exit(1);
    }
  }
  return o;
}

// X_q, Z_q for every qubit, then one Pauli string X Z Y X Z Y on each full window of 6 qubits.
template <typename FP>
std::vector<std::vector<qsim::OpString<FP>>> Observables(unsigned n) {
  using namespace qsim;
  std::vector<std::vector<OpString<FP>>> obs;
  for (unsigned q = 0; q < n; ++q) obs.push_back({{{1.0, 0.0}, {GateX<FP>::Create(0, q)}}});
  for (unsigned q = 0; q < n; ++q) obs.push_back({{{1.0, 0.0}, {GateZ<FP>::Create(0, q)}}});
  for (unsigned w = 0; w + 6 <= n; w += 6) {
    OpString<FP> s{{1.0, 0.0}, {}};
    for (unsigned j = 0; j < 6; ++j) {
      switch (j % 3) {
        case 0: s.ops.push_back(GateX<FP>::Create(0, w + j)); break;
        case 1: s.ops.push_back(GateZ<FP>::Create(0, w + j)); break;
        default: s.ops.push_back(GateY<FP>::Create(0, w + j)); break;
      }
    }
    obs.push_back({s});
  }
  return obs;
}

// Counts the passes the drivers issue (algorithmic HBM bytes of the run):
// gate pass = 16 * 2^n B, expectation pass = 8 * 2^n B in fp32.
struct PassCount { uint64_t gates = 0, expects = 0, moment_calls = 0; };
thread_local PassCount g_passes;  // per worker, added up at the end

template <typename Base>
struct Counting {
  using StateSpace = typename Base::StateSpace;
  using State = typename Base::State;
  using fp_type = typename Base::fp_type;
  template <typename... Args>
  explicit Counting(Args&&... args) : base(std::forward<Args>(args)...) {}
  void ApplyGate(const std::vector<unsigned>& qs, const fp_type* m, State& s) const {
    ++g_passes.gates;
    base.ApplyGate(qs, m, s);
  }
  void ApplyControlledGate(const std::vector<unsigned>& qs, const std::vector<unsigned>& cqs, uint64_t cvals,
                           const fp_type* m, State& s) const {
    ++g_passes.gates;
    base.ApplyControlledGate(qs, cqs, cvals, m, s);
  }
  std::complex<double> ExpectationValue(const std::vector<unsigned>& qs, const fp_type* m, const State& s) const {
    ++g_passes.expects;
    return base.ExpectationValue(qs, m, s);
  }
  static unsigned SIMDRegisterSize() { return Base::SIMDRegisterSize(); }
#ifndef QTRAJ_REFERENCE_CPU
  void BeginExpectationBatch(unsigned expected) const { base.BeginExpectationBatch(expected); }
  std::vector<double> OneQubitMoments(const State& s) const {
    g_passes.moment_calls += 1;
    return base.OneQubitMoments(s);
  }
  std::vector<std::complex<double>> EndExpectationBatch(unsigned max_count) const {
    return base.EndExpectationBatch(max_count);
  }
  void RegisterOperatorGroup(const std::vector<const fp_type*>& ms, unsigned nq) const { base.RegisterOperatorGroup(ms, nq); }
  uint64_t OperatorGroupHits() const { return base.OperatorGroupHits(); }
#endif
  Base base;
};

}  // namespace

int main(int argc, char* argv[]) {
  using namespace qsim;
  using fp_type = float;
  const Options opt = Parse(argc, argv);

#if defined(QTRAJ_REFERENCE_IS_CUDA)
  struct Factory {
    using Simulator = Counting<qsim::SimulatorCUDA<fp_type>>;
    using StateSpace = Simulator::StateSpace;
    explicit Factory(unsigned) {}
    StateSpace CreateStateSpace() const { return StateSpace(StateSpace::Parameter{}); }
    Simulator CreateSimulator() const { return Simulator(); }
  };
  Factory factory(0);
#elif defined(QTRAJ_REFERENCE_CPU)
  struct Factory {
    using Simulator = Counting<qsim::Simulator<For>>;
    using StateSpace = Simulator::StateSpace;
    explicit Factory(unsigned t) : threads(t) {}
    StateSpace CreateStateSpace() const { return StateSpace(threads); }
    Simulator CreateSimulator() const { return Simulator(threads); }
    unsigned threads;
  };
  Factory factory(opt.threads ? opt.threads : 1);
#else
  struct Factory {
    using Simulator = Counting<qsim::SimulatorB200<fp_type>>;
    using StateSpace = Simulator::StateSpace;
    StateSpace CreateStateSpace() const { return StateSpace(param); }
    Simulator CreateSimulator() const { return Simulator(); }
    StateSpace::Parameter param;
  };
  Factory factory;  // device: the launcher pins one GPU per process with CUDA_VISIBLE_DEVICES
#endif
  using Simulator = Factory::Simulator;
  using StateSpace = Factory::StateSpace;
  using Fuser = MultiQubitGateFuser<IO>;
#ifdef QTRAJ_REFERENCE_CPU
  using Runner = QSimRunner<IO, Fuser, Factory>;
#else
  using Runner = PrefixSharingRunner<IO, Fuser, Factory>;  // QSimRunner's flow unless a cache is armed
#endif
  using QTSimulator = QuantumTrajectorySimulator<IO, Runner>;

  Circuit<Operation<fp_type>> circuit;
  if (opt.circuit_file.empty() ||
      !CircuitQsimParser<IOFile>::FromFile(opt.maxtime, opt.circuit_file, circuit)) {
    std::fprintf(stderr, "cannot read circuit\n");
    return 1;
  }
  if (opt.channel != "depolarize" && opt.channel != "amplitude_damp") {
    std::fprintf(stderr, "unknown channel %s\n", opt.channel.c_str());
    return 1;
  }
  const auto ncircuit = opt.channel == "depolarize" ? MakeNoisy(circuit, Cirq::DepolarizingChannel<fp_type>(opt.p))
                                                    : MakeNoisy(circuit, Cirq::AmplitudeDampingChannel<fp_type>(opt.p));
  const auto observables = Observables<fp_type>(circuit.num_qubits);
#ifndef QTRAJ_REFERENCE_CPU
  // the observables are the same for every trajectory: reduce their strings to (qubits, matrix) once
  const auto plan = MakeObservablePlan<IO, MultiQubitGateFuser<IO>>(observables, circuit.num_qubits);
#endif

  typename QTSimulator::Parameter param;
  param.max_fused_size = opt.max_fused_size;
  param.verbosity = opt.verbosity;
  param.apply_last_deferred_ops = true;

#ifdef QTRAJ_REFERENCE_CPU
  const unsigned workers = 1;
#else
  const unsigned workers = std::max(1u, std::min(opt.workers, std::max(1u, opt.num)));
#endif
  using Clock = std::chrono::steady_clock;
  std::vector<std::complex<double>> sums(observables.size(), 0.0);
  PassCount passes;
  Clock::time_point first_start = Clock::time_point::max(), last_end = Clock::time_point::min();
  std::mutex merge;
  uint64_t group_hits = 0;   // Kraus probabilities served by another operator's read pass (operator groups, -k 1)
  uint64_t prefix_skipped = 0, prefix_clean = 0;  // fused gates not applied thanks to the shared prefix; noiseless trajectories
  std::atomic<unsigned> ready{0};
  std::atomic<bool> failed{false};

  auto work = [&](unsigned w) {
    // contiguous sub-slice of [traj0, traj0 + num)
    const unsigned base = opt.num / workers, extra = opt.num % workers;
    const unsigned first = opt.traj0 + w * base + std::min(w, extra), count = base + (w < extra ? 1 : 0);
    Simulator simulator = factory.CreateSimulator();
    StateSpace state_space = factory.CreateStateSpace();
#ifndef QTRAJ_REFERENCE_CPU
    if (workers > 1) {  // CUDA's per-thread default stream: the workers' kernels interleave on the GPU
      simulator.base.SetStream(qsim::b200::kStreamPerThread);
      state_space.SetStream(qsim::b200::kStreamPerThread);
    }
#endif
#ifndef QTRAJ_REFERENCE_CPU
    if (opt.kraus_groups) {
      // the K^dagger K of a channel's non-unitary operators form a group: whichever the sampling loop of
      // lib/qtrajectory.h:344-352 asks for first, all of them come out of that one read pass
      for (const auto& op : ncircuit.ops) {
        const auto* ch = OpGetAlternative<Channel<fp_type>>(op);
        if (!ch) continue;
        std::vector<const fp_type*> ms;
        unsigned nq = 0;
        bool same = true;
        for (const auto& kop : ch->kops) {
          if (kop.unitary) continue;
          if (!ms.empty() && kop.qubits.size() != nq) same = false;
          nq = (unsigned) kop.qubits.size();
          ms.push_back(kop.kd_k.data());
        }
        if (same && ms.size() >= 2) simulator.RegisterOperatorGroup(ms, nq);
      }
    }
#endif
    auto state = state_space.Create(circuit.num_qubits);
    bool ok = !state_space.IsNull(state);
    if (!ok) std::fprintf(stderr, "not enough memory\n");
    typename QTSimulator::Stat stat;
    std::vector<std::complex<double>> local(observables.size(), 0.0);
#ifndef QTRAJ_REFERENCE_CPU
    // noiseless prefix shared between this worker's trajectories (only the unitary-mixture channel defers the
    // whole circuit into one flush; amplitude damping flushes at every channel and gains nothing)
    typename Runner::Cache cache;
    const bool share = opt.prefix != 0 && opt.channel == "depolarize" && ok;
    std::vector<std::complex<double>> clean_evals;
    if (share) {
      std::vector<std::variant<const Gate<fp_type>*, const Operation<fp_type>*>> clean_ops;
      for (const auto& op : ncircuit.ops)
        if (!OpGetAlternative<Channel<fp_type>>(op)) clean_ops.push_back(&op);
      ok = cache.template Build<Fuser>(param, circuit.num_qubits, clean_ops, state_space, simulator);
    }
#endif
    if (ok) {
      // one untimed trajectory: context creation, kernel loading, scratch growth
      state_space.SetStateZero(state);
      ok = QTSimulator::RunOnce(param, ncircuit, first, state_space, simulator, state, stat);
      if (ok) (void) ExpectationValue<IO, Fuser>(observables.back(), simulator, state);
    }
    if (!ok) failed = true;
    ++ready;
    while (ready.load() < workers) std::this_thread::yield();  // all workers start their timed loops together
    g_passes = PassCount{};
    const auto t0 = Clock::now();
    for (unsigned i = 0; i < count && !failed.load(); ++i) {
      state_space.SetStateZero(state);
#ifndef QTRAJ_REFERENCE_CPU
      if (share) Runner::Arm(&cache);
#endif
      // seed = repetition id, as QuantumTrajectorySimulator::RunBatch does (lib/qtrajectory.h:268)
      if (!QTSimulator::RunOnce(param, ncircuit, uint64_t{first} + i, state_space, simulator, state, stat)) {
        failed = true;
        break;
      }
#ifndef QTRAJ_REFERENCE_CPU
      if (opt.batch) {
        // a trajectory without any noise event ends in the noiseless state: its observables are computed once
        const bool clean = share && Runner::was_clean();
        if (clean && !clean_evals.empty()) {
          for (std::size_t k = 0; k < observables.size(); ++k) local[k] += clean_evals[k];
          continue;
        }
        const auto evals = ExpectationValues(plan, simulator, state, opt.batch >= 2);
        if (clean) clean_evals = evals;
        for (std::size_t k = 0; k < observables.size(); ++k) local[k] += evals[k];
        continue;
      }
#endif
      for (std::size_t k = 0; k < observables.size(); ++k) {
        local[k] += ExpectationValue<IO, Fuser>(observables[k], simulator, state);
      }
    }
    const auto t1 = Clock::now();
    std::lock_guard<std::mutex> lock(merge);
#ifndef QTRAJ_REFERENCE_CPU
    prefix_skipped += cache.gates_skipped;
    prefix_clean += cache.clean_runs;
    group_hits += simulator.OperatorGroupHits();
#endif
    for (std::size_t k = 0; k < sums.size(); ++k) sums[k] += local[k];
    passes.gates += g_passes.gates;
    passes.expects += g_passes.expects;
    passes.moment_calls += g_passes.moment_calls;
    first_start = std::min(first_start, t0);
    last_end = std::max(last_end, t1);
  };

  if (workers == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    for (unsigned w = 0; w < workers; ++w) pool.emplace_back(work, w);
    for (auto& t : pool) t.join();
  }
  if (failed) return 1;
  const double seconds = std::chrono::duration<double>(last_end - first_start).count();

  std::printf("{\"n\": %u, \"traj0\": %u, \"num\": %u, \"num_ops\": %zu, \"num_observables\": %zu, "
              "\"gate_passes\": %llu, \"expect_passes\": %llu, \"moment_calls\": %llu, \"workers\": %u, "
              "\"prefix_gates_skipped\": %llu, \"noiseless_trajectories\": %llu, \"kraus_group_hits\": %llu, \"channel\": \"%s\", "
              "\"seconds\": %.6f, \"sums\": [",
              circuit.num_qubits, opt.traj0, opt.num, ncircuit.ops.size(), observables.size(),
              (unsigned long long) passes.gates, (unsigned long long) passes.expects,
              (unsigned long long) passes.moment_calls, workers, (unsigned long long) prefix_skipped,
              (unsigned long long) prefix_clean, (unsigned long long) group_hits, opt.channel.c_str(), seconds);
  for (std::size_t k = 0; k < sums.size(); ++k) {
    std::printf("%s%.9g, %.9g", k ? ", " : "", sums[k].real(), sums[k].imag());
  }
  std::printf("]}\n");
  return 0;
}
