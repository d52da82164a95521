// qtrajectory_b200.h -- prefix sharing for Monte-Carlo trajectories (the device-side
// trajectory loop).  A runner for lib/qtrajectory.h's QuantumTrajectorySimulator<IO, Runner> that keeps the
// reference's host logic (Kraus sampling in RunOnce, MultiQubitGateFuser per flush, lib/qtrajectory.h:270-411)
// and changes only what reaches the GPU:
//
// At weak noise most sampled Kraus operators are the identity, so the fused gate list of a trajectory is the
// noiseless list except from the first fused gate that absorbed a sampled Pauli.  PrefixCache applies the
// noiseless fused list ONCE, keeping a copy of the state after every fused gate (26 qubits: 36 x 512 MiB of the
// 180 GB); a trajectory whose fused list agrees with the noiseless one on its first j gates (same qubits, same
// matrix bits) starts from checkpoint j and applies only the rest.  The arithmetic a trajectory's state goes
// through is the one the plain runner would do, gate for gate, so the results are bit-identical
// (tests/test_traj_farm.py); a trajectory with no event at all (37 % at 26 qubits, p = 0.001) costs one copy.
#ifndef QSIM_B200_QTRAJECTORY_B200_H_
#define QSIM_B200_QTRAJECTORY_B200_H_

#include <cstdint>
#include <cstring>
#include <random>
#include <variant>
#include <vector>

#include "circuit.h"         // reference
#include "gate.h"            // reference
#include "gate_appl.h"       // reference
#include "operation_base.h"  // reference
#include "util.h"            // reference

namespace qsim {

template <typename Factory>
class PrefixCache {
 public:
  using Simulator = typename Factory::Simulator;
  using StateSpace = typename Simulator::StateSpace;
  using State = typename StateSpace::State;
  using fp_type = typename Simulator::fp_type;

  struct Entry {
    std::vector<unsigned> qubits;
    std::vector<fp_type> matrix;
  };

  // Fuses `clean_ops` (the ops a trajectory without any noise event defers: every non-channel op of the noisy
  // circuit), applies them to |0...0> and keeps a checkpoint after each fused gate while memory lasts.
  template <typename Fuser, typename FuserParameter, typename Ops>
  bool Build(const FuserParameter& param, unsigned num_qubits, const Ops& clean_ops,
             const StateSpace& state_space, const Simulator& simulator, std::size_t max_checkpoints = ~std::size_t{0}) {
    entries_.clear();
    checkpoints_.clear();
    clean_ptrs_.clear();
    for (const auto& op : clean_ops) clean_ptrs_.push_back(Address(op));
    auto fused = Fuser::FuseGates(param, num_qubits, clean_ops);
    if (fused.size() == 0 && clean_ops.size() > 0) return false;
    auto state = state_space.Create(num_qubits);
    if (state_space.IsNull(state)) return false;
    state_space.SetStateZero(state);
    bool keep = true;
    for (std::size_t j = 0; j < fused.size(); ++j) {
      const auto* pg = OpGetAlternative<FusedGate<fp_type>>(fused[j]);
      if (pg == nullptr) break;  // measurement or controlled gate: the shared prefix ends here
      entries_.push_back({pg->qubits, std::vector<fp_type>(pg->matrix.begin(), pg->matrix.end())});
      simulator.ApplyGate(pg->qubits, pg->matrix.data(), state);
      if (keep && checkpoints_.size() < max_checkpoints) {
        auto cp = state_space.Create(num_qubits);
        if (state_space.IsNull(cp)) {
          keep = false;  // out of memory: later gates are not checkpointed
        } else {
          state_space.Copy(state, cp);
          checkpoints_.push_back(std::move(cp));
        }
      }
    }
    total_ = fused.size();
    return true;
  }

  std::size_t num_checkpoints() const { return checkpoints_.size(); }
  std::size_t num_gates() const { return total_; }

  // Longest usable prefix of `fused`: the number of leading fused gates identical to the noiseless list,
  // limited to the checkpoints that exist.
  template <typename FusedOps>
  std::size_t SharedPrefix(const FusedOps& fused) const {
    std::size_t j = 0;
    const std::size_t limit = std::min(fused.size(), std::min(entries_.size(), checkpoints_.size()));
    for (; j < limit; ++j) {
      const auto* pg = OpGetAlternative<FusedGate<fp_type>>(fused[j]);
      if (pg == nullptr || pg->qubits != entries_[j].qubits || pg->matrix.size() != entries_[j].matrix.size() ||
          std::memcmp(pg->matrix.data(), entries_[j].matrix.data(), pg->matrix.size() * sizeof(fp_type)) != 0) {
        break;
      }
    }
    return j;
  }

  // state after the first j (>= 1) noiseless fused gates
  const State& Checkpoint(std::size_t j) const { return checkpoints_[j - 1]; }

  // True when `ops` (what a trajectory deferred) IS the noiseless list, operation for operation: every sampled Kraus
  // operator was one without gates (the identity of a depolarizing channel defers nothing, lib/channels_cirq.h:137),
  // so the list holds the very pointers Build() saw.  Decided without fusing (1 ms of host time per trajectory at
  // 26 qubits, depth 20); needs the checkpoint behind the last fused gate.
  template <typename Ops>
  bool IsCleanList(const Ops& ops) const {
    if (total_ == 0 || checkpoints_.size() != total_ || entries_.size() != total_) return false;
    if (ops.size() != clean_ptrs_.size()) return false;
    for (std::size_t i = 0; i < ops.size(); ++i)
      if (Address(ops[i]) != clean_ptrs_[i]) return false;
    return true;
  }

  // statistics of the runs served so far
  uint64_t runs = 0, gates_skipped = 0, gates_applied = 0, clean_runs = 0;

 private:
  template <typename... Ts>
  static const void* Address(const std::variant<Ts...>& v) {
    return std::visit([](auto p) -> const void* { return p; }, v);
  }
  template <typename T>
  static const void* Address(const T* p) { return p; }
  template <typename T>
  static const void* Address(const T& op) { return &op; }

  std::vector<Entry> entries_;
  std::vector<State> checkpoints_;
  std::vector<const void*> clean_ptrs_;
  std::size_t total_ = 0;
};

/**
 * Runner with QSimRunner's `Run(param, ops, state_space, simulator, state)` entry (lib/run_qsim.h:305-315, the
 * one lib/qtrajectory.h:405 calls).  When a cache is armed for the calling thread (Arm(), once per trajectory,
 * right after SetStateZero) the first flush of the trajectory starts from the deepest matching checkpoint.
 */
template <typename IO, typename Fuser, typename Factory, typename RGen = std::mt19937>
struct PrefixSharingRunner final {
  using Simulator = typename Factory::Simulator;
  using StateSpace = typename Simulator::StateSpace;
  using State = typename StateSpace::State;
  using MeasurementResult = typename StateSpace::MeasurementResult;
  using fp_type = typename Simulator::fp_type;
  using Cache = PrefixCache<Factory>;

  struct Parameter : public Fuser::Parameter {
    uint64_t seed;
  };

  // The next Run of this thread may use `cache` (the state passed to it must be |0...0>); was_clean() tells
  // afterwards whether the trajectory turned out to be the noiseless one.
  static void Arm(Cache* cache) { armed() = cache; clean() = false; }
  static bool was_clean() { return clean(); }

  template <typename Circuit>
  static bool Run(const Parameter& param, const Circuit& circuit, const StateSpace& state_space,
                  const Simulator& simulator, State& state) {
    std::vector<MeasurementResult> discarded;
    return Run(param, circuit, state_space, simulator, state, discarded);
  }

  template <typename Circuit>
  static bool Run(const Parameter& param, const Circuit& circuit, const StateSpace& state_space,
                  const Simulator& simulator, State& state, std::vector<MeasurementResult>& measure_results) {
    RGen rgen(param.seed);
    const auto& ops = Operations<Circuit>::get(circuit);
    if (armed() != nullptr && armed()->IsCleanList(ops)) {
      // the noiseless trajectory, recognised before fusing: one copy, no fuser, no gate pass
      Cache* cache = armed();
      armed() = nullptr;
      RestoreCheckpoint(state_space, cache->Checkpoint(cache->num_gates()), state, 0);
      ++cache->runs;
      cache->gates_skipped += cache->num_gates();
      clean() = true;
      ++cache->clean_runs;
      return true;
    }
    auto fused_ops = Fuser::FuseGates(param, state.num_qubits(), ops);
    if (fused_ops.size() == 0 && ops.size() > 0) return false;
    measure_results.reserve(fused_ops.size());

    std::size_t start = 0;
    Cache* cache = armed();
    armed() = nullptr;  // one flush per arming
    if (cache != nullptr) {
      start = cache->SharedPrefix(fused_ops);
      if (start > 0) RestoreCheckpoint(state_space, cache->Checkpoint(start), state, 0);
      ++cache->runs;
      cache->gates_skipped += start;
      cache->gates_applied += fused_ops.size() - start;
      if (start == fused_ops.size() && start == cache->num_gates()) {
        clean() = true;
        ++cache->clean_runs;
      }
    }
    for (std::size_t i = start; i < fused_ops.size(); ++i) {
      if (!ApplyGate(state_space, simulator, fused_ops[i], rgen, state, measure_results)) {
        IO::errorf("measurement failed.\n");
        return false;
      }
    }
    return true;
  }

 private:
  // enqueued on the state space's stream when the backend can (StateSpaceB200::CopyAsync): the gates that follow
  // are ordered after it on the same stream, the host goes on
  template <typename SS>
  static auto RestoreCheckpoint(const SS& ss, const State& from, State& to, int) -> decltype(ss.CopyAsync(from, to), void()) {
    ss.CopyAsync(from, to);
  }
  template <typename SS>
  static void RestoreCheckpoint(const SS& ss, const State& from, State& to, long) {
    ss.Copy(from, to);
  }

  static Cache*& armed() {
    thread_local Cache* c = nullptr;
    return c;
  }
  static bool& clean() {
    thread_local bool c = false;
    return c;
  }
};

}  // namespace qsim

#endif  // QSIM_B200_QTRAJECTORY_B200_H_
