// simulator_b200.h -- drop-in for qsim's SimulatorCUDA (lib/simulator_cuda.h:35-919):
// same member set, so it compiles behind lib/run_qsim.h, lib/qtrajectory.h,
// lib/expect.h, lib/hybrid.h, the pybind layer and every tests/*_testfixture.h.
#ifndef QSIM_B200_SIMULATOR_B200_H_
#define QSIM_B200_SIMULATOR_B200_H_

#include <complex>
#include <cstdint>
#include <map>
#include <memory>
#include <vector>

#include "statespace_b200.h"

namespace qsim {

template <typename FP = float>
class SimulatorB200 final {
 public:
  using StateSpace = StateSpaceB200<FP>;
  using State = typename StateSpace::State;
  using fp_type = typename StateSpace::fp_type;

  SimulatorB200() : ctx_(b200::MakeContext()) {}

  // lib/simulator_cuda.h:70-125; qs sorted ascending, up to 6 qubits (more: ignored).
  void ApplyGate(const std::vector<unsigned>& qs, const fp_type* matrix, State& state) const {
    QB200_CHECK(ctx(), qb200_apply_gate(ctx(), b200::DType<FP>::value, state.get(), state.num_qubits(),
                                        qs.data(), (unsigned) qs.size(), matrix));
  }

  // lib/simulator_cuda.h:135-207
  void ApplyControlledGate(const std::vector<unsigned>& qs, const std::vector<unsigned>& cqs,
                           uint64_t cvals, const fp_type* matrix, State& state) const {
    QB200_CHECK(ctx(), qb200_apply_controlled_gate(ctx(), b200::DType<FP>::value, state.get(),
                                                   state.num_qubits(), qs.data(), (unsigned) qs.size(),
                                                   cqs.data(), (unsigned) cqs.size(), cvals, matrix));
  }

  // lib/simulator_cuda.h:216-260
  std::complex<double> ExpectationValue(const std::vector<unsigned>& qs, const fp_type* matrix,
                                        const State& state) const {
    if (groups_ && !in_batch_ && !groups_->member.empty()) {   // (a batch pairs results with calls by position)
      std::complex<double> v;
      if (GroupValue(qs, matrix, state, &v)) return v;
    }
    double out[2] = {0, 0};
    QB200_CHECK(ctx(), qb200_expectation_value(ctx(), b200::DType<FP>::value, state.get(), state.num_qubits(),
                                               qs.data(), (unsigned) qs.size(), matrix, out));
    return {out[0], out[1]};
  }

  // Batched expectation values (no reference counterpart; used by expect_b200.h): between Begin and End
  // ExpectationValue enqueues its pass and returns NaN at once; End synchronises the stream once and
  // returns the values in call order.
  void BeginExpectationBatch(unsigned expected = 0) const {
    QB200_CHECK(ctx(), qb200_reduce_batch_begin(ctx(), expected));
    in_batch_ = true;
  }
  std::vector<std::complex<double>> EndExpectationBatch(unsigned max_count) const {
    std::vector<double> buf(2 * std::size_t{max_count} + 2);
    uint32_t count = 0;
    in_batch_ = false;
    QB200_CHECK(ctx(), qb200_reduce_batch_end(ctx(), buf.data(), max_count, &count));
    std::vector<std::complex<double>> out(count);
    for (uint32_t i = 0; i < count; ++i) out[i] = {buf[2 * i], buf[2 * i + 1]};
    return out;
  }

  // Reduced density matrices of all qubits from 3-4 read-only passes (qb200_one_qubit_moments): 4 doubles per
  // qubit, S00, S11, Re S01, Im S01.  Empty when the state cannot take the path (unaligned wrapped memory).
  std::vector<double> OneQubitMoments(const State& state) const {
    std::vector<double> out(4 * std::size_t{state.num_qubits()} + 4);
    int rc = qb200_one_qubit_moments(ctx(), b200::DType<FP>::value, state.get(), state.num_qubits(), out.data());
    if (rc == QB200_ERR_UNSUPPORTED) return {};
    QB200_CHECK(ctx(), rc);
    out.resize(4 * std::size_t{state.num_qubits()});
    return out;
  }

  // Operator groups (no reference counterpart).  The Kraus-operator sampling of a non-unitary channel asks for
  // <psi|K_i^dagger K_i|psi> one operator at a time (lib/qtrajectory.h:344-352) -- same qubits, same state, one
  // read pass and one host synchronisation each.  After RegisterOperatorGroup({kd_k(0).data(), kd_k(1).data(), ...})
  // (host pointers that stay valid and keep their contents, 2..8 matrices on one or two qubits), ExpectationValue
  // called with ANY member evaluates ALL members in one pass (qb200_expectation_values_multi) and serves the others
  // from that result for as long as nothing has written a state (qb200_mutation_epoch) and state pointer and
  // qubits are the same.  Copies of a simulator share the groups.
  void RegisterOperatorGroup(const std::vector<const fp_type*>& matrices, unsigned num_qubits) const {
    if (matrices.size() < 2 || matrices.size() > 8 || num_qubits < 1 || num_qubits > 2) return;
    if (!groups_) groups_ = std::make_shared<Groups>();
    const std::size_t id = groups_->all.size();
    groups_->all.push_back({matrices, num_qubits});
    for (std::size_t i = 0; i < matrices.size(); ++i) groups_->member[matrices[i]] = {id, i};
  }
  void ClearOperatorGroups() const { groups_.reset(); }
  // read passes saved so far: group members answered from a pass that another member paid for
  uint64_t OperatorGroupHits() const { return groups_ ? groups_->hits : 0; }

  // lib/simulator_cuda.h:265-267 (the reference's tests size their sweeps from it)
  static unsigned SIMDRegisterSize() { return 32; }

  void SetStream(void* cuda_stream) const { QB200_CHECK(ctx(), qb200_ctx_set_stream(ctx(), cuda_stream)); }

 private:
  struct Group { std::vector<const fp_type*> matrices; unsigned num_qubits; };
  struct Groups {
    std::vector<Group> all;
    std::map<const fp_type*, std::pair<std::size_t, std::size_t>> member;  // matrix -> (group, index)
    // the last evaluation
    std::size_t group = ~std::size_t{0};
    const void* state = nullptr;
    std::vector<unsigned> qubits;
    uint64_t epoch = 0;
    std::vector<double> values;
    uint64_t hits = 0;
  };

  bool GroupValue(const std::vector<unsigned>& qs, const fp_type* matrix, const State& state,
                  std::complex<double>* v) const {
    const auto it = groups_->member.find(matrix);
    if (it == groups_->member.end()) return false;
    Groups& g = *groups_;
    const Group& grp = g.all[it->second.first];
    if (grp.num_qubits != qs.size()) return false;
    const std::size_t i = it->second.second;
    if (g.group == it->second.first && g.state == state.get() && g.qubits == qs && g.epoch == qb200_mutation_epoch()) {
      ++g.hits;
    } else {
      const std::size_t per = std::size_t{2} << (2 * qs.size());
      std::vector<fp_type> packed(per * grp.matrices.size());
      for (std::size_t k = 0; k < grp.matrices.size(); ++k)
        for (std::size_t j = 0; j < per; ++j) packed[k * per + j] = grp.matrices[k][j];
      g.values.assign(2 * grp.matrices.size(), 0.0);
      int rc = qb200_expectation_values_multi(ctx(), b200::DType<FP>::value, state.get(), state.num_qubits(), qs.data(),
                                              (unsigned) qs.size(), packed.data(), (unsigned) grp.matrices.size(),
                                              g.values.data());
      g.group = ~std::size_t{0};
      if (rc == QB200_ERR_UNSUPPORTED) return false;
      QB200_CHECK(ctx(), rc);
      g.group = it->second.first;
      g.state = state.get();
      g.qubits = qs;
      g.epoch = qb200_mutation_epoch();
    }
    *v = {g.values[2 * i], g.values[2 * i + 1]};
    return true;
  }

  qb200_ctx* ctx() const { return ctx_.get(); }
  std::shared_ptr<qb200_ctx> ctx_;
  mutable std::shared_ptr<Groups> groups_;
  mutable bool in_batch_ = false;
};

}  // namespace qsim

#endif  // QSIM_B200_SIMULATOR_B200_H_
