// statespace_b200.h -- drop-in for qsim's StateSpaceCUDA (lib/statespace_cuda.h:43-468)
// on top of the reference's own CRTP base lib/statespace.h (Measure / VirtualMeasure /
// Norm stay the reference's code).  State layout is normal order (see qsim_b200.h).
#ifndef QSIM_B200_STATESPACE_B200_H_
#define QSIM_B200_STATESPACE_B200_H_

#include <cmath>
#include <complex>
#include <cstdint>
#include <type_traits>
#include <vector>

#include "statespace.h"  // reference: lib/statespace.h
#include "util.h"        // reference: lib/util.h (GenerateRandomValues)
#include "vectorspace_b200.h"

namespace qsim {

template <typename FP = float>
class StateSpaceB200 : public StateSpace<StateSpaceB200<FP>, VectorSpaceB200, FP> {
 private:
  using Base = StateSpace<StateSpaceB200<FP>, qsim::VectorSpaceB200, FP>;
  static constexpr int kDT = b200::DType<FP>::value;

 public:
  using State = typename Base::State;
  using fp_type = typename Base::fp_type;
  using MeasurementResult = typename Base::MeasurementResult;

  // Kept so code written for StateSpaceCUDA::Parameter (lib/statespace_cuda.h:59-70)
  // compiles unchanged; launch shapes are chosen by the library.
  struct Parameter {
    unsigned num_threads = 512;
    unsigned num_dblocks = 16;
  };

  StateSpaceB200() : Base() {}
  explicit StateSpaceB200(const Parameter&) : Base() {}

  static uint64_t MinSize(unsigned num_qubits) { return qb200_min_size(num_qubits); }

  void InternalToNormalOrder(State& state) const {
    QB200_CHECK(this->ctx(), qb200_internal_to_normal_order(this->ctx(), kDT, state.get(), state.num_qubits()));
  }
  void NormalToInternalOrder(State& state) const {
    QB200_CHECK(this->ctx(), qb200_normal_to_internal_order(this->ctx(), kDT, state.get(), state.num_qubits()));
  }

  void SetAllZeros(State& state) const {
    QB200_CHECK(this->ctx(), qb200_set_all_zeros(this->ctx(), kDT, state.get(), state.num_qubits()));
  }
  void SetStateUniform(State& state) const {
    QB200_CHECK(this->ctx(), qb200_set_state_uniform(this->ctx(), kDT, state.get(), state.num_qubits()));
  }
  void SetStateZero(State& state) const {
    QB200_CHECK(this->ctx(), qb200_set_state_zero(this->ctx(), kDT, state.get(), state.num_qubits()));
  }

  static std::complex<fp_type> GetAmpl(const State& state, uint64_t i) {
    double out[2];
    qb200_ctx* c = b200::ThreadContext();
    QB200_CHECK(c, qb200_get_ampl(c, kDT, state.get(), i, out));
    return std::complex<fp_type>((fp_type) out[0], (fp_type) out[1]);
  }
  static void SetAmpl(State& state, uint64_t i, const std::complex<fp_type>& ampl) {
    SetAmpl(state, i, std::real(ampl), std::imag(ampl));
  }
  static void SetAmpl(State& state, uint64_t i, fp_type re, fp_type im) {
    qb200_ctx* c = b200::ThreadContext();
    QB200_CHECK(c, qb200_set_ampl(c, kDT, state.get(), i, re, im));
  }

  void BulkSetAmpl(State& state, uint64_t mask, uint64_t bits, const std::complex<fp_type>& val,
                   bool exclude = false) const {
    BulkSetAmpl(state, mask, bits, std::real(val), std::imag(val), exclude);
  }
  void BulkSetAmpl(State& state, uint64_t mask, uint64_t bits, fp_type re, fp_type im,
                   bool exclude = false) const {
    QB200_CHECK(this->ctx(), qb200_bulk_set_ampl(this->ctx(), kDT, state.get(), state.num_qubits(), mask, bits,
                                                 re, im, exclude));
  }

  bool Add(const State& src, State& dest) const {
    if (src.num_qubits() != dest.num_qubits()) return false;
    QB200_CHECK(this->ctx(), qb200_add(this->ctx(), kDT, src.get(), dest.get(), src.num_qubits()));
    return true;
  }

  void Multiply(fp_type a, State& state) const {
    QB200_CHECK(this->ctx(), qb200_multiply(this->ctx(), kDT, a, state.get(), state.num_qubits()));
  }

  std::complex<double> InnerProduct(const State& state1, const State& state2) const {
    if (state1.num_qubits() != state2.num_qubits()) return std::nan("");
    double out[2];
    QB200_CHECK(this->ctx(), qb200_inner_product(this->ctx(), kDT, state1.get(), state2.get(),
                                                 state1.num_qubits(), out));
    return {out[0], out[1]};
  }

  double RealInnerProduct(const State& state1, const State& state2) const {
    if (state1.num_qubits() != state2.num_qubits()) return std::nan("");
    double out;
    QB200_CHECK(this->ctx(), qb200_real_inner_product(this->ctx(), kDT, state1.get(), state2.get(),
                                                      state1.num_qubits(), &out));
    return out;
  }

  double Norm(const State& state) const {
    double out;
    QB200_CHECK(this->ctx(), qb200_norm(this->ctx(), kDT, state.get(), state.num_qubits(), &out));
    return out;
  }

  template <typename DistrRealType = double>
  std::vector<uint64_t> Sample(const State& state, uint64_t num_samples, unsigned seed) const {
    std::vector<uint64_t> bitstrings;
    if (num_samples > 0) {
      bitstrings.resize(num_samples, 0);
      if (std::is_same<DistrRealType, double>::value) {
        // the reference's TODO (lib/statespace_cuda.h:292): the values GenerateRandomValues<double> would draw on
        // the host are drawn and sorted on the device, bit for bit (csrc/sample_rng.cu); norm = -1: the upper bound
        // is the total of the sampler's own partial sums (Norm(state) without a third pass over the state)
        QB200_CHECK(this->ctx(), qb200_sample_seeded(this->ctx(), kDT, state.get(), state.num_qubits(), num_samples,
                                                     seed, -1.0, bitstrings.data()));
      } else {
        double norm = Norm(state);
        // any other distribution type: host RNG exactly as the reference draws it (lib/statespace_cuda.h:293)
        auto rs = GenerateRandomValues<DistrRealType>(num_samples, seed, norm);
        std::vector<double> rsd(rs.begin(), rs.begin() + num_samples);
        QB200_CHECK(this->ctx(), qb200_sample(this->ctx(), kDT, state.get(), state.num_qubits(), rsd.data(),
                                              num_samples, bitstrings.data()));
      }
    }
    return bitstrings;
  }

  void Collapse(const MeasurementResult& mr, State& state) const {
    QB200_CHECK(this->ctx(), qb200_collapse(this->ctx(), kDT, state.get(), state.num_qubits(), mr.mask, mr.bits,
                                            nullptr));
  }

  std::vector<double> PartialNorms(const State& state) const {
    std::vector<double> norms(qb200_partial_norms_count(state.num_qubits()));
    QB200_CHECK(this->ctx(), qb200_partial_norms(this->ctx(), kDT, state.get(), state.num_qubits(), norms.data()));
    return norms;
  }

  uint64_t FindMeasuredBits(unsigned m, double r, uint64_t mask, const State& state) const {
    uint64_t bits = 0;
    QB200_CHECK(this->ctx(), qb200_find_measured_bits(this->ctx(), kDT, state.get(), state.num_qubits(), m, r, mask,
                                                      &bits));
    return bits;
  }
};

}  // namespace qsim

#endif  // QSIM_B200_STATESPACE_B200_H_
