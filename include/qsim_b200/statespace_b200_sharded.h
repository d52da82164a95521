// statespace_b200_sharded.h -- StateSpace / VectorSpace for a state sharded over several B200s by global qubits,
// behind the same duck-typed API: the counterpart of lib/vectorspace_custatevecex.h (multi-device State, wire
// ordering, :163-287) and lib/statespace_custatevecex.h:55-420, over the qb200_sv_* C ABI (csrc/sharded.cu).
// The reference's CRTP base lib/statespace.h supplies Measure / VirtualMeasure on top of PartialNorms /
// FindMeasuredBits / Collapse.  Plain C++17; this process drives every shard (single-process multi-device, the
// mode lib/vectorspace_custatevecex.h:385-470 calls kMultiDevice).
#ifndef QSIM_B200_STATESPACE_B200_SHARDED_H_
#define QSIM_B200_STATESPACE_B200_SHARDED_H_

#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "statespace.h"  // reference: lib/statespace.h
#include "util.h"        // reference: lib/util.h (GenerateRandomValues)
#include "vectorspace_b200.h"

namespace qsim {

namespace b200 {

inline void CheckSv(int status, const qb200_sv* sv, const char* file, int line) {
  if (status == QB200_OK || status == QB200_ERR_UNSUPPORTED) return;
  std::fprintf(stderr, "CUDA error: %s at %s %d\n",
               status == QB200_ERR_CUDA ? "sharded state: CUDA runtime failure"
               : status == QB200_ERR_OOM ? "out of memory" : "invalid argument", file, line);
  std::exit(status == QB200_ERR_CUDA && sv ? qb200_sv_last_cuda_error(sv) : status);
}
#define QB200_SV_CHECK(sv, call) ::qsim::b200::CheckSv((call), (sv), __FILE__, __LINE__)

// Devices a sharded state spreads over: an explicit list, or the first 2^k visible devices.
struct ShardedParameter {
  std::vector<int> devices;   // devices[r] hosts shard r; empty = all visible devices (rounded down to 2^k)
  int swap_mode = -1;         // qb200_sv_set_option("swap_mode")
  int reorder = 1;            // the runner may reorder commuting gates to save exchanges
};

inline std::vector<int> DefaultDevices() {
  int count = 0;
  if (qb200_device_count(&count) != QB200_OK || count < 1) {
    std::fprintf(stderr, "CUDA error: no usable CUDA device (qsim_b200 has no CPU fallback) at %s %d\n", __FILE__, __LINE__);
    std::exit(1);
  }
  unsigned p = 1;
  while (2 * p <= (unsigned) count) p *= 2;
  std::vector<int> d(p);
  for (unsigned i = 0; i < p; ++i) d[i] = (int) i;
  return d;
}

}  // namespace b200

template <typename Impl, typename FP>
class VectorSpaceB200Sharded {
 public:
  using fp_type = FP;
  using Parameter = b200::ShardedParameter;

 private:
  struct Deleter { void operator()(qb200_sv* p) const { qb200_sv_destroy(p); } };
  using Pointer = std::unique_ptr<qb200_sv, Deleter>;

 public:
  class Vector {
   public:
    Vector() = delete;
    Vector(Pointer&& ptr, unsigned num_qubits) : ptr_(std::move(ptr)), num_qubits_(num_qubits) {}

    // the opaque sharded state (there is no single device pointer to hand out)
    qb200_sv* get() { return ptr_.get(); }
    qb200_sv* get() const { return ptr_.get(); }
    qb200_sv* release() { num_qubits_ = 0; return ptr_.release(); }
    unsigned num_qubits() const { return num_qubits_; }
    unsigned num_substates() const { return ptr_ ? qb200_sv_num_shards(ptr_.get()) : 0; }
    static constexpr bool requires_copy_to_host() { return true; }

    // wire ordering, lib/vectorspace_custatevecex.h:163-177: physical bit of every logical qubit
    std::vector<unsigned> get_wire_ordering() const {
      std::vector<unsigned> pos(num_qubits_);
      if (ptr_) qb200_sv_qubit_map(ptr_.get(), pos.data());
      return pos;
    }
    void to_normal_order() { if (ptr_) QB200_SV_CHECK(ptr_.get(), qb200_sv_canonicalize(ptr_.get())); }

   private:
    Pointer ptr_;
    unsigned num_qubits_;
  };

  VectorSpaceB200Sharded() : param_() {}
  explicit VectorSpaceB200Sharded(const Parameter& param) : param_(param) {}

  // A gate must fit one shard (lib/simulator_custatevecex.h:67-71) and the Simulator API promises gates of up
  // to 6 qubits, so a state is spread over 2^g shards only while every shard keeps at least 6 local qubits:
  // small states get fewer shards (down to one) instead of being refused
  // (cf. lib/multiprocess_custatevecex.h:160-163).
  Vector Create(unsigned num_qubits) const {
    std::vector<int> dev = param_.devices.empty() ? b200::DefaultDevices() : param_.devices;
    unsigned p = 1, g = 0;
    while (2 * p <= dev.size()) { p *= 2; ++g; }
    while (p > 1 && num_qubits < g + 6) { p /= 2; --g; }
    qb200_sv* sv = nullptr;
    int rc = qb200_sv_create(dev.data(), p, num_qubits, b200::DType<FP>::value, &sv);
    if (rc != QB200_OK) return Null();  // "not enough memory" is the caller's message (lib/run_qsim.h:93-97)
    qb200_sv_set_option(sv, "swap_mode", param_.swap_mode);
    qb200_sv_set_option(sv, "reorder", param_.reorder);
    return Vector{Pointer{sv}, num_qubits};
  }

  // wrapping caller-owned device memory makes no sense for a multi-device state
  Vector Create(fp_type*, unsigned) const { return Null(); }

  static Vector Null() { return Vector{Pointer{nullptr}, 0}; }
  static bool IsNull(const Vector& vector) { return vector.get() == nullptr; }
  static void Free(fp_type* ptr) { std::free(ptr); }

  bool Copy(const Vector& src, Vector& dest) const {
    if (src.num_qubits() != dest.num_qubits()) return false;
    QB200_SV_CHECK(src.get(), qb200_sv_copy(src.get(), dest.get()));
    return true;
  }
  // host buffers hold the whole state in normal order (what pybind hands to Python)
  bool Copy(const Vector& src, fp_type* dest) const {
    QB200_SV_CHECK(src.get(), qb200_sv_copy_to_host(src.get(), dest));
    return true;
  }
  bool Copy(const fp_type* src, Vector& dest) const {
    QB200_SV_CHECK(dest.get(), qb200_sv_copy_from_host(dest.get(), src));
    return true;
  }
  bool Copy(const fp_type* src, uint64_t size, Vector& dest) const {
    if (size < Impl::MinSize(dest.num_qubits())) return false;
    return Copy(src, dest);
  }

  static void DeviceSync() {
    int count = 0;
    qb200_device_count(&count);
    for (int d = 0; d < count; ++d) QB200_CHECK(nullptr, qb200_device_sync_on(d));
  }

 protected:
  Parameter param_;
};

template <typename FP = float>
class StateSpaceB200Sharded : public StateSpace<StateSpaceB200Sharded<FP>, VectorSpaceB200Sharded, FP> {
 private:
  using Base = StateSpace<StateSpaceB200Sharded<FP>, qsim::VectorSpaceB200Sharded, FP>;

 public:
  using State = typename Base::State;
  using fp_type = typename Base::fp_type;
  using MeasurementResult = typename Base::MeasurementResult;
  using Parameter = b200::ShardedParameter;

  StateSpaceB200Sharded() : Base() {}
  explicit StateSpaceB200Sharded(const Parameter& param) : Base(param) {}

  static uint64_t MinSize(unsigned num_qubits) { return qb200_min_size(num_qubits); }

  // lib/statespace_custatevecex.h:79-85
  void InternalToNormalOrder(State& state) const { state.to_normal_order(); }
  void NormalToInternalOrder(State&) const {}

  void SetAllZeros(State& state) const { QB200_SV_CHECK(state.get(), qb200_sv_set_all_zeros(state.get())); }
  void SetStateUniform(State& state) const { QB200_SV_CHECK(state.get(), qb200_sv_set_state_uniform(state.get())); }
  void SetStateZero(State& state) const { QB200_SV_CHECK(state.get(), qb200_sv_set_state_zero(state.get())); }

  // through the qubit map (lib/statespace_custatevecex.h:120-147)
  std::complex<fp_type> GetAmpl(const State& state, uint64_t i) const {
    double out[2];
    QB200_SV_CHECK(state.get(), qb200_sv_get_ampl(state.get(), i, out));
    return std::complex<fp_type>((fp_type) out[0], (fp_type) out[1]);
  }
  void SetAmpl(State& state, uint64_t i, const std::complex<fp_type>& ampl) const {
    SetAmpl(state, i, std::real(ampl), std::imag(ampl));
  }
  void SetAmpl(State& state, uint64_t i, fp_type re, fp_type im) const {
    QB200_SV_CHECK(state.get(), qb200_sv_set_ampl(state.get(), i, re, im));
  }

  // implemented here (the reference leaves it empty, lib/statespace_custatevecex.h:195-208)
  void BulkSetAmpl(State& state, uint64_t mask, uint64_t bits, const std::complex<fp_type>& val,
                   bool exclude = false) const {
    BulkSetAmpl(state, mask, bits, std::real(val), std::imag(val), exclude);
  }
  void BulkSetAmpl(State& state, uint64_t mask, uint64_t bits, fp_type re, fp_type im, bool exclude = false) const {
    QB200_SV_CHECK(state.get(), qb200_sv_bulk_set_ampl(state.get(), mask, bits, re, im, exclude));
  }

  bool Add(const State& src, State& dest) const {
    if (src.num_qubits() != dest.num_qubits()) return false;
    QB200_SV_CHECK(dest.get(), qb200_sv_add(src.get(), dest.get()));
    return true;
  }
  void Multiply(fp_type a, State& state) const { QB200_SV_CHECK(state.get(), qb200_sv_multiply(state.get(), a)); }

  std::complex<double> InnerProduct(const State& state1, const State& state2) const {
    if (state1.num_qubits() != state2.num_qubits()) return std::nan("");
    double out[2];
    QB200_SV_CHECK(state1.get(), qb200_sv_inner_product(state1.get(), state2.get(), out));
    return {out[0], out[1]};
  }
  double RealInnerProduct(const State& state1, const State& state2) const {
    return std::real(InnerProduct(state1, state2));
  }
  double Norm(const State& state) const {
    double out;
    QB200_SV_CHECK(state.get(), qb200_sv_norm(state.get(), &out));
    return out;
  }

  template <typename DistrRealType = double>
  std::vector<uint64_t> Sample(const State& state, uint64_t num_samples, unsigned seed) const {
    std::vector<uint64_t> bitstrings;
    if (num_samples > 0) {
      double norm = Norm(state);
      auto rs = GenerateRandomValues<DistrRealType>(num_samples, seed, norm);  // lib/statespace_cuda.h:293
      std::vector<double> rsd(rs.begin(), rs.begin() + num_samples);
      bitstrings.resize(num_samples, 0);
      QB200_SV_CHECK(state.get(), qb200_sv_sample(state.get(), rsd.data(), num_samples, bitstrings.data()));
    }
    return bitstrings;
  }

  void Collapse(const MeasurementResult& mr, State& state) const {
    QB200_SV_CHECK(state.get(), qb200_sv_collapse(state.get(), mr.mask, mr.bits, nullptr));
  }
  std::vector<double> PartialNorms(const State& state) const {
    std::vector<double> norms(qb200_sv_partial_norms_count(state.get()));
    QB200_SV_CHECK(state.get(), qb200_sv_partial_norms(state.get(), norms.data()));
    return norms;
  }
  uint64_t FindMeasuredBits(unsigned m, double r, uint64_t mask, const State& state) const {
    uint64_t bits = 0;
    QB200_SV_CHECK(state.get(), qb200_sv_find_measured_bits(state.get(), m, r, mask, &bits));
    return bits;
  }
};

}  // namespace qsim

#endif  // QSIM_B200_STATESPACE_B200_SHARDED_H_
