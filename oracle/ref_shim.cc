// ref_shim.cc -- TEST INFRASTRUCTURE.  Thin C-ABI wrapper around the UNMODIFIED
// reference CPU backends, compiled from the headers where they lie
// (-I$QSIM_REF/lib); the outputs go to oracle/_ref/ only.  Nothing from the
// reference is copied into this repository.
//
// Three engines are exposed behind one handle type:
//   kind 0: SimulatorBasic<For,float>  / StateSpaceBasic<For,float>
//   kind 1: SimulatorBasic<For,double> / StateSpaceBasic<For,double>
//   kind 2: qsim::Simulator<For> from lib/simmux.h (AVX512 / AVX / SSE as the
//           -m flags of this build select) -- the reference's fast CPU path,
//           used as the CPU baseline.
// All state I/O through this shim is in NORMAL order (interleaved re,im); the
// SIMD engine converts with its own NormalToInternalOrder/InternalToNormalOrder.
#include <complex>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>

#include "formux.h"
#include "simmux.h"
#include "simulator_basic.h"
#include "statespace_basic.h"
#include "util_cpu.h"

namespace {

using namespace qsim;

struct Engine {
  virtual ~Engine() {}
  virtual unsigned n() const = 0;
  virtual bool ok() const = 0;
  virtual void set_zero() = 0;
  virtual void set_uniform() = 0;
  virtual void set_all_zeros() = 0;
  virtual void from_normal(const void* src) = 0;
  virtual void to_normal(void* dst) = 0;
  virtual void apply(const std::vector<unsigned>& qs, const void* m) = 0;
  virtual void apply_c(const std::vector<unsigned>& qs,
                       const std::vector<unsigned>& cqs, uint64_t cvals,
                       const void* m) = 0;
  virtual std::complex<double> expect(const std::vector<unsigned>& qs,
                                      const void* m) = 0;
  virtual double norm() = 0;
  virtual std::complex<double> inner(Engine& o) = 0;
  virtual double real_inner(Engine& o) = 0;
  virtual void multiply(double a) = 0;
  virtual bool add_from(Engine& src) = 0;
  virtual void sample(uint64_t num, unsigned seed, uint64_t* out) = 0;
  virtual int measure(const std::vector<unsigned>& qs, unsigned seed,
                      uint64_t* mask, uint64_t* bits) = 0;
  virtual void collapse(uint64_t mask, uint64_t bits) = 0;
  virtual void get_ampl(uint64_t i, double out[2]) = 0;
  virtual void set_ampl(uint64_t i, double re, double im) = 0;
  virtual void bulk_set(uint64_t mask, uint64_t bits, double re, double im,
                        bool exclude) = 0;
  virtual void* raw() = 0;
};

template <typename Simulator>
struct EngineT final : Engine {
  using StateSpace = typename Simulator::StateSpace;
  using State = typename StateSpace::State;
  using fp = typename Simulator::fp_type;

  EngineT(unsigned nq, unsigned threads)
      : ss(threads), sim(threads), st(ss.Create(nq)) {}

  StateSpace ss;
  Simulator sim;
  State st;

  unsigned n() const override { return st.num_qubits(); }
  bool ok() const override { return !StateSpace::IsNull(st); }
  void set_zero() override { ss.SetStateZero(st); }
  void set_uniform() override { ss.SetStateUniform(st); }
  void set_all_zeros() override { ss.SetAllZeros(st); }
  void from_normal(const void* src) override {
    ss.SetAllZeros(st);  // SIMD backends pad small states; keep the padding clean
    std::memcpy(st.get(), src, sizeof(fp) * (uint64_t{2} << n()));
    ss.NormalToInternalOrder(st);
  }
  void to_normal(void* dst) override {
    ss.InternalToNormalOrder(st);
    std::memcpy(dst, st.get(), sizeof(fp) * (uint64_t{2} << n()));
    ss.NormalToInternalOrder(st);
  }
  void apply(const std::vector<unsigned>& qs, const void* m) override {
    sim.ApplyGate(qs, (const fp*) m, st);
  }
  void apply_c(const std::vector<unsigned>& qs,
               const std::vector<unsigned>& cqs, uint64_t cvals,
               const void* m) override {
    sim.ApplyControlledGate(qs, cqs, cvals, (const fp*) m, st);
  }
  std::complex<double> expect(const std::vector<unsigned>& qs,
                              const void* m) override {
    return sim.ExpectationValue(qs, (const fp*) m, st);
  }
  double norm() override { return ss.Norm(st); }
  std::complex<double> inner(Engine& o) override {
    return ss.InnerProduct(st, static_cast<EngineT&>(o).st);
  }
  double real_inner(Engine& o) override {
    return ss.RealInnerProduct(st, static_cast<EngineT&>(o).st);
  }
  void multiply(double a) override { ss.Multiply((fp) a, st); }
  bool add_from(Engine& src) override {
    return ss.Add(static_cast<EngineT&>(src).st, st);
  }
  void sample(uint64_t num, unsigned seed, uint64_t* out) override {
    auto v = ss.Sample(st, num, seed);
    std::memcpy(out, v.data(), v.size() * sizeof(uint64_t));
  }
  int measure(const std::vector<unsigned>& qs, unsigned seed, uint64_t* mask,
              uint64_t* bits) override {
    std::mt19937 rgen(seed);
    auto r = ss.Measure(qs, rgen, st);
    if (!r.valid) return 1;
    *mask = r.mask;
    *bits = r.bits;
    return 0;
  }
  void collapse(uint64_t mask, uint64_t bits) override {
    typename StateSpace::MeasurementResult mr;
    mr.mask = mask;
    mr.bits = bits;
    mr.valid = true;
    ss.Collapse(mr, st);
  }
  void get_ampl(uint64_t i, double out[2]) override {
    auto a = ss.GetAmpl(st, i);
    out[0] = std::real(a);
    out[1] = std::imag(a);
  }
  void set_ampl(uint64_t i, double re, double im) override {
    ss.SetAmpl(st, i, (fp) re, (fp) im);
  }
  void bulk_set(uint64_t mask, uint64_t bits, double re, double im,
                bool exclude) override {
    ss.BulkSetAmpl(st, mask, bits, (fp) re, (fp) im, exclude);
  }
  void* raw() override { return st.get(); }
};

std::vector<unsigned> vec(const unsigned* p, unsigned n) {
  return std::vector<unsigned>(p, p + n);
}

}  // namespace

extern "C" {

const char* ref_simd_name() {
#if defined(__AVX512F__)
  return "SimulatorAVX512";
#elif defined(__AVX2__)
  return "SimulatorAVX";
#elif defined(__SSE4_1__)
  return "SimulatorSSE";
#else
  return "SimulatorBasic";
#endif
}

void* ref_create(int kind, unsigned num_qubits, unsigned threads) {
  Engine* e = nullptr;
  switch (kind) {
    case 0: e = new EngineT<SimulatorBasic<For, float>>(num_qubits, threads); break;
    case 1: e = new EngineT<SimulatorBasic<For, double>>(num_qubits, threads); break;
    case 2: e = new EngineT<qsim::Simulator<For>>(num_qubits, threads); break;
    default: return nullptr;
  }
  if (!e->ok()) { delete e; return nullptr; }
  return e;
}
void ref_destroy(void* h) { delete (Engine*) h; }
void ref_set_zero(void* h) { ((Engine*) h)->set_zero(); }
void ref_set_uniform(void* h) { ((Engine*) h)->set_uniform(); }
void ref_set_all_zeros(void* h) { ((Engine*) h)->set_all_zeros(); }
void ref_from_normal(void* h, const void* src) { ((Engine*) h)->from_normal(src); }
void ref_to_normal(void* h, void* dst) { ((Engine*) h)->to_normal(dst); }
void ref_apply_gate(void* h, const unsigned* qs, unsigned nq, const void* m) {
  ((Engine*) h)->apply(vec(qs, nq), m);
}
void ref_apply_controlled_gate(void* h, const unsigned* qs, unsigned nq,
                               const unsigned* cqs, unsigned nc, uint64_t cvals,
                               const void* m) {
  ((Engine*) h)->apply_c(vec(qs, nq), vec(cqs, nc), cvals, m);
}
void ref_expectation_value(void* h, const unsigned* qs, unsigned nq,
                           const void* m, double out[2]) {
  auto r = ((Engine*) h)->expect(vec(qs, nq), m);
  out[0] = r.real();
  out[1] = r.imag();
}
double ref_norm(void* h) { return ((Engine*) h)->norm(); }
void ref_inner_product(void* h1, void* h2, double out[2]) {
  auto r = ((Engine*) h1)->inner(*(Engine*) h2);
  out[0] = r.real();
  out[1] = r.imag();
}
double ref_real_inner_product(void* h1, void* h2) {
  return ((Engine*) h1)->real_inner(*(Engine*) h2);
}
void ref_multiply(void* h, double a) { ((Engine*) h)->multiply(a); }
int ref_add(void* src, void* dest) { return ((Engine*) dest)->add_from(*(Engine*) src) ? 0 : 1; }
void ref_sample(void* h, uint64_t num, unsigned seed, uint64_t* out) {
  ((Engine*) h)->sample(num, seed, out);
}
int ref_measure(void* h, const unsigned* qs, unsigned nq, unsigned seed,
                uint64_t* mask, uint64_t* bits) {
  return ((Engine*) h)->measure(vec(qs, nq), seed, mask, bits);
}
void ref_collapse(void* h, uint64_t mask, uint64_t bits) { ((Engine*) h)->collapse(mask, bits); }
void ref_get_ampl(void* h, uint64_t i, double out[2]) { ((Engine*) h)->get_ampl(i, out); }
void ref_set_ampl(void* h, uint64_t i, double re, double im) { ((Engine*) h)->set_ampl(i, re, im); }
void ref_bulk_set_ampl(void* h, uint64_t mask, uint64_t bits, double re, double im, int exclude) {
  ((Engine*) h)->bulk_set(mask, bits, re, im, exclude != 0);
}

// The sorted uniform[0,max) sequence Sample() draws (lib/util.h:67-85), so the
// CUDA path and the C oracle can be fed the very same values.
void ref_generate_random_values(uint64_t num, unsigned seed, double max_value, double* out) {
  auto rs = GenerateRandomValues<double>(num, seed, max_value);
  std::memcpy(out, rs.data(), num * sizeof(double));
}

}  // extern "C"
