// vectorspace_b200.h -- drop-in for qsim's VectorSpaceCUDA (lib/vectorspace_cuda.h:43-167)
// over the libqsim_b200.so C ABI.  Plain C++17: client code is compiled by g++,
// no nvcc and no CUDA headers needed.  Consumed next to the reference headers
// (-I$QSIM/lib -I<this repo>/include).
#ifndef QSIM_B200_VECTORSPACE_B200_H_
#define QSIM_B200_VECTORSPACE_B200_H_

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <type_traits>
#include <utility>

#include "../qsim_b200.h"

namespace qsim {
namespace b200 {

template <typename FP> struct DType;
template <> struct DType<float> { static constexpr int value = QB200_F32; };
template <> struct DType<double> { static constexpr int value = QB200_F64; };

// Error convention of the reference (lib/util_cuda.h:31-39): a failed CUDA call
// prints "CUDA error: ..." and exits; invalid-argument statuses are reported the
// same way because the reference would have crashed there.
inline void Check(int status, const qb200_ctx* ctx, const char* file, int line) {
  if (status == QB200_OK || status == QB200_ERR_UNSUPPORTED) return;
  const char* what = status == QB200_ERR_CUDA ? qb200_last_cuda_error_string(ctx)
                   : status == QB200_ERR_OOM ? "out of memory" : "invalid argument";
  std::fprintf(stderr, "CUDA error: %s at %s %d\n", what, file, line);
  std::exit(status == QB200_ERR_CUDA ? qb200_last_cuda_error(ctx) : status);
}
#define QB200_CHECK(ctx, call) ::qsim::b200::Check((call), (ctx), __FILE__, __LINE__)

// SetStream(kStreamPerThread): CUDA's per-thread default stream (cudaStreamPerThread) -- every host thread gets its
// own stream without creating one; what the multi-worker trajectory app uses (apps/qsim_qtrajectory_b200.cc -j).
static void* const kStreamPerThread = reinterpret_cast<void*>(0x2);

// One context per backend object; copies of the object share it.
inline std::shared_ptr<qb200_ctx> MakeContext(int device = -1) {
  qb200_ctx* ctx = nullptr;
  int rc = qb200_ctx_create(device, &ctx);
  if (rc != QB200_OK) {
    std::fprintf(stderr, "CUDA error: no usable CUDA device (qsim_b200 has no CPU fallback) at %s %d\n",
                 __FILE__, __LINE__);
    std::exit(1);
  }
  return std::shared_ptr<qb200_ctx>(ctx, [](qb200_ctx* c) { qb200_ctx_destroy(c); });
}

// Context used by the reference's *static* members (GetAmpl, SetAmpl, DeviceSync).
inline qb200_ctx* ThreadContext() {
  thread_local std::shared_ptr<qb200_ctx> ctx = MakeContext();
  return ctx.get();
}

namespace detail {
inline void do_not_free(void*) {}
inline void free(void* ptr) { qb200_state_free(ptr); }
}  // namespace detail

}  // namespace b200

// lib/vectorspace_cuda.h:31-39 publishes its deleter as qsim::detail::free and the pybind layer
// names it (pybind_interface/pybind_main.cpp:455); a backend header provides it, exactly one
// backend header per translation unit (the reference's CPU and CUDA headers clash the same way).
#if !defined(VECTORSPACE_H_) && !defined(VECTORSPACE_CUDA_H_)
namespace detail {
inline void do_not_free(void*) {}
inline void free(void* ptr) { qb200_state_free(ptr); }
}  // namespace detail
#endif

template <typename Impl, typename FP>
class VectorSpaceB200 {
 public:
  using fp_type = FP;

 private:
  using Pointer = std::unique_ptr<fp_type, decltype(&b200::detail::free)>;

 public:
  class Vector {
   public:
    Vector() = delete;
    Vector(Pointer&& ptr, unsigned num_qubits) : ptr_(std::move(ptr)), num_qubits_(num_qubits) {}

    fp_type* get() { return ptr_.get(); }
    const fp_type* get() const { return ptr_.get(); }

    fp_type* release() {
      num_qubits_ = 0;
      return ptr_.release();
    }

    unsigned num_qubits() const { return num_qubits_; }

    // device memory: pybind copies the state out (pybind_main.cpp:431-460)
    static constexpr bool requires_copy_to_host() { return true; }

   private:
    Pointer ptr_;
    unsigned num_qubits_;
  };

  template <typename... Args>
  VectorSpaceB200(Args&&...) : ctx_(b200::MakeContext()) {}

  static Vector Create(unsigned num_qubits) {
    void* p = nullptr;
    int rc = qb200_state_alloc(num_qubits, b200::DType<FP>::value, &p);
    if (rc == QB200_OK) {
      return Vector{Pointer{(fp_type*) p, &b200::detail::free}, num_qubits};
    }
    return Null();  // "not enough memory" is the caller's message (lib/run_qsim.h:93-97)
  }

  // It is the client's responsibility to make sure that p has at least
  // Impl::MinSize(num_qubits) elements of device memory.
  static Vector Create(fp_type* p, unsigned num_qubits) {
    return Vector{Pointer{p, &b200::detail::do_not_free}, num_qubits};
  }

  static Vector Null() { return Vector{Pointer{nullptr, &b200::detail::free}, 0}; }
  static bool IsNull(const Vector& vector) { return vector.get() == nullptr; }
  static void Free(fp_type* ptr) { b200::detail::free(ptr); }

  bool Copy(const Vector& src, Vector& dest) const {
    if (src.num_qubits() != dest.num_qubits()) return false;
    QB200_CHECK(ctx(), qb200_copy_d2d(ctx(), b200::DType<FP>::value, src.get(), dest.get(),
                                      Impl::MinSize(src.num_qubits())));
    return true;
  }

  // Not in the reference: the same copy enqueued on this object's stream, the host does not wait
  // (qtrajectory_b200.h restores checkpoints with it).
  bool CopyAsync(const Vector& src, Vector& dest) const {
    if (src.num_qubits() != dest.num_qubits()) return false;
    QB200_CHECK(ctx(), qb200_copy_d2d_async(ctx(), b200::DType<FP>::value, src.get(), dest.get(),
                                            Impl::MinSize(src.num_qubits())));
    return true;
  }

  bool Copy(const Vector& src, fp_type* dest) const {
    QB200_CHECK(ctx(), qb200_copy_d2h(ctx(), b200::DType<FP>::value, src.get(), dest,
                                      Impl::MinSize(src.num_qubits())));
    return true;
  }

  bool Copy(const fp_type* src, Vector& dest) const {
    QB200_CHECK(ctx(), qb200_copy_h2d(ctx(), b200::DType<FP>::value, src, dest.get(),
                                      Impl::MinSize(dest.num_qubits())));
    return true;
  }

  bool Copy(const fp_type* src, uint64_t size, Vector& dest) const {
    size = std::min(size, Impl::MinSize(dest.num_qubits()));
    QB200_CHECK(ctx(), qb200_copy_h2d(ctx(), b200::DType<FP>::value, src, dest.get(), size));
    return true;
  }

  static void DeviceSync() { QB200_CHECK(nullptr, qb200_device_sync()); }

  // Not in the reference: lets a caller put this object's work on its own stream.
  void SetStream(void* cuda_stream) const { QB200_CHECK(ctx(), qb200_ctx_set_stream(ctx(), cuda_stream)); }

 protected:
  qb200_ctx* ctx() const { return ctx_.get(); }
  std::shared_ptr<qb200_ctx> ctx_;
};

}  // namespace qsim

#endif  // QSIM_B200_VECTORSPACE_B200_H_
