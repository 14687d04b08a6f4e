// Minimal stand-in for jaxlib's <xla/ffi/api/ffi.h> (absent from this image): just enough of the binding DSL to
// TYPE-CHECK snnquantprune_b200/csrc/xla_ffi_shim.cc.  Ffi::Bind() accumulates the C++ argument types a handler will
// be called with (context, arguments, results, attributes, in binding order) and XLA_FFI_DEFINE_HANDLER_SYMBOL
// static_asserts that the implementation is invocable with exactly those types and returns ffi::Error -- the same
// compile-time contract the real header enforces.  Test infrastructure only (tests/test_host_cpu.py).
#pragma once
#include <cstdint>
#include <string>
#include <type_traits>
#include <vector>

namespace xla {
namespace ffi {

enum class ErrorCode { kOk, kInvalidArgument, kInternal };
class Error {
 public:
  Error() = default;
  Error(ErrorCode c, std::string m) : code_(c), msg_(std::move(m)) {}
  static Error Success() { return Error(); }
  bool failure() const { return code_ != ErrorCode::kOk; }
 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string msg_;
};

enum DataType { U8, S8, S32, S64, F32, U32 };
template <DataType> struct NativeOf;
template <> struct NativeOf<U8> { using type = uint8_t; };
template <> struct NativeOf<S8> { using type = int8_t; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<S64> { using type = int64_t; };
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<U32> { using type = uint32_t; };

template <DataType dt>
class Buffer {
 public:
  using T = typename NativeOf<dt>::type;
  T *typed_data() const { return data_; }
  const std::vector<int64_t> &dimensions() const { return dims_; }
  size_t element_count() const { size_t n = 1; for (auto d : dims_) n *= (size_t)d; return n; }
 private:
  T *data_ = nullptr;
  std::vector<int64_t> dims_;
};
template <DataType dt>
class ResultBuffer {
 public:
  Buffer<dt> *operator->() { return &b_; }
 private:
  Buffer<dt> b_;
};
template <typename T> struct PlatformStream { using type = T; };

template <typename... Ts>
struct Binding {
  template <typename C> Binding<Ts..., typename C::type> Ctx() const { return {}; }
  template <typename A> Binding<Ts..., A> Arg() const { return {}; }
  template <typename A> Binding<Ts..., A> OptionalArg() const { return {}; }
  template <typename R> auto Ret() const { return RetHelper<R>::next(*this); }
  template <typename A> Binding<Ts..., A> Attr(const char *) const { return {}; }
  template <typename R> struct RetHelper;
  template <DataType dt> struct RetHelper<Buffer<dt>> {
    static Binding<Ts..., ResultBuffer<dt>> next(const Binding &) { return {}; }
  };
  template <typename Fn> static constexpr bool Matches() { return std::is_invocable_r<Error, Fn, Ts...>::value; }
};
struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                              \
  static_assert(decltype(binding)::template Matches<decltype(&impl)>(),                                 \
                #impl " is not callable with the types bound for " #name);                             \
  extern "C" void *name() { return reinterpret_cast<void *>(&impl); }
