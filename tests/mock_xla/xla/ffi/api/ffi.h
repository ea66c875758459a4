// MOCK of the subset of jaxlib's xla/ffi/api/ffi.h that dynode_b200/csrc/xla_ffi_shim.cc uses.  TEST INFRASTRUCTURE.
//
// jaxlib (and with it the real header) is absent from this image, so the shim could not meet a compiler at all.
// This model lets tests/test_xla_shim.py compile it here and check what a header-only mock can check: the file is
// valid C++, every C-ABI call in it matches include/*.h, and -- the error-prone part -- every handler's parameter
// list agrees, position by position and type by type, with the Ffi::Bind() chain it is registered with
// (Ctx<PlatformStream<T>> -> T, Arg<B> -> B, Attr<T> -> T, Ret<B> -> Result<B>), as the real binding machinery
// demands.  It does not execute anything and says nothing about XLA's runtime behaviour.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;

namespace xla {
namespace ffi {

enum DataType { PRED, S8, S16, S32, S64, U8, U16, U32, U64, F16, F32, F64, BF16 };
namespace internal {
template <DataType> struct Native;
template <> struct Native<F64> { using type = double; };
template <> struct Native<F32> { using type = float; };
template <> struct Native<S32> { using type = int32_t; };
template <> struct Native<S64> { using type = int64_t; };
template <> struct Native<U8> { using type = uint8_t; };
}  // namespace internal

template <typename T>
class Span {
 public:
  Span() = default;
  Span(T* p, size_t n) : p_(p), n_(n) {}
  T* begin() const { return p_; }
  T* end() const { return p_ + n_; }
  size_t size() const { return n_; }
  T& operator[](size_t i) const { return p_[i]; }

 private:
  T* p_ = nullptr;
  size_t n_ = 0;
};

template <DataType dtype>
class Buffer {
 public:
  using T = typename internal::Native<dtype>::type;
  Buffer() = default;
  // mock-only: wrap caller memory (tests/mock_xla/shim_driver.cc drives the handlers' bodies with device pointers)
  Buffer(T* data, std::vector<int64_t> dims) : data_(data), dims_(std::move(dims)) {}
  T* typed_data() const { return data_; }
  void* untyped_data() const { return data_; }
  Span<const int64_t> dimensions() const { return Span<const int64_t>(dims_.data(), dims_.size()); }
  size_t element_count() const {
    size_t n = 1;
    for (int64_t d : dims_) n *= (size_t)d;
    return n;
  }

 private:
  T* data_ = nullptr;
  std::vector<int64_t> dims_;
};

template <typename T>
class Result {
 public:
  Result() = default;
  explicit Result(T v) : v_(std::move(v)) {}
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_;
};
template <DataType dtype>
using ResultBuffer = Result<Buffer<dtype>>;

enum class ErrorCode { kOk, kCancelled, kUnknown, kInvalidArgument, kInternal, kUnimplemented };
class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  bool failure() const { return code_ != ErrorCode::kOk; }
  const std::string& message() const { return message_; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

template <typename T>
struct PlatformStream {};

template <typename... Ts>
struct Binding {
  template <typename C>
  struct CtxType;
  template <typename T>
  struct CtxType<PlatformStream<T>> { using type = T; };
  template <typename C>
  Binding<Ts..., typename CtxType<C>::type> Ctx() const { return {}; }
  template <typename T>
  Binding<Ts..., T> Arg() const { return {}; }
  template <typename T>
  Binding<Ts..., T> Attr(const char*) const { return {}; }
  template <typename T>
  Binding<Ts..., Result<T>> Ret() const { return {}; }
  // the check the real Bind().To(fn) performs: fn takes exactly the bound parameter list (no implicit conversions:
  // an int32_t attribute bound to a double parameter is a decoding error at run time in XLA)
  template <typename Fn>
  struct Match : std::false_type {};
  template <typename... As>
  struct Match<Error (*)(As...)> : std::is_same<std::tuple<As...>, std::tuple<Ts...>> {};
  template <typename Fn>
  static constexpr bool matches = Match<Fn>::value;
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                                     \
  static_assert(decltype(binding)::template matches<decltype(&impl)>,                                         \
                #name ": the handler's parameters do not match its Ffi::Bind() chain");                        \
  extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame*) { return nullptr; }
