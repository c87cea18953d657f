// Minimal stand-in for jaxlib's xla/ffi/api/ffi.h -- TEST INFRASTRUCTURE ONLY (tests/test_abi_cpu.py compiles
// meanflow_audio_codec_b200/jax_ffi/mfac_jax_ffi.cc against it with -fsyntax-only).  It models just the surface the handlers
// use -- typed buffers, result buffers, Error, the Bind() builder and the handler macro -- so that the handler bodies, and above
// all their calls into include/mfac.h, are type-checked in an image that has no jaxlib.  The binder checks that a handler's
// parameter list matches its Ctx / Arg / Ret / Attr chain in count; it does not dispatch anything.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <vector>

namespace xla::ffi {
enum DataType { F32, U8, S32, BF16 };
template <DataType> struct NativeOf;
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<U8> { using type = uint8_t; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<BF16> { using type = uint16_t; };

template <DataType T> class Buffer {
 public:
  using Native = typename NativeOf<T>::type;
  Native* typed_data() const { return data_; }
  const std::vector<int64_t>& dimensions() const { return dims_; }
  size_t size_bytes() const { return bytes_; }
  size_t element_count() const { return bytes_ / sizeof(Native); }
 private:
  Native* data_ = nullptr;
  std::vector<int64_t> dims_;
  size_t bytes_ = 0;
};
template <typename B> class Result {
 public:
  B* operator->() { return &b_; }
  B& operator*() { return b_; }
 private:
  B b_;
};
template <DataType T> using ResultBuffer = Result<Buffer<T>>;

enum class ErrorCode { kOk, kInternal, kInvalidArgument };
class Error {
 public:
  Error() = default;
  Error(ErrorCode c, std::string m) : code_(c), msg_(std::move(m)) {}
  static Error Success() { return Error(); }
 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string msg_;
};
template <typename T> struct PlatformStream {};

template <int N> struct Binding {
  template <typename T> Binding<N + 1> Ctx() const { return {}; }
  template <typename T> Binding<N + 1> Arg() const { return {}; }
  template <typename T> Binding<N + 1> Ret() const { return {}; }
  template <typename T> Binding<N + 1> Attr(const char*) const { return {}; }
  static constexpr int arity = N;
};
struct Ffi { static Binding<0> Bind() { return {}; } };

template <typename F> struct arity_of;
template <typename R, typename... A> struct arity_of<R (*)(A...)> { static constexpr int value = sizeof...(A); };
template <typename R, typename... A> struct arity_of<R(A...)> { static constexpr int value = sizeof...(A); };
}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, impl, binding)                                                          \
  static_assert(xla::ffi::arity_of<decltype(impl)>::value == decltype(binding)::arity,                             \
                #sym ": handler parameters do not match the Ctx/Arg/Ret/Attr chain");                              \
  extern "C" void* sym(void* call_frame) { (void)call_frame; return nullptr; }
