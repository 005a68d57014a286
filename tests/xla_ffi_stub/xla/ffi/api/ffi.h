// MINIMAL STAND-IN for jaxlib's xla/ffi/api/ffi.h - test infrastructure only (tests/test_xla_ffi_source.py).
// jaxlib is not installable in this image, so the real header is absent; this stub reproduces just enough of the typed
// FFI surface (Buffer / Result / Span / Error / Ffi::Bind() builder / XLA_FFI_DEFINE_HANDLER_SYMBOL) for
// `g++ -fsyntax-only` to type-check jaxrl_b200/csrc/xla_ffi/xla_ffi.cc against include/mapox_b200.h, and it
// static_asserts that every handler's C++ signature is exactly what its Bind() chain decodes:
//   Ctx<PlatformStream<S>> -> S, Arg<T> -> T, Attr<T> -> T, Ret<T> -> Result<T>.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>

namespace xla {
namespace ffi {

enum DataType { PRED, U8, S32, U64, F32, F64, BF16 };
template <DataType dt> struct NativeOf;
template <> struct NativeOf<PRED> { using type = bool; };
template <> struct NativeOf<U8> { using type = uint8_t; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<U64> { using type = uint64_t; };
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<F64> { using type = double; };
template <> struct NativeOf<BF16> { using type = uint16_t; };

template <typename T> struct Span {
  const T* ptr = nullptr;
  size_t n = 0;
  size_t size() const { return n; }
  const T& operator[](size_t i) const { return ptr[i]; }
  const T* begin() const { return ptr; }
  const T* end() const { return ptr + n; }
};

template <DataType dt> struct Buffer {
  using T = typename NativeOf<dt>::type;
  T* typed_data() const { return nullptr; }
  void* untyped_data() const { return nullptr; }
  Span<const int64_t> dimensions() const { return {}; }
  size_t element_count() const { return 0; }
};

template <typename T> struct Result {
  T value;
  T& operator*() { return value; }
  T* operator->() { return &value; }
};
template <DataType dt> using ResultBuffer = Result<Buffer<dt>>;

enum class ErrorCode { kInternal, kInvalidArgument };
struct Error {
  Error() = default;
  Error(ErrorCode, std::string) {}
  static Error Success() { return Error(); }
};

template <typename S> struct PlatformStream {};
enum class Traits { kCmdBufferCompatible };

template <typename... Ts> struct Binding {
  template <typename C> constexpr auto Ctx() const { return ctx_impl(static_cast<C*>(nullptr)); }
  template <typename T> constexpr Binding<Ts..., T> Arg() const { return {}; }
  template <typename T> constexpr Binding<Ts..., Result<T>> Ret() const { return {}; }
  template <typename T> constexpr Binding<Ts..., T> Attr(const char*) const { return {}; }
  template <typename Fn> static constexpr bool matches() { return std::is_same<Fn, Error (*)(Ts...)>::value; }

 private:
  template <typename S> constexpr Binding<Ts..., S> ctx_impl(PlatformStream<S>*) const { return {}; }
};

struct Ffi {
  static constexpr Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, impl, binding, ...)                                         \
  static_assert(decltype(binding)::template matches<decltype(&impl)>(),                                \
                #sym ": the handler's C++ signature does not match its Bind() chain");                 \
  extern "C" void* sym(void* call_frame) { return call_frame; }
