// TEST INFRASTRUCTURE: a minimal stand-in for jaxlib's xla/ffi/api/ffi.h (not installable in this image), just
// enough of the typed-FFI binder to TYPE-CHECK brax_tracking_b200/csrc/xla_ffi_shim.cc: the binder accumulates the decoded
// parameter types of Ctx / Attr / Arg / Ret in order and `To(handler)` requires the handler to be invocable with exactly those
// types, which is the rule the real header enforces (an arity or type mismatch between a binding and its handler -- ADVICE r1 --
// fails the compile here as it would there).  No runtime behaviour.
#pragma once
#include <cstdint>
#include <string>
#include <type_traits>
#include <vector>

namespace xla {
namespace ffi {

enum DataType { F32, S32, U32 };
template <DataType> struct NativeOf;
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<U32> { using type = uint32_t; };

struct Dims { const int64_t* p; int64_t n; int64_t operator[](int i) const { return p[i]; } int64_t size() const { return n; } };

template <DataType T>
struct Buffer {
  using Native = typename NativeOf<T>::type;
  Native* data = nullptr;
  std::vector<int64_t> dims;
  Native* typed_data() const { return data; }
  Dims dimensions() const { return Dims{dims.data(), (int64_t)dims.size()}; }
  size_t size_bytes() const { size_t n = sizeof(Native); for (auto d : dims) n *= (size_t)d; return n; }
};
template <typename T> struct Result { T value; T* operator->() { return &value; } T& operator*() { return value; } };
template <DataType T> using ResultBuffer = Result<Buffer<T>>;

enum class ErrorCode { kInternal, kInvalidArgument };
struct Error {
  bool ok = true;
  Error() = default;
  Error(ErrorCode, std::string) : ok(false) {}
  static Error Success() { return Error(); }
};

template <typename T> struct PlatformStream {};
template <typename T> struct CtxDecode;
template <typename T> struct CtxDecode<PlatformStream<T>> { using type = T; };

template <typename... Ts>
struct Binding {
  template <typename C> Binding<Ts..., typename CtxDecode<C>::type> Ctx() { return {}; }
  template <typename A> Binding<Ts..., A> Attr(const char*) { return {}; }
  template <typename A> Binding<Ts..., A> Arg() { return {}; }
  template <typename R> Binding<Ts..., Result<R>> Ret() { return {}; }
  template <typename F> int To(F) {
    static_assert(std::is_invocable_r_v<Error, F, Ts...>, "handler signature does not match the FFI binding");
    return (int)sizeof...(Ts);
  }
};
struct Ffi { static Binding<> Bind() { return {}; } };

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, fn, binding) \
  extern "C" int sym##_arity() { return (binding).To(fn); }
