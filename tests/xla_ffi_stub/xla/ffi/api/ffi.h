// TEST STUB - not XLA.  The handful of declarations of jaxlib's "xla/ffi/api/ffi.h" that smolyax_b200/csrc/smx_xla_ffi.cc
// uses, with the same names and shapes, so that tests/test_xla_ffi.py can check on a machine without jaxlib that the
// translation unit is well-formed C++ and binds its handlers with the argument lists it declares.  The real header is found
// through jax.ffi.include_dir() (smolyax_b200/xla_ffi.py); this file is never on an include path of the product.
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

namespace xla::ffi {

enum class DataType { F64, S64 };
inline constexpr DataType F64 = DataType::F64;
inline constexpr DataType S64 = DataType::S64;

template <typename T>
class Span {
public:
    Span(const T* d, size_t n) : d_(d), n_(n) {}
    size_t size() const { return n_; }
    const T& operator[](size_t i) const { return d_[i]; }
    const T* begin() const { return d_; }
    const T* end() const { return d_ + n_; }

private:
    const T* d_;
    size_t n_;
};

class AnyBuffer {
public:
    using Dimensions = Span<int64_t>;
};

template <DataType dtype>
struct NativeOf;
template <>
struct NativeOf<DataType::F64> {
    using type = double;
};
template <>
struct NativeOf<DataType::S64> {
    using type = int64_t;
};

template <DataType dtype>
class Buffer {
public:
    using T = typename NativeOf<dtype>::type;
    T* typed_data() const { return data_; }
    AnyBuffer::Dimensions dimensions() const { return AnyBuffer::Dimensions(dims_.data(), dims_.size()); }

private:
    T* data_ = nullptr;
    std::vector<int64_t> dims_;
};

template <typename T>
class Result {
public:
    T* operator->() { return &value_; }
    T& operator*() { return value_; }

private:
    T value_;
};
template <DataType dtype>
using ResultBuffer = Result<Buffer<dtype>>;

class Error {
public:
    static Error Success() { return Error(false, ""); }
    static Error Internal(std::string m) { return Error(true, std::move(m)); }
    static Error InvalidArgument(std::string m) { return Error(true, std::move(m)); }
    bool failure() const { return failed_; }
    bool success() const { return !failed_; }
    const std::string& message() const { return msg_; }

private:
    Error(bool f, std::string m) : failed_(f), msg_(std::move(m)) {}
    bool failed_;
    std::string msg_;
};

template <typename T>
struct PlatformStream {};

// Binding: records the handler's signature as a type list; the handler macro checks that the implementation is callable
// with exactly those arguments (stream first, then the buffers, then the results)
template <typename S>
struct StreamOf;
template <typename S>
struct StreamOf<PlatformStream<S>> {
    using type = S;
};
template <typename... Ts>
struct BindingWithStream {
    template <typename T>
    BindingWithStream<Ts..., T> Arg() const { return {}; }
    template <typename T>
    BindingWithStream<Ts..., Result<T>> Ret() const { return {}; }
    template <typename Fn>
    static constexpr bool Accepts() { return std::is_invocable_r_v<Error, Fn, Ts...>; }
};
struct Ffi {
    struct Start {
        template <typename T>
        BindingWithStream<typename StreamOf<T>::type> Ctx() const { return {}; }
    };
    static Start Bind() { return {}; }
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                                      \
    static_assert(decltype(binding)::template Accepts<decltype(&impl)>(), #impl " does not match its binding"); \
    extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame*) { return nullptr; }
