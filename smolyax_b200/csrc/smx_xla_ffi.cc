// XLA FFI (jax.ffi) handlers around the seam twins of libsmolyax_b200.so - for a caller that keeps the reference's
// Python above the two compiled callables and stays inside jax.jit.
//
// Replaces, call for call:
//   reference interpolation.py:243-245 + 293-302   res = jit(vmap(evaluate_tensor_product_interpolant))(x, F, nodes, weights,
//                                                   dims, degs, zetas);  I += jnp.sum(res, axis=0)       -> smx_group_eval
//   reference interpolation.py:246-248 + 334-343   the same with evaluate_tensor_product_gradient         -> smx_group_gradient
//   reference interpolation.py:389                 the einsum of the quadrature                           -> smx_group_integral
// The operands are the reference's own device arrays of one group n (or of a slice [start_s:end_s] of its summands); the
// result is the SUM over the summands of the slice, i.e. what `jnp.sum(res, axis=0)` yields - the (summands, N, d_out)
// intermediate of the vmap is never formed.
//
// Built only where jaxlib's headers exist (smolyax_b200/xla_ffi.py: jax.ffi.include_dir()); this image has no jax, so the
// repository's CPU test compiles the file against a stub of the few FFI declarations it uses (tests/xla_ffi_stub) and the
// real build + parity test is skipped there.  No torch, no Python in this file: XLA FFI types in, C-ABI calls out.
#include <cuda_runtime_api.h>

#include <cstdint>
#include <string>
#include <vector>

#include "smolyax_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

// F is (nn, d_out, tau_1 + 1, .., tau_n + 1); nodes / weights (nn, n, m); dims / degs (nn, n); zetas (nn)
struct Group {
    std::vector<int64_t> tau;
    smx_group_desc desc;
    int64_t d_out;
};

ffi::Error describe(const ffi::AnyBuffer::Dimensions& fd, const double* F, const double* nodes, const double* weights, const int64_t* dims,
                    const int64_t* degs, const int64_t* zetas, const double* quad, ffi::AnyBuffer::Dimensions nd, Group* g) {
    if (fd.size() < 3) return ffi::Error::InvalidArgument("F must have shape (summands, d_out, tau_1 + 1, ..)");
    if (nd.size() != 3 || nd[0] != fd[0] || nd[1] != (int64_t)fd.size() - 2)
        return ffi::Error::InvalidArgument("nodes must have shape (summands, n, max degree + 1) with n = F.ndim - 2");
    g->tau.assign(fd.begin() + 2, fd.end());
    for (int64_t& t : g->tau) t -= 1;
    g->d_out = fd[1];
    g->desc.n = (int32_t)g->tau.size();
    g->desc.nn = fd[0];
    g->desc.tau = g->tau.data();
    g->desc.F = F, g->desc.nodes = nodes, g->desc.weights = weights;
    g->desc.dims = dims, g->desc.degs = degs, g->desc.zetas = zetas, g->desc.quad = quad;
    return ffi::Error::Success();
}

ffi::Error status(int rc) {
    if (rc == SMX_OK) return ffi::Error::Success();
    const std::string msg = std::string("smolyax_b200: ") + smx_last_error();
    return rc == SMX_ERR_INVALID_ARG ? ffi::Error::InvalidArgument(msg) : ffi::Error::Internal(msg);
}

ffi::Error GroupEvalImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F64> F, ffi::Buffer<ffi::F64> nodes,
                         ffi::Buffer<ffi::F64> weights, ffi::Buffer<ffi::S64> dims, ffi::Buffer<ffi::S64> degs,
                         ffi::Buffer<ffi::S64> zetas, ffi::ResultBuffer<ffi::F64> y) {
    const auto xd = x.dimensions();
    if (xd.size() != 2) return ffi::Error::InvalidArgument("x must have shape (N, d_in)");
    Group g;
    if (ffi::Error e = describe(F.dimensions(), F.typed_data(), nodes.typed_data(), weights.typed_data(), dims.typed_data(),
                                degs.typed_data(), zetas.typed_data(), nullptr, nodes.dimensions(), &g);
        e.failure())
        return e;
    const auto yd = y->dimensions();
    if (yd.size() != 2 || yd[0] != xd[0] || yd[1] != g.d_out) return ffi::Error::InvalidArgument("result must have shape (N, d_out)");
    return status(smx_group_eval(x.typed_data(), xd[0], xd[1], xd[1], &g.desc, g.d_out, y->typed_data(), /*accumulate=*/0, stream));
}

ffi::Error GroupGradientImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F64> F, ffi::Buffer<ffi::F64> nodes,
                             ffi::Buffer<ffi::F64> weights, ffi::Buffer<ffi::S64> dims, ffi::Buffer<ffi::S64> degs,
                             ffi::Buffer<ffi::S64> zetas, ffi::ResultBuffer<ffi::F64> J) {
    const auto xd = x.dimensions();
    if (xd.size() != 2) return ffi::Error::InvalidArgument("x must have shape (N, d_in)");
    Group g;
    if (ffi::Error e = describe(F.dimensions(), F.typed_data(), nodes.typed_data(), weights.typed_data(), dims.typed_data(),
                                degs.typed_data(), zetas.typed_data(), nullptr, nodes.dimensions(), &g);
        e.failure())
        return e;
    const auto jd = J->dimensions();
    if (jd.size() != 3 || jd[0] != xd[0] || jd[1] != g.d_out || jd[2] != xd[1])
        return ffi::Error::InvalidArgument("result must have shape (N, d_out, d_in)");
    return status(smx_group_gradient(x.typed_data(), xd[0], xd[1], xd[1], &g.desc, g.d_out, J->typed_data(), /*accumulate=*/0, stream));
}

// quad: (nn, n, m) quadrature weights per summand and slot (interpolation.py:361-379)
ffi::Error GroupIntegralImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> F, ffi::Buffer<ffi::F64> quad, ffi::Buffer<ffi::S64> zetas,
                             ffi::ResultBuffer<ffi::F64> q) {
    Group g;
    if (ffi::Error e = describe(F.dimensions(), F.typed_data(), nullptr, nullptr, nullptr, nullptr, zetas.typed_data(), quad.typed_data(),
                                quad.dimensions(), &g);
        e.failure())
        return e;
    const auto qd = q->dimensions();
    if (qd.size() != 1 || qd[0] != g.d_out) return ffi::Error::InvalidArgument("result must have shape (d_out,)");
    return status(smx_group_integral(&g.desc, g.d_out, q->typed_data(), /*accumulate=*/0, stream));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(SmxGroupEval, GroupEvalImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()   // x        (N, d_in)
                                  .Arg<ffi::Buffer<ffi::F64>>()   // F        (nn, d_out, tau_1 + 1, ..)
                                  .Arg<ffi::Buffer<ffi::F64>>()   // nodes    (nn, n, m)
                                  .Arg<ffi::Buffer<ffi::F64>>()   // weights  (nn, n, m)
                                  .Arg<ffi::Buffer<ffi::S64>>()   // dims     (nn, n)
                                  .Arg<ffi::Buffer<ffi::S64>>()   // degs     (nn, n)
                                  .Arg<ffi::Buffer<ffi::S64>>()   // zetas    (nn)
                                  .Ret<ffi::Buffer<ffi::F64>>());  // y       (N, d_out)

XLA_FFI_DEFINE_HANDLER_SYMBOL(SmxGroupGradient, GroupGradientImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::S64>>()
                                  .Arg<ffi::Buffer<ffi::S64>>()
                                  .Arg<ffi::Buffer<ffi::S64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());  // J       (N, d_out, d_in)

XLA_FFI_DEFINE_HANDLER_SYMBOL(SmxGroupIntegral, GroupIntegralImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()   // F
                                  .Arg<ffi::Buffer<ffi::F64>>()   // quad
                                  .Arg<ffi::Buffer<ffi::S64>>()   // zetas
                                  .Ret<ffi::Buffer<ffi::F64>>());  // q       (d_out)
