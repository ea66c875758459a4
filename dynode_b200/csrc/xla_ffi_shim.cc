// xla_ffi_shim.cc -- typed XLA-FFI handlers over the C ABI of include/dynode_b200.h, so that
// DynODE's JAX code reaches the sm_100a kernels through jax.ffi.ffi_call.
//
// NOT compiled in this image: the XLA FFI headers ship with jaxlib (jax.ffi.include_dir()), which is
// absent here.  dynode_b200/_build.py builds libdynode_b200_xla.so from this file only when
// `import jax` works.  The handlers only enqueue on the stream XLA passes in; XLA owns every buffer.
//
// Replaces, at the JAX level, diffrax.diffeqsolve(...) in reference src/dynode/simulation/odes.py:133-144
// under jax.vmap (leading batch axis on y0 and the rate arrays).
#include <cuda_runtime.h>

#include "../../include/dynode_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

using F64 = ffi::Buffer<ffi::F64>;
using F64Out = ffi::ResultBuffer<ffi::F64>;
using S32Out = ffi::ResultBuffer<ffi::S32>;

// [B][row] buffer -> DynodeArray; a buffer without the batch axis (or with B == 1) is shared.
DynodeArray as_array(const F64& b, int64_t B) {
  DynodeArray a;
  a.ptr = b.element_count() > 0 ? b.typed_data() : nullptr;
  const int64_t lead = b.dimensions().size() > 0 ? b.dimensions()[0] : 1;
  a.batch_stride = (lead == B && B > 1) ? (int64_t)(b.element_count() / B) : 0;
  return a;
}

DynodeParams pack(const F64& beta, const F64& gamma, const F64& sigma, const F64& omega, const F64& season,
                  const F64& contact, int64_t B) {
  DynodeParams p{};
  p.beta = as_array(beta, B);
  p.gamma = as_array(gamma, B);
  p.sigma = as_array(sigma, B);
  p.omega = as_array(omega, B);
  if (season.element_count() > 0) {  // [B][3] = (amp, phase, period)
    const int64_t stride = as_array(season, B).batch_stride;
    p.season_amp = {season.typed_data() + 0, stride};
    p.season_phase = {season.typed_data() + 1, stride};
    p.season_period = {season.typed_data() + 2, stride};
  }
  p.contact = contact.element_count() > 0 ? contact.typed_data() : nullptr;
  return p;
}

ffi::Error SolveImpl(cudaStream_t stream, F64 y0, F64 beta, F64 gamma, F64 sigma, F64 omega, F64 season,
                     F64 contact, F64 save_ts, int32_t flow, int32_t flags, int32_t n_groups,
                     int32_t n_strains, int64_t save_mask, double t0, double t1, double rtol, double atol,
                     double const_dt, int64_t max_steps, double save_dt, F64Out ys, S32Out stats) {
  const DynodeModelDesc model{flow, flags, n_groups, n_strains};
  const DynodeSolverDesc solver{t0, t1, rtol, atol, const_dt, max_steps, save_dt};
  const int64_t B = stats->dimensions()[0];
  const DynodeParams p = pack(beta, gamma, sigma, omega, season, contact, B);
  if (dynode_solve_f64(&model, &solver, B, as_array(y0, B), &p, save_ts.typed_data(),
                       (int32_t)save_ts.element_count(), (uint32_t)save_mask, ys->typed_data(),
                       stats->typed_data(), stream) != 0)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, dynode_last_error());
  return ffi::Error::Success();
}

ffi::Error LoglikGradImpl(cudaStream_t stream, F64 y0, F64 beta, F64 gamma, F64 sigma, F64 omega, F64 season,
                          F64 contact, F64 save_ts, F64 obs, F64 dy0, ffi::Span<const int32_t> wrt, int32_t flow,
                          int32_t flags, int32_t n_groups, int32_t n_strains, int32_t obs_comp, double lp_const,
                          double t0, double t1, double rtol, double atol, double const_dt, int64_t max_steps,
                          double save_dt, F64Out lp, F64Out grad, S32Out stats) {
  const DynodeModelDesc model{flow, flags, n_groups, n_strains};
  const DynodeSolverDesc solver{t0, t1, rtol, atol, const_dt, max_steps, save_dt};
  const int64_t B = stats->dimensions()[0];
  const DynodeParams p = pack(beta, gamma, sigma, omega, season, contact, B);
  if (dynode_poisson_loglik_grad_f64(&model, &solver, B, as_array(y0, B), &p, save_ts.typed_data(),
                                     (int32_t)save_ts.element_count(), obs_comp, obs.typed_data(), lp_const,
                                     (int32_t)wrt.size(), wrt.begin(),
                                     dy0.element_count() > 0 ? dy0.typed_data() : nullptr, lp->typed_data(),
                                     grad->typed_data(), stats->typed_data(), stream) != 0)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, dynode_last_error());
  return ffi::Error::Success();
}

}  // namespace

#define DYN_MODEL_ATTRS \
  .Attr<int32_t>("flow").Attr<int32_t>("flags").Attr<int32_t>("n_groups").Attr<int32_t>("n_strains")
#define DYN_SOLVER_ATTRS                                                                         \
  .Attr<double>("t0").Attr<double>("t1").Attr<double>("rtol").Attr<double>("atol")                \
      .Attr<double>("const_dt").Attr<int64_t>("max_steps").Attr<double>("save_dt")
#define DYN_INPUTS                                                                               \
  .Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>() \
      .Arg<F64>().Arg<F64>().Arg<F64>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(DynodeSolve, SolveImpl,
                              ffi::Ffi::Bind() DYN_INPUTS DYN_MODEL_ATTRS.Attr<int64_t>("save_mask")
                                  DYN_SOLVER_ATTRS.Ret<F64>().Ret<ffi::Buffer<ffi::S32>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(DynodePoissonLoglikGrad, LoglikGradImpl,
                              ffi::Ffi::Bind() DYN_INPUTS.Arg<F64>().Arg<F64>()
                                  .Attr<ffi::Span<const int32_t>>("wrt") DYN_MODEL_ATTRS.Attr<int32_t>("obs_comp")
                                  .Attr<double>("lp_const") DYN_SOLVER_ATTRS.Ret<F64>().Ret<F64>()
                                  .Ret<ffi::Buffer<ffi::S32>>());
