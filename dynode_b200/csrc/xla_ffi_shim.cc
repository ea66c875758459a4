// xla_ffi_shim.cc -- typed XLA-FFI handlers over the C ABI (include/dynode_b200*.h), so that DynODE's JAX code
// reaches the sm_100a kernels through jax.ffi.ffi_call.
//
// One handler per launch entry point of the path:
//   DynodeSolve                 dynode_solve_f64                   diffeqsolve + SaveAt (odes.py:133-144)
//   DynodeSolveSens             dynode_solve_sens_f64              the same with forward tangents (jvp / small-P vjp)
//   DynodePoissonLoglikGrad     dynode_poisson_loglik_grad_f64     fused NUTS log-likelihood, forward mode
//   DynodePoissonLoglikAdjoint  dynode_poisson_loglik_adjoint_f64  fused NUTS log-likelihood, discrete adjoint
//   DynodeSeipSolve             dynode_seip_solve_f64              immune-history family (one CTA per trajectory)
//   DynodeSiteLogdensity / DynodeSiteLogdensityVjp                 a latent site's bijector + prior, each way
// Every solver handler takes `jump_ts` (SolverParams.discontinuity_points, odes.py:120-131) and `only` (row mask) as
// trailing input buffers; a zero-element buffer means "none".  Scratch of the adjoint arrives as extra results, so XLA
// owns and reuses it.
//
// The real headers ship with jaxlib (jax.ffi.include_dir()), absent from this image: dynode_b200/_build.py::
// build_xla_shim compiles this file against them wherever `import jax` works; tests/test_xla_shim.py compiles it here
// against tests/mock_xla (a model of the subset of xla/ffi/api/ffi.h used below that checks every handler's signature
// against its binding), and says loudly which of the two it did.  The handlers only enqueue on the stream XLA passes
// in; XLA owns every buffer.
#include <cuda_runtime.h>

#include "../../include/dynode_b200.h"
#include "../../include/dynode_b200_ppl.h"
#include "../../include/dynode_b200_seip.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

using F64 = ffi::Buffer<ffi::F64>;
using U8 = ffi::Buffer<ffi::U8>;
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

DynodeSolverDesc solver_desc(double t0, double t1, double rtol, double atol, double const_dt, int64_t max_steps,
                             double save_dt, const F64& jump_ts, const U8& only) {
  DynodeSolverDesc s{};
  s.t0 = t0; s.t1 = t1; s.rtol = rtol; s.atol = atol; s.const_dt = const_dt; s.max_steps = max_steps;
  s.save_dt = save_dt;
  // the reference ignores discontinuity points in constant-step mode (odes.py:113-131)
  const bool jumps = jump_ts.element_count() > 0 && !(const_dt > 0.0);
  s.jump_ts = jumps ? jump_ts.typed_data() : nullptr;
  s.n_jump = jumps ? (int32_t)jump_ts.element_count() : 0;
  s.only = only.element_count() > 0 ? only.typed_data() : nullptr;
  return s;
}

ffi::Error status(int rc) {
  return rc == 0 ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInvalidArgument, dynode_last_error());
}

#define DYN_MODEL_PARAMS int32_t flow, int32_t flags, int32_t n_groups, int32_t n_strains
#define DYN_SOLVER_PARAMS \
  double t0, double t1, double rtol, double atol, double const_dt, int64_t max_steps, double save_dt
#define DYN_SOLVER_ARGS t0, t1, rtol, atol, const_dt, max_steps, save_dt

ffi::Error SolveImpl(cudaStream_t stream, F64 y0, F64 beta, F64 gamma, F64 sigma, F64 omega, F64 season, F64 contact,
                     F64 save_ts, F64 jump_ts, U8 only, DYN_MODEL_PARAMS, int64_t save_mask, DYN_SOLVER_PARAMS,
                     F64Out ys, S32Out stats) {
  const DynodeModelDesc model{flow, flags, n_groups, n_strains};
  const DynodeSolverDesc solver = solver_desc(DYN_SOLVER_ARGS, jump_ts, only);
  const int64_t B = stats->dimensions()[0];
  const DynodeParams p = pack(beta, gamma, sigma, omega, season, contact, B);
  return status(dynode_solve_f64(&model, &solver, B, as_array(y0, B), &p, save_ts.typed_data(),
                                 (int32_t)save_ts.element_count(), (uint32_t)save_mask, ys->typed_data(),
                                 stats->typed_data(), stream));
}

ffi::Error SolveSensImpl(cudaStream_t stream, F64 y0, F64 beta, F64 gamma, F64 sigma, F64 omega, F64 season,
                         F64 contact, F64 save_ts, F64 jump_ts, U8 only, F64 dy0, ffi::Span<const int32_t> wrt,
                         DYN_MODEL_PARAMS, int64_t save_mask, DYN_SOLVER_PARAMS, F64Out ys, F64Out dys,
                         S32Out stats) {
  const DynodeModelDesc model{flow, flags, n_groups, n_strains};
  const DynodeSolverDesc solver = solver_desc(DYN_SOLVER_ARGS, jump_ts, only);
  const int64_t B = stats->dimensions()[0];
  const DynodeParams p = pack(beta, gamma, sigma, omega, season, contact, B);
  return status(dynode_solve_sens_f64(&model, &solver, B, as_array(y0, B), &p, save_ts.typed_data(),
                                      (int32_t)save_ts.element_count(), (uint32_t)save_mask, (int32_t)wrt.size(),
                                      wrt.begin(), dy0.element_count() > 0 ? dy0.typed_data() : nullptr,
                                      ys->typed_data(), dys->typed_data(), stats->typed_data(), stream));
}

ffi::Error LoglikGradImpl(cudaStream_t stream, F64 y0, F64 beta, F64 gamma, F64 sigma, F64 omega, F64 season,
                          F64 contact, F64 save_ts, F64 jump_ts, U8 only, F64 obs, F64 dy0,
                          ffi::Span<const int32_t> wrt, DYN_MODEL_PARAMS, int32_t obs_comp, double lp_const,
                          DYN_SOLVER_PARAMS, F64Out lp, F64Out grad, S32Out stats) {
  const DynodeModelDesc model{flow, flags, n_groups, n_strains};
  const DynodeSolverDesc solver = solver_desc(DYN_SOLVER_ARGS, jump_ts, only);
  const int64_t B = stats->dimensions()[0];
  const DynodeParams p = pack(beta, gamma, sigma, omega, season, contact, B);
  return status(dynode_poisson_loglik_grad_f64(
      &model, &solver, B, as_array(y0, B), &p, save_ts.typed_data(), (int32_t)save_ts.element_count(), obs_comp,
      obs.typed_data(), lp_const, (int32_t)wrt.size(), wrt.begin(),
      dy0.element_count() > 0 ? dy0.typed_data() : nullptr, lp->typed_data(), grad->typed_data(),
      stats->typed_data(), stream));
}

// ckpt [B][cap][n + 2] and vsave [B][T][m] are results XLA allocates (scratch); grad_y0 may be a zero-element result.
ffi::Error LoglikAdjointImpl(cudaStream_t stream, F64 y0, F64 beta, F64 gamma, F64 sigma, F64 omega, F64 season,
                             F64 contact, F64 save_ts, F64 jump_ts, U8 only, F64 obs, DYN_MODEL_PARAMS,
                             int32_t obs_comp, double lp_const, int32_t cap, DYN_SOLVER_PARAMS, F64Out lp,
                             F64Out grad, F64Out grad_y0, S32Out stats, F64Out ckpt, F64Out vsave) {
  const DynodeModelDesc model{flow, flags, n_groups, n_strains};
  const DynodeSolverDesc solver = solver_desc(DYN_SOLVER_ARGS, jump_ts, only);
  const int64_t B = stats->dimensions()[0];
  const DynodeParams p = pack(beta, gamma, sigma, omega, season, contact, B);
  return status(dynode_poisson_loglik_adjoint_f64(
      &model, &solver, B, as_array(y0, B), &p, save_ts.typed_data(), (int32_t)save_ts.element_count(), obs_comp,
      obs.typed_data(), lp_const, lp->typed_data(), grad->typed_data(),
      grad_y0->element_count() > 0 ? grad_y0->typed_data() : nullptr, stats->typed_data(), ckpt->typed_data(), cap,
      vsave->typed_data(), stream));
}

// zero-element buffers = absent (no vaccination tables / no introductions / no discontinuity points / no row mask)
ffi::Error SeipSolveImpl(cudaStream_t stream, F64 y0, F64 beta, F64 sigma, F64 gamma, F64 omega, F64 contact, F64 pop,
                         F64 immunity, F64 vax_base, F64 vax_knots, F64 vax_coef, F64 intro_time, F64 intro_scale,
                         F64 intro_pct, F64 intro_ages, F64 save_ts, F64 jump_ts, U8 only, int32_t n_ages,
                         int32_t n_strains, int32_t n_wane, int32_t n_vax, int32_t n_knots, int64_t save_mask,
                         double season_tau, double season_on, DYN_SOLVER_PARAMS, F64Out ys, S32Out stats) {
  DynodeSeipDesc model{};
  model.n_ages = n_ages; model.n_strains = n_strains; model.n_wane = n_wane; model.n_vax = n_vax;
  model.n_knots = n_knots; model.save_mask = (uint32_t)save_mask;
  const DynodeSolverDesc solver = solver_desc(DYN_SOLVER_ARGS, jump_ts, only);
  const int64_t B = stats->dimensions()[0];
  auto opt = [](const F64& b) -> const double* { return b.element_count() > 0 ? b.typed_data() : nullptr; };
  DynodeSeipParams p{};
  p.beta = as_array(beta, B); p.sigma = as_array(sigma, B); p.gamma = as_array(gamma, B); p.omega = as_array(omega, B);
  p.contact = contact.typed_data(); p.pop = pop.typed_data(); p.immunity = immunity.typed_data();
  p.vax_base = opt(vax_base); p.vax_knots = opt(vax_knots); p.vax_coef = opt(vax_coef);
  p.intro_time = as_array(intro_time, B); p.intro_scale = as_array(intro_scale, B);
  p.intro_pct = as_array(intro_pct, B); p.intro_ages = opt(intro_ages);
  p.season_tau = season_tau; p.season_on = season_on;
  return status(dynode_seip_solve_f64(&model, &solver, B, as_array(y0, B), &p, save_ts.typed_data(),
                                      (int32_t)save_ts.element_count(), ys->typed_data(), stats->typed_data(),
                                      stream));
}

#define DYN_SITE_PARAMS                                                                                          \
  int32_t bijector, int32_t family, double a, double b, double p0, double p1, double c, double aff_loc, \
      double aff_scale
#define DYN_SITE_INIT {bijector, family, a, b, p0, p1, c, aff_loc, aff_scale}

ffi::Error SiteImpl(cudaStream_t stream, F64 z, DYN_SITE_PARAMS, F64Out x, F64Out lp) {
  const DynodeSiteDesc site DYN_SITE_INIT;
  return status(dynode_site_logdensity_f64(&site, (int64_t)z.element_count(), z.typed_data(), 1, x->typed_data(),
                                           lp->typed_data(), stream));
}

ffi::Error SiteVjpImpl(cudaStream_t stream, F64 z, F64 gx, F64 glp, DYN_SITE_PARAMS, F64Out gz) {
  const DynodeSiteDesc site DYN_SITE_INIT;
  return status(dynode_site_logdensity_vjp_f64(&site, (int64_t)z.element_count(), z.typed_data(), 1, gx.typed_data(),
                                               glp.typed_data(), gz->typed_data(), stream));
}

}  // namespace

#define DYN_MODEL_ATTRS \
  .Attr<int32_t>("flow").Attr<int32_t>("flags").Attr<int32_t>("n_groups").Attr<int32_t>("n_strains")
#define DYN_SOLVER_ATTRS                                                                          \
  .Attr<double>("t0").Attr<double>("t1").Attr<double>("rtol").Attr<double>("atol")                 \
      .Attr<double>("const_dt").Attr<int64_t>("max_steps").Attr<double>("save_dt")
// stream, y0, beta, gamma, sigma, omega, season, contact, save_ts, jump_ts, only
#define DYN_INPUTS                                                                                     \
  .Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>() \
      .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<U8>()
#define DYN_SITE_ATTRS                                                                                       \
  .Attr<int32_t>("bijector").Attr<int32_t>("family").Attr<double>("a").Attr<double>("b").Attr<double>("p0") \
      .Attr<double>("p1").Attr<double>("c").Attr<double>("aff_loc").Attr<double>("aff_scale")
#define S32 ffi::Buffer<ffi::S32>

XLA_FFI_DEFINE_HANDLER_SYMBOL(DynodeSolve, SolveImpl,
                              ffi::Ffi::Bind() DYN_INPUTS DYN_MODEL_ATTRS.Attr<int64_t>("save_mask")
                                  DYN_SOLVER_ATTRS.Ret<F64>().Ret<S32>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(DynodeSolveSens, SolveSensImpl,
                              ffi::Ffi::Bind() DYN_INPUTS.Arg<F64>().Attr<ffi::Span<const int32_t>>("wrt")
                                  DYN_MODEL_ATTRS.Attr<int64_t>("save_mask") DYN_SOLVER_ATTRS.Ret<F64>()
                                  .Ret<F64>().Ret<S32>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(DynodePoissonLoglikGrad, LoglikGradImpl,
                              ffi::Ffi::Bind() DYN_INPUTS.Arg<F64>().Arg<F64>()
                                  .Attr<ffi::Span<const int32_t>>("wrt") DYN_MODEL_ATTRS.Attr<int32_t>("obs_comp")
                                  .Attr<double>("lp_const") DYN_SOLVER_ATTRS.Ret<F64>().Ret<F64>().Ret<S32>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(DynodePoissonLoglikAdjoint, LoglikAdjointImpl,
                              ffi::Ffi::Bind() DYN_INPUTS.Arg<F64>() DYN_MODEL_ATTRS.Attr<int32_t>("obs_comp")
                                  .Attr<double>("lp_const").Attr<int32_t>("cap") DYN_SOLVER_ATTRS.Ret<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<S32>().Ret<F64>().Ret<F64>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(DynodeSeipSolve, SeipSolveImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Arg<F64>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Arg<F64>().Arg<U8>().Attr<int32_t>("n_ages").Attr<int32_t>("n_strains")
                                  .Attr<int32_t>("n_wane").Attr<int32_t>("n_vax").Attr<int32_t>("n_knots")
                                  .Attr<int64_t>("save_mask").Attr<double>("season_tau").Attr<double>("season_on")
                                  DYN_SOLVER_ATTRS.Ret<F64>().Ret<S32>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(DynodeSiteLogdensity, SiteImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>()
                                  DYN_SITE_ATTRS.Ret<F64>().Ret<F64>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(DynodeSiteLogdensityVjp, SiteVjpImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F64>().Arg<F64>()
                                  .Arg<F64>() DYN_SITE_ATTRS.Ret<F64>());
