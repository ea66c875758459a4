// Test driver (TEST INFRASTRUCTURE): runs the BODIES of the XLA-FFI handlers of dynode_b200/csrc/xla_ffi_shim.cc on
// real device memory, with tests/mock_xla's Buffer standing in for XLA's.  What it proves that compiling alone does
// not: the shim's own plumbing -- batch-stride detection, the season / contact / jump_ts / row-mask conventions,
// attribute order -- hands the C ABI exactly what the ctypes binding hands it (tests/test_gpu_parity.py compares the
// two on a B200).  It says nothing about XLA's runtime.
#include "../../dynode_b200/csrc/xla_ffi_shim.cc"

#include <cstring>
#undef S32

namespace {
F64 f64(const double* p, int64_t d0, int64_t d1 = -1, int64_t d2 = -1) {
  std::vector<int64_t> dims;
  if (d0 >= 0) dims.push_back(d0);
  if (d1 >= 0) dims.push_back(d1);
  if (d2 >= 0) dims.push_back(d2);
  if (!p) dims.assign(1, 0);
  return F64(const_cast<double*>(p), dims);
}
thread_local std::string g_msg;
int finish(const ffi::Error& e) {
  g_msg = e.message();
  return e.failure() ? 1 : 0;
}
}  // namespace

extern "C" {

const char* shim_driver_last_error() { return g_msg.c_str(); }

// DynodeSolve's body: [y0_rows][n] y0 (1 row = shared), [B][S] rates, optional [B][3] season, [G][G] contact, [T] grid, [J] jumps, [B] mask
int shim_driver_solve(void* stream, int64_t B, int64_t y0_rows, int64_t n, int64_t S, int64_t G, int64_t T, int64_t n_saved,
                      const double* y0, const double* beta, const double* gamma, const double* sigma,
                      const double* omega, const double* season, const double* contact, const double* save_ts,
                      const double* jump_ts, int64_t n_jump, const uint8_t* only, int32_t flow, int32_t flags,
                      int64_t save_mask, double t1, double rtol, double atol, double const_dt, int64_t max_steps,
                      double save_dt, double* ys, int32_t* stats) {
  U8 mask(const_cast<uint8_t*>(only), std::vector<int64_t>{only ? B : 0});
  F64Out ys_r(F64(ys, {B, T, n_saved}));
  S32Out st_r(ffi::Buffer<ffi::S32>(stats, {B, 4}));
  return finish(SolveImpl((cudaStream_t)stream, f64(y0, y0_rows, n), f64(beta, B, S), f64(gamma, B, S), f64(sigma, B, S),
                          f64(omega, B, S), f64(season, B, 3), f64(contact, G, G), f64(save_ts, T),
                          f64(jump_ts, n_jump), mask, flow, flags, (int32_t)G, (int32_t)S, save_mask, 0.0, t1, rtol,
                          atol, const_dt, max_steps, save_dt, ys_r, st_r));
}

// DynodePoissonLoglikGrad's body
int shim_driver_loglik_grad(void* stream, int64_t B, int64_t y0_rows, int64_t n, int64_t S, int64_t G, int64_t T, int64_t m,
                            const double* y0, const double* beta, const double* gamma, const double* sigma,
                            const double* omega, const double* contact, const double* save_ts, const double* obs,
                            const int32_t* wrt, int64_t n_wrt, int32_t flow, int32_t flags, int32_t obs_comp,
                            double lp_const, double t1, double rtol, double atol, int64_t max_steps, double save_dt,
                            double* lp, double* grad, int32_t* stats) {
  U8 mask(nullptr, std::vector<int64_t>{0});
  F64Out lp_r(F64(lp, {B})), g_r(F64(grad, {B, n_wrt}));
  S32Out st_r(ffi::Buffer<ffi::S32>(stats, {B, 4}));
  return finish(LoglikGradImpl((cudaStream_t)stream, f64(y0, y0_rows, n), f64(beta, B, S), f64(gamma, B, S),
                               f64(sigma, B, S), f64(omega, B, S), f64(nullptr, 0), f64(contact, G, G),
                               f64(save_ts, T), f64(nullptr, 0), mask, f64(obs, T - 1, m), f64(nullptr, 0),
                               ffi::Span<const int32_t>(wrt, (size_t)n_wrt), flow, flags, (int32_t)G, (int32_t)S,
                               obs_comp, lp_const, 0.0, t1, rtol, atol, 0.0, max_steps, save_dt, lp_r, g_r, st_r));
}

// DynodeSeipSolve's body: one shared y0 [n], per-draw rates [B][K] / [B][W] and introduction parameters [B][K], shared
// tables; NULL pointers become zero-element buffers (= absent), which is the handler's own convention
int shim_driver_seip(void* stream, int64_t B, int64_t n, int64_t A, int64_t K, int64_t W, int64_t V, int64_t NK,
                     int64_t T, int64_t n_saved, const double* y0, const double* beta, const double* sigma,
                     const double* gamma, const double* omega, const double* contact, const double* pop,
                     const double* imm, const double* vbase, const double* vknots, const double* vcoef,
                     const double* itime, const double* iscale, const double* ipct, const double* iages,
                     const double* save_ts, const double* jump_ts, int64_t n_jump, int64_t save_mask,
                     double season_tau, double season_on, double t1, double rtol, double atol, double const_dt,
                     int64_t max_steps, double save_dt, double* ys, int32_t* stats) {
  const int64_t H = int64_t(1) << K;
  U8 mask(nullptr, std::vector<int64_t>{0});
  F64Out ys_r(F64(ys, {B, T, n_saved}));
  S32Out st_r(ffi::Buffer<ffi::S32>(stats, {B, 4}));
  return finish(SeipSolveImpl((cudaStream_t)stream, f64(y0, n), f64(beta, B, K), f64(sigma, B, K), f64(gamma, B, K),
                              f64(omega, B, W), f64(contact, A, A), f64(pop, A), f64(imm, H * V, W, K),
                              f64(vbase, A, V, 4), f64(vknots, A, V, NK), f64(vcoef, A, V, NK), f64(itime, B, K),
                              f64(iscale, B, K), f64(ipct, B, K), f64(iages, K, A), f64(save_ts, T),
                              f64(jump_ts, n_jump), mask, (int32_t)A, (int32_t)K, (int32_t)W, (int32_t)V,
                              (int32_t)NK, save_mask, season_tau, season_on, 0.0, t1, rtol, atol, const_dt, max_steps,
                              save_dt, ys_r, st_r));
}

}  // extern "C"
