// ppl_kernels.cu -- the bijector of a latent site and its log-Jacobian in one elementwise kernel each way
// (include/dynode_b200_ppl.h).  Restates numpyro's biject_to(interval / greater_than / less_than) =
// SigmoidTransform / ExpTransform followed by AffineTransform, with SigmoidTransform.log_abs_det_jacobian
// = -softplus(z) - softplus(-z) (what the reference's NUTS evaluates around every model call,
// src/dynode/infer/inference.py:149-163).
#include <cuda_runtime.h>
#include <math_constants.h>

#include "../../include/dynode_b200_ppl.h"

namespace dynode {
int fail_msg(const char* fmt, ...);  // capi.cu

namespace {

__device__ __forceinline__ double softplus(double v) {  // log(1 + e^v) without overflow
  return fmax(v, 0.0) + log1p(exp(-fabs(v)));
}
__device__ __forceinline__ double sigmoid(double v) {
  const double e = exp(-fabs(v));
  const double s = 1.0 / (1.0 + e);  // sigmoid(|v|)
  return v >= 0.0 ? s : e * s;
}

// x, dx/dz, log|dx/dz| and its derivative for one element
struct Bij { double x, dxdz, ladj, dladj; };
__device__ __forceinline__ Bij bijector(int kind, double v, double a, double b, double logb) {
  Bij r;
  if (kind == DYNODE_BIJ_INTERVAL) {
    const double s = sigmoid(v);
    r.x = fma(b, s, a);
    r.dxdz = b * s * (1.0 - s);
    r.ladj = logb - softplus(v) - softplus(-v);
    r.dladj = 1.0 - 2.0 * s;
  } else if (kind == DYNODE_BIJ_REAL) {
    r.x = v; r.dxdz = 1.0; r.ladj = 0.0; r.dladj = 0.0;
  } else {
    const double e = exp(v);
    r.x = (kind == DYNODE_BIJ_GREATER_THAN) ? a + e : a - e;
    r.dxdz = (kind == DYNODE_BIJ_GREATER_THAN) ? e : -e;
    r.ladj = v;
    r.dladj = 1.0;
  }
  return r;
}
__global__ void __launch_bounds__(256) bijector_kernel(int kind, int64_t n, const double* __restrict__ z, double a,
                                                        double b, double* __restrict__ x, double* __restrict__ ladj) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const double logb = (kind == DYNODE_BIJ_INTERVAL) ? log(fabs(b)) : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const Bij r = bijector(kind, z[i], a, b, logb);
    x[i] = r.x;
    ladj[i] = r.ladj;
  }
}

__global__ void __launch_bounds__(256) bijector_vjp_kernel(int kind, int64_t n, const double* __restrict__ z, double b,
                                                            const double* __restrict__ gx,
                                                            const double* __restrict__ gl, double* __restrict__ gz) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const Bij r = bijector(kind, z[i], 0.0, b, 0.0);
    gz[i] = fma(gx[i], r.dxdz, gl[i] * r.dladj);
  }
}

__device__ __forceinline__ double xlogy(double c, double y) { return c == 0.0 ? 0.0 : c * log(y); }
// f_family(u) and its derivative
__device__ __forceinline__ void family(int fam, double u, double p0, double p1, double& f, double& df) {
  switch (fam) {
    case DYNODE_FAM_NORMAL: { const double t = (u - p0) / p1; f = -0.5 * t * t; df = -t / p1; break; }
    case DYNODE_FAM_UNIFORM: f = 0.0; df = 0.0; break;
    case DYNODE_FAM_BETA:
      f = xlogy(p0 - 1.0, u) + xlogy(p1 - 1.0, 1.0 - u);
      df = (p0 - 1.0) / u - (p1 - 1.0) / (1.0 - u);
      break;
    case DYNODE_FAM_GAMMA: f = (p0 - 1.0) * log(u) - p1 * u; df = (p0 - 1.0) / u - p1; break;
    case DYNODE_FAM_LOGNORMAL: {
      const double lv = log(u), t = (lv - p0) / p1;
      f = -0.5 * t * t - lv; df = -(t / p1 + 1.0) / u; break;
    }
    case DYNODE_FAM_HALFNORMAL: { const double t = u / p0; f = -0.5 * t * t; df = -t / p0; break; }
    default: f = -p0 * u; df = -p0; break;  // EXPONENTIAL
  }
}

__global__ void __launch_bounds__(256) site_kernel(const DynodeSiteDesc s, int64_t n, const double* __restrict__ z,
                                                    int64_t zs, double* __restrict__ x, double* __restrict__ lp) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const double logb = (s.bijector == DYNODE_BIJ_INTERVAL) ? log(fabs(s.b)) : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const Bij r = bijector(s.bijector, z[i * zs], s.a, s.b, logb);
    double f, df;
    family(s.family, (r.x - s.aff_loc) / s.aff_scale, s.p0, s.p1, f, df);
    x[i] = r.x;
    lp[i] = r.ladj + (f + s.c);
  }
}
__global__ void __launch_bounds__(256) site_vjp_kernel(const DynodeSiteDesc s, int64_t n, const double* __restrict__ z,
                                                        int64_t zs, const double* __restrict__ gx,
                                                        const double* __restrict__ glp,
                                                        double* __restrict__ gz) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const Bij r = bijector(s.bijector, z[i * zs], s.a, s.b, 0.0);
    double f, df;
    family(s.family, (r.x - s.aff_loc) / s.aff_scale, s.p0, s.p1, f, df);
    // a site pushed to the edge of its support (x == bound in floating point) has dx/dz == 0: no 0 * inf
    const double chain = (r.dxdz == 0.0) ? 0.0 : (df / s.aff_scale) * r.dxdz;
    gz[i] = fma(gx[i], r.dxdz, glp[i] * (r.dladj + chain));
  }
}

// ---- whole-model evaluation around the ODE launch (DynodePotentialPlan) -------------------------------------
struct PlanCols { int32_t col[DYNODE_PLAN_MAX_RATES]; };

__global__ void __launch_bounds__(128) potential_pre_kernel(const DynodePotentialPlan p, int64_t C,
                                                             const double* __restrict__ z, int64_t zs,
                                                             double* __restrict__ theta, double* __restrict__ aux,
                                                             const uint8_t* __restrict__ only) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (only && only[c] == 0) return;
  const int D = p.n_sites, K = p.n_rates;
  double x[DYNODE_PLAN_MAX_SITES];
  double* a = aux + c * (3 * D + 1);
  double prior = 0.0;
#pragma unroll 1
  for (int j = 0; j < D; ++j) {
    const DynodeSiteDesc& s = p.site[j];
    const double logb = (s.bijector == DYNODE_BIJ_INTERVAL) ? log(fabs(s.b)) : 0.0;
    const Bij r = bijector(s.bijector, z[c * zs + j], s.a, s.b, logb);
    double f, df;
    family(s.family, (r.x - s.aff_loc) / s.aff_scale, s.p0, s.p1, f, df);
    const double chain = (r.dxdz == 0.0) ? 0.0 : (df / s.aff_scale) * r.dxdz;
    x[j] = r.x;
    a[3 * j + 0] = r.x;
    a[3 * j + 1] = r.dxdz;
    a[3 * j + 2] = r.dladj + chain;
    prior += r.ladj + (f + s.c);
  }
  a[3 * D] = prior;
#pragma unroll 1
  for (int k = 0; k < K; ++k) {
    double num = p.rate_c[k], den = 1.0;
    for (int j = 0; j < D; ++j) {
      const int e = p.rate_e[k][j];
      if (e > 0) num *= x[j];
      if (e < 0) den *= x[j];
    }
    theta[c * K + k] = num / den;
  }
}

__global__ void __launch_bounds__(128) potential_post_kernel(
    const DynodePotentialPlan p, int64_t C, const double* __restrict__ theta, const double* __restrict__ aux,
    const double* __restrict__ lp, const double* __restrict__ grad, int64_t gs, const PlanCols gc,
    const double* __restrict__ lp_fb, const double* __restrict__ grad_fb, int64_t gs_fb, const PlanCols gc_fb,
    const int32_t* __restrict__ stats, const uint8_t* __restrict__ only, double* __restrict__ U,
    double* __restrict__ dU) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int D = p.n_sites, K = p.n_rates;
  if (only && only[c] == 0) {
    U[c] = 0.0;
    for (int j = 0; j < D; ++j) dU[c * D + j] = 0.0;
    return;
  }
  const bool fb = stats && lp_fb && stats[c * 4 + 0] == 2 /* DYNODE_RESULT_ADJOINT_CAPACITY */;
  const double* g = fb ? grad_fb + c * gs_fb : grad + c * gs;
  const int32_t* col = fb ? gc_fb.col : gc.col;
  const double* a = aux + c * (3 * D + 1);
  U[c] = -(a[3 * D] + (fb ? lp_fb[c] : lp[c]));
  double acc[DYNODE_PLAN_MAX_SITES];
  for (int j = 0; j < D; ++j) acc[j] = 0.0;
#pragma unroll 1
  for (int k = 0; k < K; ++k) {
    if (col[k] < 0) continue;
    const double w = g[col[k]] * theta[c * K + k];  // d lp / d log theta_k
    for (int j = 0; j < D; ++j) {
      const int e = p.rate_e[k][j];
      if (e != 0) acc[j] += (e > 0 ? w : -w) / a[3 * j];  // theta_k * e_kj / x_j
    }
  }
  for (int j = 0; j < D; ++j) dU[c * D + j] = -fma(acc[j], a[3 * j + 1], a[3 * j + 2]);
}

int grid_for(int64_t n) {
  const int64_t g = (n + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

}  // namespace
}  // namespace dynode

using namespace dynode;

extern "C" {

int dynode_bijector_f64(int32_t kind, int64_t n, const double* z, double a, double b, double* x, double* ladj,
                        void* stream) {
  if (kind < 0 || kind > DYNODE_BIJ_REAL) return fail_msg("unknown bijector kind %d", kind);
  if (n < 0 || (n > 0 && (!z || !x || !ladj))) return fail_msg("bijector: null buffer");
  if (n == 0) return 0;
  bijector_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(kind, n, z, a, b, x, ladj);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail_msg("bijector launch failed: %s", cudaGetErrorString(e));
}

int dynode_bijector_vjp_f64(int32_t kind, int64_t n, const double* z, double b, const double* gx, const double* gl,
                            double* gz, void* stream) {
  if (kind < 0 || kind > DYNODE_BIJ_REAL) return fail_msg("unknown bijector kind %d", kind);
  if (n < 0 || (n > 0 && (!z || !gx || !gl || !gz))) return fail_msg("bijector vjp: null buffer");
  if (n == 0) return 0;
  bijector_vjp_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(kind, n, z, b, gx, gl, gz);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail_msg("bijector vjp launch failed: %s", cudaGetErrorString(e));
}

static int check_site(const DynodeSiteDesc* s) {
  if (!s) return fail_msg("null site descriptor");
  if (s->bijector < 0 || s->bijector > DYNODE_BIJ_REAL) return fail_msg("unknown bijector kind %d", s->bijector);
  if (s->family < 0 || s->family > DYNODE_FAM_EXPONENTIAL) return fail_msg("unknown prior family %d", s->family);
  if (!(s->aff_scale != 0.0)) return fail_msg("site: aff_scale must be non-zero");
  return 0;
}

int dynode_site_logdensity_f64(const DynodeSiteDesc* site, int64_t n, const double* z, int64_t z_stride, double* x,
                               double* lp, void* stream) {
  if (int rc = check_site(site)) return rc;
  if (n < 0 || (n > 0 && (!z || !x || !lp))) return fail_msg("site: null buffer");
  if (n == 0) return 0;
  site_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(*site, n, z, z_stride, x, lp);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail_msg("site launch failed: %s", cudaGetErrorString(e));
}

int dynode_site_logdensity_vjp_f64(const DynodeSiteDesc* site, int64_t n, const double* z, int64_t z_stride,
                                   const double* gx, const double* glp, double* gz, void* stream) {
  if (int rc = check_site(site)) return rc;
  if (n < 0 || (n > 0 && (!z || !gx || !glp || !gz))) return fail_msg("site vjp: null buffer");
  if (n == 0) return 0;
  site_vjp_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(*site, n, z, z_stride, gx, glp, gz);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail_msg("site vjp launch failed: %s", cudaGetErrorString(e));
}

static int check_plan(const DynodePotentialPlan* p) {
  if (!p) return fail_msg("null potential plan");
  if (p->n_sites < 1 || p->n_sites > DYNODE_PLAN_MAX_SITES) return fail_msg("plan: n_sites must be in [1, %d]", DYNODE_PLAN_MAX_SITES);
  if (p->n_rates < 1 || p->n_rates > DYNODE_PLAN_MAX_RATES) return fail_msg("plan: n_rates must be in [1, %d]", DYNODE_PLAN_MAX_RATES);
  for (int j = 0; j < p->n_sites; ++j)
    if (int rc = check_site(&p->site[j])) return rc;
  for (int k = 0; k < p->n_rates; ++k)
    for (int j = 0; j < p->n_sites; ++j)
      if (p->rate_e[k][j] < -1 || p->rate_e[k][j] > 1) return fail_msg("plan: rate exponents must be -1, 0 or 1");
  return 0;
}

int dynode_potential_pre_f64(const DynodePotentialPlan* plan, int64_t C, const double* z, int64_t z_stride,
                             double* theta, double* aux, const uint8_t* only, void* stream) {
  if (int rc = check_plan(plan)) return rc;
  if (C < 0 || (C > 0 && (!z || !theta || !aux))) return fail_msg("potential pre: null buffer");
  if (z_stride < plan->n_sites) return fail_msg("potential pre: z_stride smaller than the number of sites");
  if (C == 0) return 0;
  potential_pre_kernel<<<(unsigned)((C + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*plan, C, z, z_stride, theta,
                                                                                       aux, only);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail_msg("potential pre launch failed: %s", cudaGetErrorString(e));
}

int dynode_potential_post_f64(const DynodePotentialPlan* plan, int64_t C, const double* theta, const double* aux,
                              const double* lp, const double* grad, int64_t grad_stride, const int32_t* grad_col,
                              const double* lp_fb, const double* grad_fb, int64_t grad_fb_stride,
                              const int32_t* grad_fb_col, const int32_t* stats, const uint8_t* only, double* U,
                              double* dU, void* stream) {
  if (int rc = check_plan(plan)) return rc;
  if (C < 0 || (C > 0 && (!theta || !aux || !lp || !grad || !grad_col || !U || !dU)))
    return fail_msg("potential post: null buffer");
  if ((lp_fb != nullptr) != (grad_fb != nullptr) || (lp_fb && (!grad_fb_col || !stats)))
    return fail_msg("potential post: the fallback needs lp, grad, its column map and stats together");
  if (C == 0) return 0;
  PlanCols gc, gf;
  for (int k = 0; k < DYNODE_PLAN_MAX_RATES; ++k) {
    gc.col[k] = k < plan->n_rates ? grad_col[k] : -1;
    gf.col[k] = (lp_fb && k < plan->n_rates) ? grad_fb_col[k] : -1;
    if (gc.col[k] >= grad_stride || (lp_fb && gf.col[k] >= grad_fb_stride))
      return fail_msg("potential post: gradient column outside the row");
  }
  potential_post_kernel<<<(unsigned)((C + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      *plan, C, theta, aux, lp, grad, grad_stride, gc, lp_fb, grad_fb, grad_fb_stride, gf, stats, only, U, dU);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail_msg("potential post launch failed: %s", cudaGetErrorString(e));
}

}  // extern "C"
