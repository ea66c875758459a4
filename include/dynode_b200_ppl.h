/* dynode_b200_ppl.h -- fused elementwise pieces of the log-density around the ODE launch.
 *
 * numpyro evaluates a model's potential energy in unconstrained space: every latent site z is mapped to its
 * support by biject_to(support) and log|dx/dz| is added (numpyro.infer.util.potential_energy, which the
 * reference reaches through MCMC(NUTS(model)), src/dynode/infer/inference.py:149-163).  Written with tensor
 * operations that is ~13 tiny kernels forward and ~12 backward per site -- at a few thousand chains the NUTS round
 * is nothing but such launches.  One kernel each way here.
 *
 *   kind 0  interval      x = a + b * sigmoid(z)      log|dx/dz| = log b - softplus(z) - softplus(-z)
 *   kind 1  greater_than  x = a + exp(z)              log|dx/dz| = z            (b unused; positive: a = 0)
 *   kind 2  less_than     x = a - exp(z)              log|dx/dz| = z
 *   kind 3  real          x = z                       log|dx/dz| = 0
 *
 * All pointers are DEVICE pointers to n doubles; launches are enqueued on `stream`, allocate nothing.
 */
#ifndef DYNODE_B200_PPL_H_
#define DYNODE_B200_PPL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { DYNODE_BIJ_INTERVAL = 0, DYNODE_BIJ_GREATER_THAN = 1, DYNODE_BIJ_LESS_THAN = 2, DYNODE_BIJ_REAL = 3 };

/* x[i], ladj[i] from z[i] */
int dynode_bijector_f64(int32_t kind, int64_t n, const double* z, double a, double b, double* x, double* ladj,
                        void* stream);
/* gz[i] = gx[i] * dx/dz + gl[i] * d ladj/dz */
int dynode_bijector_vjp_f64(int32_t kind, int64_t n, const double* z, double b, const double* gx, const double* gl,
                            double* gz, void* stream);

/* A latent site with a prior of constant parameters, whole: x = bijector(z) and
 *   lp = log|dx/dz| + log p(x),   log p(x) = f_family((x - aff_loc) / aff_scale; p0, p1) + c
 * where c collects every term that does not depend on x (normalisers, -log|aff_scale|, the truncation mass of
 * a TruncatedNormal), computed once on the host.  aff_loc / aff_scale describe a TransformedDistribution(base,
 * AffineTransform(loc, scale)); (0, 1) otherwise.
 *   NORMAL      f = -t^2/2, t = (u - p0)/p1          (TruncatedNormal: same f, c carries the truncation)
 *   UNIFORM     f = 0
 *   BETA        f = xlogy(p0 - 1, u) + xlogy(p1 - 1, 1 - u)
 *   GAMMA       f = (p0 - 1) log u - p1 u
 *   LOGNORMAL   f = -t^2/2 - log u, t = (log u - p0)/p1
 *   HALFNORMAL  f = -(u/p0)^2/2
 *   EXPONENTIAL f = -p0 u                                                                            */
enum { DYNODE_FAM_NORMAL = 0, DYNODE_FAM_UNIFORM = 1, DYNODE_FAM_BETA = 2, DYNODE_FAM_GAMMA = 3,
       DYNODE_FAM_LOGNORMAL = 4, DYNODE_FAM_HALFNORMAL = 5, DYNODE_FAM_EXPONENTIAL = 6 };

typedef struct {
  int32_t bijector; /* DYNODE_BIJ_* */
  int32_t family;   /* DYNODE_FAM_* */
  double a, b;      /* bijector constants */
  double p0, p1, c; /* family parameters, constant term */
  double aff_loc, aff_scale;
} DynodeSiteDesc;

/* x[i], lp[i] from z[i * z_stride]  (a site is usually a column of the sampler's [chains][D] position array) */
int dynode_site_logdensity_f64(const DynodeSiteDesc* site, int64_t n, const double* z, int64_t z_stride, double* x,
                               double* lp, void* stream);
/* gz[i] = gx[i] * dx/dz + glp[i] * d lp/dz */
int dynode_site_logdensity_vjp_f64(const DynodeSiteDesc* site, int64_t n, const double* z, int64_t z_stride,
                                   const double* gx, const double* glp, double* gz, void* stream);

/* ---- a whole model evaluation around ONE ODE launch --------------------------------------------------------
 * numpyro's potential energy of a DynODE model (reference src/dynode/infer/inference.py:149-163 ->
 * examples/sir_infer_parameters.py:21-59) is
 *     U(z) = -( sum_j [ log p_j(x_j) + log|dx_j/dz_j| ]  +  log L(obs | theta(x)) ),   x_j = bijector_j(z_j)
 * with the kernel's rates a MONOMIAL map of the constrained sites, theta_k = c_k * prod_j x_j^e_kj, e in {-1,0,1}
 * (every get_odeparams of the reference: beta = r0 / infectious_period, gamma = 1 / infectious_period,
 * sigma = 1 / latent_period, omega = 1 / waning_period; examples/seirs_multi_strain_age_stratified.py:187-209).
 * Evaluated through tensor operations that is ~35 launches per gradient evaluation; here it is three:
 *     dynode_potential_pre_f64   z -> theta, and per site x, dx/dz, d(lp_j)/dz, plus the summed prior term
 *     dynode_poisson_loglik_*    theta -> log L, d log L / d theta        (include/dynode_b200.h)
 *     dynode_potential_post_f64  chain rule back to z: U, dU/dz
 * Rows with only[c] == 0 (the sampler's finished chains) are skipped by all three and come back as zeros.      */
#define DYNODE_PLAN_MAX_SITES 16
#define DYNODE_PLAN_MAX_RATES 32

typedef struct {
  int32_t n_sites; /* scalar latent sites = columns of z */
  int32_t n_rates; /* columns of theta */
  DynodeSiteDesc site[DYNODE_PLAN_MAX_SITES];
  double rate_c[DYNODE_PLAN_MAX_RATES];
  int8_t rate_e[DYNODE_PLAN_MAX_RATES][DYNODE_PLAN_MAX_SITES];
} DynodePotentialPlan;

/* aux row (3 * n_sites + 1 doubles): x_j, dx_j/dz_j, d(log p_j + ladj_j)/dz_j for every site, then the prior sum */
int dynode_potential_pre_f64(const DynodePotentialPlan* plan, int64_t C, const double* z, int64_t z_stride,
                             double* theta, double* aux, const uint8_t* only, void* stream);

/* lp [C], grad [C][grad_stride] with theta column k's derivative at grad_col[k] (-1: not differentiated).
 * A second (lp, grad, col map) triple may be given together with stats [C][4]: rows whose stats result code is
 * DYNODE_RESULT_ADJOINT_CAPACITY take it instead (the forward-sensitivity re-evaluation of rows that overflowed the
 * adjoint's checkpoint scratch).  Writes U [C] and dU [C][n_sites].                                             */
int dynode_potential_post_f64(const DynodePotentialPlan* plan, int64_t C, const double* theta, const double* aux,
                              const double* lp, const double* grad, int64_t grad_stride, const int32_t* grad_col,
                              const double* lp_fb, const double* grad_fb, int64_t grad_fb_stride,
                              const int32_t* grad_fb_col, const int32_t* stats, const uint8_t* only, double* U,
                              double* dU, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DYNODE_B200_PPL_H_ */
