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

#ifdef __cplusplus
}
#endif
#endif /* DYNODE_B200_PPL_H_ */
