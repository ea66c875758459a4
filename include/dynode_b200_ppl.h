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
 *
 * All pointers are DEVICE pointers to n doubles; launches are enqueued on `stream`, allocate nothing.
 */
#ifndef DYNODE_B200_PPL_H_
#define DYNODE_B200_PPL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { DYNODE_BIJ_INTERVAL = 0, DYNODE_BIJ_GREATER_THAN = 1, DYNODE_BIJ_LESS_THAN = 2 };

/* x[i], ladj[i] from z[i] */
int dynode_bijector_f64(int32_t kind, int64_t n, const double* z, double a, double b, double* x, double* ladj,
                        void* stream);
/* gz[i] = gx[i] * dx/dz + gl[i] * d ladj/dz */
int dynode_bijector_vjp_f64(int32_t kind, int64_t n, const double* z, double b, const double* gx, const double* gl,
                            double* gz, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DYNODE_B200_PPL_H_ */
