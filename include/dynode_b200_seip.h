/* dynode_b200_seip.h -- C ABI of the CTA-per-trajectory kernel for the immune-history / waning family.
 *
 * The model is the SEIP structure of the reference's prose document (reference ode_model.md:15-53: S indexed
 * age x immune history x waning stage, E/I/C indexed age x immune history x strain; ode_model.md:100-118: the
 * immune-history update eta(x, y) = x | 2^y; ode_model.md:179-211: force of infection reduced by immunity).
 * The reference ships NO implementation of it (moved to a private repository, CHANGELOG:120-122), so the exact
 * equations are this repository's reading, stated in oracle/dynode_oracle.cpp (FAM_SEIP) and repeated here;
 * the vaccination dimension of the prose model is not included.
 *
 *   H = 2^K immune histories (bit k set = recovered from strain k at least once)
 *   state row (n = A*H*W + 3*A*H*K doubles):  S[A][H][W], E[A][H][K], I[A][H][K], C[A][H][K]
 *   foi[a][k]   = beta_k * sum_b contact[a][b] * (sum_j I[b][j][k]) / pop[b]
 *   x[a][j][w][k] = foi[a][k] * (1 - immunity[j][w][k]) * S[a][j][w]
 *   dS[a][j][w] = -sum_k x[a][j][w][k] + omega[w-1] S[a][j][w-1] - omega[w] S[a][j][w]          (waning chain)
 *                 + [w == 0] sum_{k in j} gamma_k (I[a][j][k] + I[a][j \ k][k])                 (recovery, eta)
 *   dE[a][j][k] = sum_w x[a][j][w][k] - sigma_k E;  dI = sigma_k E - gamma_k I;  dC = sum_w x[a][j][w][k]
 *
 * Integration is the same diffeqsolve restatement as dynode_solve_f64 (Tsit5, I-controller, Hairer initial step,
 * SaveAt by dense output); one thread block integrates one trajectory with the state and the 7 stage derivatives
 * staged in shared memory, block-wide reductions for the infectious totals and the RMS error norm.
 */
#ifndef DYNODE_B200_SEIP_H_
#define DYNODE_B200_SEIP_H_

#include <stdint.h>

#include "dynode_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define DYNODE_SEIP_MAX_STRAINS 4
#define DYNODE_SEIP_MAX_STATE 1536 /* 9 * n doubles must fit the 227 KB of shared memory of one CTA */

typedef struct {
  int32_t n_ages;    /* A */
  int32_t n_strains; /* K <= 4, H = 2^K */
  int32_t n_wane;    /* W */
} DynodeSeipDesc;

typedef struct {
  DynodeArray beta, sigma, gamma; /* [B][K] */
  DynodeArray omega;              /* [B][W]  waning rates; omega[W-1] is ignored (last stage absorbs) */
  const double* contact;          /* [A][A] shared, contact[target][source] */
  const double* pop;              /* [A]    shared, population per age group */
  const double* immunity;         /* [H][W][K] shared, protection in [0, 1] */
} DynodeSeipParams;

int dynode_seip_state_size(const DynodeSeipDesc* model);

/* y0 [B][n] (batch_stride 0 = shared), save_ts [T], ys [B][T][n] (unreached slots +inf), stats [B][4]. */
int dynode_seip_solve_f64(const DynodeSeipDesc* model, const DynodeSolverDesc* solver, int64_t B, DynodeArray y0,
                          const DynodeSeipParams* params, const double* save_ts, int32_t T, double* ys,
                          int32_t* stats, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DYNODE_B200_SEIP_H_ */
