/* dynode_b200_seip.h -- C ABI of the CTA-per-trajectory kernel for the immune-history / waning family.
 *
 * The model is the SEIP structure of the reference's prose document (reference ode_model.md:15-53: S indexed
 * age x immune history x waning stage, E/I/C indexed age x immune history x strain; ode_model.md:100-118: the
 * immune-history update eta(x, y) = x | 2^y; ode_model.md:179-211: force of infection reduced by immunity).
 * The reference ships NO implementation of it (moved to a private repository, CHANGELOG:120-122), so the exact
 * equations are this repository's reading, stated in oracle/dynode_oracle.cpp (FAM_SEIP) and repeated here;
 * the vaccination dimension, the spline vaccination rates, the seasonal tier reset and external introductions are
 * the FAM_SEIPV extension (terms marked [V] below; absent with n_vax <= 1 / NULL tables).
 *
 *   H = 2^K immune histories (bit k set = recovered from strain k at least once)
 *   state row (n = A*H*W + 3*A*H*K doubles):  S[A][H][W], E[A][H][K], I[A][H][K], C[A][H][K]
 *   foi[a][k]   = beta_k * sum_b contact[a][b] * (sum_j I[b][j][k]) / pop[b]
 *   x[a][j][w][k] = foi[a][k] * (1 - immunity[j][w][k]) * S[a][j][w]
 *   dS[a][j][w] = -sum_k x[a][j][w][k] + omega[w-1] S[a][j][w-1] - omega[w] S[a][j][w]          (waning chain)
 *                 + [w == 0] sum_{k in j} gamma_k (I[a][j][k] + I[a][j \ k][k])                 (recovery, eta)
 *   dE[a][j][k] = sum_w x[a][j][w][k] - sigma_k E;  dI = sigma_k E - gamma_k I;  dC = sum_w x[a][j][w][k]
 * [V] every cell carries a tier v: S[A][H][V][W], E/I/C[A][H][V][K], n = A*H*V*(W + 3K), immunity[j][v][w][k];
 *     infectious fraction += N(t; intro_time_k, intro_scale_k) intro_pct_k intro_ages[k][b]
 *     r[a][v] = min(nu[a][v](t) pop[a] / sum_{j,w} S[a][j][v][w], 1);  S[a][j][v][w] --r--> S[a][j][v+1][0];
 *     in the top tier a dose moves stages w >= 1 back to stage 0;  phi(t) moves the top tier of S, E, I one tier down
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
  int32_t n_vax;     /* V vaccination tiers (0 or 1 = none); the top tier V-1 is the prose model's K */
  int32_t n_knots;   /* knots of the vaccination-rate splines */
  uint32_t save_mask; /* bit c set = compartment c of (S, E, I, C) is saved (sub_save_indices, odes.py:182-193);
                         0 = all four.  ys rows then hold the saved compartments only, concatenated */
} DynodeSeipDesc;

typedef struct {
  DynodeArray beta, sigma, gamma; /* [B][K] */
  DynodeArray omega;              /* [B][W]  waning rates; omega[W-1] is ignored (last stage absorbs) */
  const double* contact;          /* [A][A] shared, contact[target][source] */
  const double* pop;              /* [A]    shared, population per age group */
  const double* immunity;         /* [H][V][W][K] shared, protection in [0, 1] */
  /* vaccination (reference ode_model.md:19-29, utils/splines.py:72-109): nu[a][v](t) = max(0, base . (1, t, t^2, t^3)
   * + sum_i coef_i max(t - knot_i, 0)^3), the proportion of age group a vaccinated per day out of tier v.
   * vax_base NULL = no vaccination. */
  const double* vax_base;  /* [A][V][4] shared */
  const double* vax_knots; /* [A][V][n_knots] */
  const double* vax_coef;  /* [A][V][n_knots] */
  /* external introductions (ode_model.md:183, config/strains.py:59-109): strain k is carried in by an untracked
   * population of relative size intro_pct[k], Gaussian in time around intro_time[k] with standard deviation
   * intro_scale[k] days, age structure intro_ages[k][a].  intro_pct.ptr NULL = none. */
  DynodeArray intro_time, intro_scale, intro_pct; /* [B][K] */
  const double* intro_ages;                       /* [K][A] shared */
  /* seasonal reset of the top tier (ode_model.md:72-75): phi(t) = season_on sin(2 pi (t + season_tau) / 730)^1000 */
  double season_tau, season_on;
} DynodeSeipParams;

int dynode_seip_state_size(const DynodeSeipDesc* model);

/* y0 [B][n] (batch_stride 0 = shared), save_ts [T], ys [B][T][n_saved] (unreached slots +inf), stats [B][4].
 * DynodeSolverDesc.jump_ts (SolverParams.discontinuity_points) is honoured as in dynode_solve_f64. */
int dynode_seip_solve_f64(const DynodeSeipDesc* model, const DynodeSolverDesc* solver, int64_t B, DynodeArray y0,
                          const DynodeSeipParams* params, const double* save_ts, int32_t T, double* ys,
                          int32_t* stats, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DYNODE_B200_SEIP_H_ */
