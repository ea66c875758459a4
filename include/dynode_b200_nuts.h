/* dynode_b200_nuts.h -- C ABI of the many-chain NUTS bookkeeping kernels.
 *
 * The reference runs numpyro's NUTS (MCMC(NUTS(model, dense_mass=True, max_tree_depth, init_to_median)),
 * reference src/dynode/infer/inference.py:149-163) one chain after another on the CPU.  Here every chain is
 * one CUDA thread: a *round* advances all chains by one leapfrog step and is
 *
 *     dynode_nuts_round_pre   (new momentum / new doubling where due, first half of the leapfrog)
 *     potential_and_grad      (the model; its ODE part is dynode_poisson_loglik_grad_f64 / dynode_solve_sens_f64)
 *     dynode_nuts_round_post  (second half of the leapfrog, multinomial sampling inside the subtree,
 *                              checkpointed U-turn tests, tree doubling, and -- when a chain's tree is
 *                              complete -- the transition commit: dual-averaging step size, Welford
 *                              covariance, window-end mass-matrix update (per-thread Cholesky), storage of
 *                              the draw)
 *
 * Both launches are stream-ordered, allocate nothing and keep no state outside the buffers the caller owns,
 * so a whole round (model included) can be captured in a CUDA graph.  All pointers are DEVICE pointers.
 * Layouts are row-major with the chain index first: [C], [C][D], [C][D][D], [C][max_depth][D], [C][N][D].
 */
#ifndef DYNODE_B200_NUTS_H_
#define DYNODE_B200_NUTS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DYNODE_NUTS_MAX_DIM 16
#define DYNODE_NUTS_MAX_DEPTH 12

/* Per-transition schedule flags (st->sched[t] for the chain's t-th transition, warm-up first).  The adaptation
 * state is per chain (numpyro semantics), so chains never wait for each other: every chain walks the schedule
 * at its own pace and does its own window-end updates inside dynode_nuts_round_post. */
#define DYNODE_NUTS_ADAPT 1u      /* dual averaging of the log step size after this transition */
#define DYNODE_NUTS_WELFORD 2u    /* the position enters the Welford covariance (slow windows) */
#define DYNODE_NUTS_SAMPLING 4u   /* the draw is stored at out_*[t - n_warmup] */
#define DYNODE_NUTS_END_SLOW 8u   /* last transition of a slow window: inverse mass matrix <- shrunk covariance
                                     (if WELFORD); step-size search from the current step size, then dual averaging
                                     restarted around 10 x the result (if ADAPT) -- numpyro _update_at_window_end */
#define DYNODE_NUTS_END_WARMUP 16u /* last warm-up transition: step size <- averaged iterate (if ADAPT) */

typedef struct {
  int32_t C, D, max_depth, N; /* chains, dimension, numpyro max_tree_depth, draws kept per chain */
  int32_t n_warmup;           /* transitions before the first stored draw */
  int32_t dense;              /* 1: dense inverse mass matrix, 0: diagonal */
  double target_accept;       /* dual averaging target (0.8) */
  /* chain state */
  double *z, *U, *g;    /* [C][D], [C], [C][D]: position, potential energy, its gradient */
  double *eps;          /* [C] step size */
  double *imm, *msqrt;  /* [C][D][D] inverse mass matrix; factor with momentum = msqrt @ N(0, I) */
  /* schedule bookkeeping (chains run asynchronously through the whole warm-up + sampling schedule) */
  int64_t *k;           /* [C] transitions completed */
  const int64_t *nwin;  /* [1] transitions each chain has to make in total */
  uint8_t *active, *need_tree; /* [C] */
  /* numpyro's find_reasonable_step_size (run at the start of warm-up and at the end of every slow window): while
   * `searching`, a round is one probe -- fresh momentum, ONE leapfrog from the current state with eps * 2^fr_dir --
   * and the step size keeps doubling / halving until the direction "accept probability above 0.8?" flips. */
  uint8_t *searching;        /* [C] */
  int64_t *fr_dir, *fr_last; /* [C] direction of the next probe / of the one before (-1, 0, +1) */
  const uint8_t *sched;   /* [*nwin] DYNODE_NUTS_* flags per transition */
  const double *sched_n;  /* [*nwin] length of the adaptation window the transition belongs to */
  /* whole tree */
  double *energy0;                  /* [C] */
  double *zL, *rL, *gL, *zR, *rR, *gR, *zP, *gP, *r_sum; /* [C][D] */
  double *UP, *weight, *sum_acc;    /* [C] */
  int64_t *depth, *nprop;           /* [C] */
  uint8_t *turning, *diverging;     /* [C] */
  /* subtree under construction */
  int64_t *s_n;                     /* [C] leaves so far */
  uint8_t *s_right, *s_turn, *s_div; /* [C] */
  double *s_z, *s_r, *s_g, *s_zP, *s_gP, *s_rsum; /* [C][D] */
  double *s_UP, *s_w, *s_acc;       /* [C] */
  double *r_ck, *rs_ck;             /* [C][max_depth][D] U-turn checkpoints */
  /* leapfrog scratch written by _pre, read by _post */
  double *z_new, *r_half;           /* [C][D] */
  /* adaptation */
  double *da_x, *da_xavg, *da_gavg, *da_t, *da_prox; /* [C] dual averaging of log step size */
  double *wf_n, *wf_mean, *wf_m2;   /* [C], [C][D], [C][D][D] Welford accumulators */
  /* outputs */
  double *out_z;                    /* [C][N][D] */
  double *out_accept, *out_steps, *out_div, *out_energy, *out_depth; /* [C][N] */
  double *last_accept, *last_steps; /* [C] */
  int64_t *n_leap;                  /* [C] leapfrogs that belonged to a tree */
  uint8_t *any_active;              /* [1] cleared by _pre, set by _post when a chain is still running afterwards */
} DynodeNutsState;

/* rnd_n [C][D] standard normals (fresh momentum), rnd_u [C][3] uniforms (direction, subtree transition,
 * top-level transition).  Returns 0 when enqueued, nonzero + dynode_last_error() otherwise. */
int dynode_nuts_round_pre(const DynodeNutsState* st, const double* rnd_n, const double* rnd_u, void* stream);

/* U_new [C], g_new [C][D]: potential energy and gradient at st->z_new (non-finite values mark a divergent leaf). */
int dynode_nuts_round_post(const DynodeNutsState* st, const double* U_new, const double* g_new,
                           const double* rnd_u, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DYNODE_B200_NUTS_H_ */
