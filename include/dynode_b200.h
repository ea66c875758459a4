/* dynode_b200.h -- C ABI of the B200-native ensemble ODE engine for DynODE's hot path.
 *
 * The reference has no FFI for this path: the boundary is the Python function
 *   dynode.simulation.simulate(ode, duration_days, initial_state, ode_parameters,
 *                              solver_parameters, sub_save_indices=None, save_step=1)
 * (reference src/dynode/simulation/odes.py:35-145) which forwards to diffrax.diffeqsolve
 * (odes.py:133-144).  The entry points below are what an FFI for that call binds: one batched
 * launch per diffeqsolve call (jax.vmap over parameter draws), one for its forward
 * sensitivities, one for the fused log-density + gradient NUTS asks for
 * (examples/sir_infer_parameters.py:21-39).  Plain C, plain pointers and sizes, no torch types.
 *
 * Contract for every dynode_*_f64 launch function:
 *   - all array pointers are DEVICE pointers unless marked HOST; the caller owns every buffer;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden synchronisation,
 *     no allocation, no global mutable state -> thread-safe and re-entrant;
 *   - returns 0 when the launch was enqueued, nonzero for an invalid argument or an unsupported
 *     model (message from dynode_last_error(), thread-local).  There is NO CPU fallback: a model
 *     outside the compiled flow family fails loudly here.
 *   - per-trajectory numerical outcomes go to `stats` (the kernel is asynchronous).
 */
#ifndef DYNODE_B200_H_
#define DYNODE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DYNODE_B200_VERSION 100 /* 0.1.0 */

/* Compartment sets of the supported flow family.  State layout of one
 * trajectory: the compartments concatenated in this order, `s` of shape [G], every other
 * compartment of shape [G][S] (C order), G = n_groups (age, or age x risk flattened),
 * S = n_strains.  n = G + (ncomp-1)*G*S. */
enum {
  DYNODE_FLOW_SIR = 0,     /* (s, i, r)        examples/sir.py:78-84, sir_age_stratified.py:127-142,
                              sir_age_risk_stratified.py:157-173, tests/test_simulation/test_odes.py:17-28 */
  DYNODE_FLOW_SEIRS = 1,   /* (s, e, i, r)     examples/seirs.py:88-95, seirs_seasonal_forcing.py:40-55 */
  DYNODE_FLOW_SEIRS_C = 2  /* (s, e, i, r, c)  examples/seirs_multi_strain_age_stratified.py:213-243 */
};
enum {
  DYNODE_FLAG_SEASONAL = 1,   /* beta_t = beta*(1 + amp*sin(2*pi*t/period + phase))  (seirs_seasonal_forcing.py:34-37) */
  DYNODE_FLAG_DENSITY_DEP = 2 /* new infections beta*s*i, no division by N           (test_odes.py:23) */
};

typedef struct {
  int32_t flow;      /* DYNODE_FLOW_* */
  int32_t flags;     /* DYNODE_FLAG_* bit set */
  int32_t n_groups;  /* G */
  int32_t n_strains; /* S */
} DynodeModelDesc;

/* diffeqsolve arguments fixed by odes.py:107-144 and config/params.py:24-67:
 * Tsit5, PIDController(rtol, atol) (const_dt == 0) or ConstantStepSize (const_dt > 0),
 * t0, t1 = duration_days, dt0 = None (automatic), max_steps. */
typedef struct {
  double t0, t1;
  double rtol, atol;
  double const_dt;
  int64_t max_steps;
  /* Optional hint: if > 0 the caller guarantees save_ts[k] == t0 + k*save_dt for k < T-1 and
   * save_ts[T-1] == t1 bit-for-bit (build_saveat's linspace grid); the kernel then generates the
   * save times arithmetically instead of loading them.  0 = read save_ts. */
  double save_dt;
  /* SolverParams.discontinuity_points (odes.py:120-131: ClipStepSizeController(jump_ts)): n_jump <= 32
   * sorted times in DEVICE memory; steps end just before a jump and restart at it.  NULL / 0 = none.
   * Honoured by every entry point: dynode_solve_f64, dynode_solve_sens_f64 and dynode_poisson_loglik_grad_f64
   * (tangents ride the clipped step sequence) and dynode_poisson_loglik_adjoint_f64 (its forward sweep clips, its
   * reverse sweep rebuilds every step from its checkpoint). */
  const double* jump_ts;
  int32_t n_jump;
  /* Optional row mask: DEVICE [B] bytes.  When non-NULL, only trajectories with only[b] != 0 are integrated;
   * nothing is computed or written for the others (their rows of every output keep what the caller put
   * there).  A many-chain sampler passes its "chain is still running" flags, so finished chains cost nothing
   * (dynode_b200/infer/nuts.py).  NULL = all B trajectories. */
  const uint8_t* only;
} DynodeSolverDesc;

/* An ensemble array: element (b, k) lives at ptr[b*batch_stride + k]; batch_stride == 0 shares
 * one row across the whole ensemble.  ptr == NULL means "absent". */
typedef struct {
  const double* ptr;
  int64_t batch_stride;
} DynodeArray;

/* The ODE parameters of the flow family (the fields of the reference's *_ODEParams dataclasses). */
typedef struct {
  DynodeArray beta;          /* [B][S]  r0 / infectious_period */
  DynodeArray gamma;         /* [B][S]  1 / infectious_period */
  DynodeArray sigma;         /* [B][S]  1 / latent period        (flows with e) */
  DynodeArray omega;         /* [B][S]  1 / waning period        (flows with waning; absent = 0) */
  DynodeArray season_amp;    /* [B]     DYNODE_FLAG_SEASONAL only */
  DynodeArray season_phase;  /* [B] */
  DynodeArray season_period; /* [B] */
  const double* contact;     /* [G][G] shared: contact[target][source]; NULL = identity */
} DynodeParams;

/* Parameter ids for sensitivities: wrt[k] = DYNODE_WRT(kind, strain), or -1 for a direction that is
 * seeded only through dy0. */
enum { DYNODE_P_BETA = 0, DYNODE_P_GAMMA = 1, DYNODE_P_SIGMA = 2, DYNODE_P_OMEGA = 3,
       DYNODE_P_SEASON_AMP = 4, DYNODE_P_SEASON_PHASE = 5 };
#define DYNODE_WRT(kind, strain) ((kind) * 16 + (strain))

/* stats row per trajectory */
enum { DYNODE_STAT_RESULT = 0, DYNODE_STAT_ACCEPTED = 1, DYNODE_STAT_REJECTED = 2, DYNODE_STAT_STEPS = 3 };
enum { DYNODE_RESULT_OK = 0, DYNODE_RESULT_MAX_STEPS = 1, DYNODE_RESULT_ADJOINT_CAPACITY = 2 };

int dynode_version(void);
const char* dynode_last_error(void);

/* n (state size), number of compartments, size of the saved row for a compartment bit mask. */
int dynode_state_size(const DynodeModelDesc* model);
int dynode_num_compartments(const DynodeModelDesc* model);
int dynode_saved_size(const DynodeModelDesc* model, uint32_t save_comp_mask);
/* 1 if kernels for this model are compiled into the library, else 0 (and last_error says why). */
int dynode_is_supported(const DynodeModelDesc* model);

/* Replaces diffrax.diffeqsolve(ODETerm(ode), Tsit5(), t0, t1, None, y0, args, PIDController|ConstantStepSize,
 * SaveAt(ts) | SaveAt(subs=SubSaveAt(ts, fn)), max_steps) batched over B draws (odes.py:133-144).
 *   y0      [B][n]                 (batch_stride 0 = one shared initial state)
 *   save_ts [T]                    build_saveat's linspace grid (odes.py:177-179)
 *   save_comp_mask                 bit c set = compartment c is saved (sub_save_indices, odes.py:182-193)
 *   ys      [B][T][n_saved]        saved compartments concatenated; slots never reached stay +inf.  Any 8-byte
 *                                  aligned pointer; 16-byte alignment lets the 4-state flows store rows in pairs
 *   stats   [B][4]                 result, accepted, rejected, steps */
int dynode_solve_f64(const DynodeModelDesc* model, const DynodeSolverDesc* solver, int64_t B,
                     DynodeArray y0, const DynodeParams* params, const double* save_ts, int32_t T,
                     uint32_t save_comp_mask, double* ys, int32_t* stats, void* stream);

/* Same solve carrying forward sensitivities of the discrete scheme (step sequence frozen, as
 * diffrax's stop_gradient on the controller factor and on the automatic dt0 implies):
 *   wrt  HOST [n_wrt]              parameter ids (DYNODE_WRT) or -1
 *   dy0  [B][n_wrt][n] or NULL     tangents of the initial state
 *   dys  [B][T][n_saved][n_wrt] */
int dynode_solve_sens_f64(const DynodeModelDesc* model, const DynodeSolverDesc* solver, int64_t B,
                          DynodeArray y0, const DynodeParams* params, const double* save_ts, int32_t T,
                          uint32_t save_comp_mask, int32_t n_wrt, const int32_t* wrt, const double* dy0,
                          double* ys, double* dys, int32_t* stats, void* stream);

/* Fused log-density + gradient for NUTS (examples/sir_infer_parameters.py:30-38): nothing but
 * lp/grad/stats is written.
 *   rate = max(diff(ys[obs_comp], axis=time), 1e-6);  lp = sum(obs*log(rate) - rate) + lp_const
 *   obs  [T-1][m]   shared observations, m = size of compartment obs_comp
 *   lp   [B],  grad [B][n_wrt]  (d lp / d wrt-parameters; chain to r0 / infectious_period on the host) */
int dynode_poisson_loglik_grad_f64(const DynodeModelDesc* model, const DynodeSolverDesc* solver, int64_t B,
                                   DynodeArray y0, const DynodeParams* params, const double* save_ts,
                                   int32_t T, int32_t obs_comp, const double* obs, double lp_const,
                                   int32_t n_wrt, const int32_t* wrt, const double* dy0, double* lp,
                                   double* grad, int32_t* stats, void* stream);

/* Same log-likelihood with its gradient w.r.t. EVERY rate and (optionally) the initial state from one
 * reverse sweep over the accepted steps -- the discrete adjoint of the frozen-step scheme, i.e. what the
 * reference's reverse-mode pass through diffeqsolve yields (odes.py:133-144 with diffrax's default
 * RecursiveCheckpointAdjoint).  Cost ~4.5 solves whatever the number of parameters.
 *   grad     [B][4*S + 2]   d lp / d (beta_s, gamma_s, sigma_s, omega_s, season_amp, season_phase)
 *   grad_y0  [B][n] or NULL d lp / d y0
 *   ckpt     scratch [B][cap][n + 2]  (tprev, tnext, y_k) of every accepted step; a trajectory that accepts
 *            more than `cap` steps reports DYNODE_RESULT_ADJOINT_CAPACITY and NaN outputs
 *   vsave    scratch [B][T][m] */
int dynode_poisson_loglik_adjoint_f64(const DynodeModelDesc* model, const DynodeSolverDesc* solver, int64_t B,
                                      DynodeArray y0, const DynodeParams* params, const double* save_ts,
                                      int32_t T, int32_t obs_comp, const double* obs, double lp_const,
                                      double* lp, double* grad, double* grad_y0, int32_t* stats, double* ckpt,
                                      int32_t cap, double* vsave, void* stream);

/* Bench / roofline helpers: dependency-free FP64 FMA loop (returns flops done per launch) and a
 * streaming write, both enqueued on `stream`. */
int64_t dynode_probe_dfma(double* sink, int32_t iters, void* stream);
int dynode_probe_hbm_write(double* dst, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DYNODE_B200_H_ */
