// solve_args.h -- launch arguments shared by the kernels (lane_solver.cuh) and the C ABI (capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dynode_b200.h"

namespace dynode {

constexpr int kPMax = 2;            // max tangent directions carried per work item (more -> several groups)
constexpr int kMaxWrt = 64;         // max tangent directions per call
// Directions per pass chosen so that (1+P) * 8 * NE doubles stay in registers without spilling:
// SIR lanes hold 3 elements (P=2 -> 72 doubles), SEIRS/SEIRS_C lanes 4-5 elements (P=1 -> 80).
constexpr int tangent_chunk(int flow) { return flow == DYNODE_FLOW_SIR ? 2 : 1; }
#ifndef DYN_THREADS
#define DYN_THREADS 64
#endif
constexpr int kThreads = DYN_THREADS;  // warps per CTA = kThreads / 32
constexpr int MODE_SAVE = 0;        // write saved trajectories (+ tangents)
constexpr int MODE_LOGLIK = 1;      // fused Poisson-incidence log-likelihood (+ gradient)
constexpr int MODE_SAVE_JUMPS = 2;  // MODE_SAVE with ClipStepSizeController(jump_ts) step clipping
constexpr int MODE_LOGLIK_JUMPS = 3;  // MODE_LOGLIK with it
constexpr int kMaxJumps = 32;

#ifndef DYN_SMEM_OFFLOAD
#define DYN_SMEM_OFFLOAD (-1)
#endif

#ifndef DYN_OVERSUBSCRIBE
#define DYN_OVERSUBSCRIBE 1
#endif
constexpr int kOversubscribe = DYN_OVERSUBSCRIBE;  // grid = resident warps x this

struct SolveArgs {
  int64_t B;
  int64_t chunk;  // trajectories per warp (contiguous); set by the launcher
  DynodeArray y0;
  DynodeParams prm;
  const double* save_ts;
  int32_t T;
  uint32_t save_mask;
  double t0, t1, rtol, atol, const_dt;
  double save_dt;  // > 0: save_ts is the uniform grid t0 + k*save_dt (k < T-1), ts[T-1] = t1
  int32_t max_steps;
  double* ys;      // [B][T][n_saved]   (written when write_primal)
  int32_t* stats;  // [B][4]            (written when write_primal)
  int32_t write_primal;
  // discontinuity points (SolverParams.discontinuity_points -> jump_ts), sorted ascending, device memory
  const double* jump_ts;
  int32_t n_jump;
  // sensitivities: P_total directions ride the primal's steps in n_pass groups of P (the kernel's template
  // chunk).  All groups run in ONE launch: work item v = trajectory * n_pass + group integrates trajectory
  // v / n_pass carrying directions group*P .. group*P+P-1; group 0 also writes the primal outputs.
  int32_t P_total, n_pass;
  int32_t wrt[kMaxWrt];
  const double* dy0;  // [B][P_total][n]
  double* dys;        // [B][T][n_saved][P_total]
  // fused log-likelihood
  int32_t obs_comp;
  const double* obs;  // [T-1][m]
  double lp_const;
  double* lp;    // [B]
  double* grad;  // [B][P_total]
  const uint8_t* only;  // [B] or NULL: row mask (DynodeSolverDesc.only)
  int32_t refill_min;   // persistent-slot instances: refill once this many slots of a warp are idle (launcher)
};

// Discrete-adjoint log-likelihood kernel (adjoint_solver.cuh)
struct AdjointArgs {
  SolveArgs s;        // B, y0, prm, save_ts, T, t0, t1, rtol, atol, max_steps, obs_comp, obs, lp_const, lp, stats
  double* grad;       // [B][4*S + 2]: d lp / d (beta_s, gamma_s, sigma_s, omega_s, season_amp, season_phase)
  double* grad_y0;    // [B][n] or NULL
  double* ckpt;       // scratch [B][cap][n + 2]: per accepted step (tprev, tnext, y_k)
  double* vsave;      // scratch [B][T][m]: observed compartment at the save times, then its cotangent
  int32_t cap;        // accepted steps that fit the scratch
};
template <int FLOW, int FLAGS, int G, int S, bool JUMPS>
cudaError_t launch_adjoint_solver(const AdjointArgs& a, cudaStream_t stream);

// The fused log-likelihood carrying ONE direction per work item, for the flows whose production chunk is two
// (FLOW_SIR): at a few chains a launch is as long as one warp's instruction stream, and a work item with one tangent
// executes two thirds of the instructions of one with two -- the primal is repeated, which costs nothing while the
// GPU is empty.  cudaErrorNotSupported for flows whose chunk is already one.
template <int FLOW, int FLAGS, int G, int S>
cudaError_t launch_loglik_single_direction(const SolveArgs& a, cudaStream_t stream);

// Defined in lane_solver.cuh, explicitly instantiated per model in inst.cu (see instances.def).
template <int FLOW, int FLAGS, int G, int S, int P, int MODE>
cudaError_t launch_lane_solver(const SolveArgs& a, cudaStream_t stream);

}  // namespace dynode
