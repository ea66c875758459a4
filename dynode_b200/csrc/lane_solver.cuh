// lane_solver.cuh -- the ensemble Tsit5 kernel: one trajectory per group of L = G*S lanes,
// every lane owning the compartments of one (population group g, strain s) cell in REGISTERS.
//
// Replaces, for a whole ensemble in one launch, what the reference runs per trajectory through
// diffrax.diffeqsolve (reference src/dynode/simulation/odes.py:133-144): the while-loop of
// adaptive Tsit5 steps, the I-controller, and the SaveAt(ts) dense-output writes
// (SURVEY.md 8a rows a3-a9), on the compartmental right-hand sides of a11.
//
// Layout (DESIGN.md "Kernel K1"):
//   lane (g, s) holds  S_g (replicated across the S lanes of a group, kept bit-identical by
//   order-fixed segmented sums), and E/I/R/C[g, s]: NE <= 5 state elements, 8*NE doubles with the
//   7 stage derivatives, all in registers -> no shared/local memory traffic in the step loop.
//   A warp has TPW = 32/L trajectory SLOTS (SIR/SEIRS 1-bin: 32, age SIR: 16, multi-strain 2x3: 5).
//   Cross-lane terms (N_g, the contact contraction, dS_g, the RMS error norm) are warp shuffles.
//   FP64 FMA on the CUDA cores; no tensor cores (contractions are <= 6x6).
//
// Persistent slots: adaptive step counts differ per draw (47..129 attempts for seasonal SEIRS), so
// a warp that integrated a fixed set of trajectories would idle its finished slots until the
// slowest one ends (measured: 10% of lane-steps for the 5-slot multi-strain case, 49% for 32
// slots).  Instead every warp owns a contiguous chunk of the ensemble and a finished slot is
// refilled with the warp's next trajectory (initial-step selection runs masked inside the loop).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/dynode_b200.h"
#include "dual.cuh"
#include "solve_args.h"
#include "tsit5.cuh"

namespace dynode {

// clamp of the Poisson rate and its logarithm (fused log-likelihood)
static __constant__ double kLikC[2] = {1e-6, -13.815510557964274};

// Taylor coefficients of sin(d)/d - 1 and cos(d) - 1 in z = d^2, highest power first (sincos_small)
static __constant__ double kSinTaylor[7] = {-1.0 / 1307674368000.0, 1.0 / 6227020800.0, -1.0 / 39916800.0,
                                            1.0 / 362880.0,         -1.0 / 5040.0,      1.0 / 120.0,
                                            -1.0 / 6.0};
static __constant__ double kCosTaylor[8] = {1.0 / 20922789888000.0, -1.0 / 87178291200.0, 1.0 / 479001600.0,
                                            -1.0 / 3628800.0,       1.0 / 40320.0,        -1.0 / 720.0,
                                            1.0 / 24.0,             -0.5};

template <int FLOW, int FLAGS, int G, int S, int P, int MODE>
struct LaneSolver {
  static constexpr bool HAS_E = FLOW != DYNODE_FLOW_SIR;
  static constexpr bool HAS_C = FLOW == DYNODE_FLOW_SEIRS_C;
  static constexpr bool WANING = FLOW != DYNODE_FLOW_SIR;
  static constexpr bool SEASONAL = (FLAGS & DYNODE_FLAG_SEASONAL) != 0;
  static constexpr bool DENSITY = (FLAGS & DYNODE_FLAG_DENSITY_DEP) != 0;
  static constexpr int NE = 3 + (HAS_E ? 1 : 0) + (HAS_C ? 1 : 0);  // elements per lane
  static constexpr int IE = 1;
  static constexpr int II = HAS_E ? 2 : 1;
  static constexpr int IR = II + 1;
  static constexpr int IC = IR + 1;
  static constexpr int L = G * S;
  static constexpr int TPW = 32 / L;  // trajectory slots per warp
  static constexpr int N = G + (NE - 1) * G * S;
  // Persistent slots pay off only where many slots share a warp (measured on B200, profiles/
  // r1_tuning.md: +11% for 32 slots, -8% for the 5-slot multi-strain case whose step counts are
  // tight); otherwise a warp integrates one generation of TPW trajectories.
  static constexpr bool JUMPS = MODE == MODE_SAVE_JUMPS || MODE == MODE_LOGLIK_JUMPS;
  static constexpr bool IS_SAVE = MODE == MODE_SAVE || MODE == MODE_SAVE_JUMPS;
#ifndef DYN_PERSIST_MIN_TPW
#define DYN_PERSIST_MIN_TPW 8
#endif
#ifndef DYN_REFILL_SMALL
#define DYN_REFILL_SMALL 1
#endif
  static constexpr bool PERSIST = TPW >= DYN_PERSIST_MIN_TPW;
  // refill as soon as this many slots are idle (masked re-initialisation costs ~1.5 steps)
  // (measured on B200, DYNODE_REFILL sweep: a quarter of the slots beats an eighth by 5 % on the 16-slot fused
  // log-likelihood of C2 -- 22-step trajectories, so refills are frequent -- and by 2 % on the 32-slot C3 solve)
  static constexpr int REFILL = TPW >= 8 ? TPW / 4 : DYN_REFILL_SMALL;
  static_assert(L >= 1 && L <= 32, "a trajectory must fit one warp");
  using D = Dual<P>;
  // Shared-memory offload: the dense-output coefficients Q (3*NE doubles, written once per step, read only
  // by the save loop) and the per-trajectory rates + contact row live in per-thread columns of shared
  // memory instead of registers.  Measured on B200 (profiles/r1/kernel_variants.md): +4 % for the 1-bin
  // SEIRS lanes (seasonal C3 workload), -3 % for the 5-element SEIRS+C lane even though it then fits 16
  // instead of 12 warps per SM -- so it is on for FLOW_SEIRS only (DYN_SMEM_OFFLOAD: -1 auto, 0 off, 1 all).
  static constexpr bool OFFLOAD = P == 0 && (DYN_SMEM_OFFLOAD == 1 || (DYN_SMEM_OFFLOAD < 0 && FLOW == DYNODE_FLOW_SEIRS));
  // the dense-output coefficients are offloaded only where whole rows are saved: the fused log-likelihood keeps the
  // observed compartment's three coefficients in registers
  static constexpr bool OFFLOAD_Q = OFFLOAD && IS_SAVE;
  static constexpr int OFF_Q = 0, OFF_PRM = 3 * NE, OFF_K = OFF_PRM + 4, OFF_SEAS = OFF_K + G;
  static constexpr int NOFF = OFFLOAD ? OFF_SEAS + (SEASONAL ? 3 : 0) : 1;
  // volatile: the rates are loop-invariant, and the whole point is that they are NOT kept in registers
  static DYN_DI double lds_v(const double* p) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
  }

  struct Geo {  // lane geometry + the shared contact row (fixed for the whole kernel)
    int base;   // lane of cell (0,0) of my slot
    int sbase;  // lane of cell (g,0)
    int g, s;
  };
  struct Prm {  // per-trajectory parameters of my cell
    D beta, gamma, sigma, omega, amp, phase;
    double period;
  };

  // sum over the S strain lanes of my group, in fixed order (identical in every lane of the group)
  static DYN_DI D sum_strains(const D& x, const Geo& c) {
    if constexpr (S == 1) {
      return x;
    } else {
      D r = dual_shfl(x, c.sbase);
#pragma unroll
      for (int k = 1; k < S; ++k) r = r + dual_shfl(x, c.sbase + k);
      return r;
    }
  }
  static DYN_DI double sum_strains(double x, const Geo& c) {
    if constexpr (S == 1) {
      return x;
    } else {
      double r = __shfl_sync(0xffffffffu, x, c.sbase);
#pragma unroll
      for (int k = 1; k < S; ++k) r += __shfl_sync(0xffffffffu, x, c.sbase + k);
      return r;
    }
  }
  // sum over groups of a value that is already identical across the strain lanes of each group
  static DYN_DI double sum_groups(double x, const Geo& c) {
    if constexpr (G == 1) {
      return x;
    } else {
      double r = __shfl_sync(0xffffffffu, x, c.base + c.s);
#pragma unroll
      for (int b = 1; b < G; ++b) r += __shfl_sync(0xffffffffu, x, c.base + b * S + c.s);
      return r;
    }
  }
  // sum of a per-lane partial over the whole trajectory; identical in all of its lanes
  static DYN_DI double traj_sum(double x, const Geo& c) { return sum_groups(sum_strains(x, c), c); }

  // 1 / N_g with N_g = s_g + sum_s (e+i+r)[g,s]  (c excluded).  N_g is a linear invariant of every
  // flow in the family (each transfer leaves one compartment of group g and enters another), and an
  // explicit Runge-Kutta stage y + h*sum a_ij f_j preserves linear invariants to rounding.  The
  // kernel therefore forms 1/N_g once per step from the step's initial state and reuses it for the
  // stages of that step; the reference re-sums N at every stage, which differs by O(1e-16) relative.
  static DYN_DI D inv_population(const D (&y)[NE], const Geo& c) {
    if constexpr (DENSITY) {
      return make_dual<P>(1.0);
    } else {
      D part = y[II] + y[IR];
      if constexpr (HAS_E) part = part + y[IE];
      const D Ng = y[0] + sum_strains(part, c);
      return drcp_fast(Ng);
    }
  }

  // ---- right-hand side of the flow family in lane layout -------------
  // sin / cos of the seasonal angle 2*pi*t/period + phase (seirs_seasonal_forcing.py:34-37)
  static DYN_DI void season_angle(double t, const Prm& p, double& sn, double& cs) {
    // quotient by reciprocal + one correction (faithfully rounded; the IEEE division sequence costs ~4x as much and
    // runs once or twice per step)
    const double w = div_fast((2.0 * CUDART_PI) * t, p.period);
    sincos(p.phase.v + w, &sn, &cs);
  }
  // sin / cos of a small increment |d| <= 0.5 by Taylor series in d^2 (remainder < 1e-18): 16 FMAs instead of
  // a libm sincos (~75 instructions, a third of them 64-bit constant moves); used with the angle-addition
  // formulas for the stage times inside a step, the base angle being evaluated exactly once per step
  static DYN_DI void sincos_small(double d, double& sd, double& cd) {
    // coefficients from constant memory (kSinTaylor / kCosTaylor below): c[bank][offset] operands of DFMA; as
    // literals they cost 24 UMOVs per call, 96 per step (SASS of the seasonal instance)
    const double z = d * d;
    double ps = kSinTaylor[0];
#pragma unroll
    for (int k = 1; k < 7; ++k) ps = fma(ps, z, kSinTaylor[k]);
    sd = fma(d * z, ps, d);
    double pc = kCosTaylor[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) pc = fma(pc, z, kCosTaylor[k]);
    cd = fma(z, pc, 1.0);
  }

  static DYN_DI void rhs(double t, const D (&y)[NE], D (&dy)[NE], const Geo& c, const double (&K)[G],
                         const Prm& p, const D& invN) {
    double sn = 0.0, cs = 1.0;
    if constexpr (SEASONAL) season_angle(t, p, sn, cs);
    rhs_sc(sn, cs, y, dy, c, K, p, invN);
  }

  // the right-hand side with the seasonal sine / cosine already known
  static DYN_DI void rhs_sc(double sn, double cs, const D (&y)[NE], D (&dy)[NE], const Geo& c,
                            const double (&K)[G], const Prm& p, const D& invN) {
    D prop;
    if constexpr (DENSITY) {
      prop = y[II];  // tests/test_simulation/test_odes.py:23  s_to_i = beta*s*i
    } else {
      prop = y[II] * invN;
    }
    // contact contraction: sum_b K[g][b] * prop[b, s]
    D acc;
    if constexpr (G == 1) {
      acc = K[0] * prop;
    } else {
      acc = K[0] * dual_shfl(prop, c.base + c.s);
#pragma unroll
      for (int b = 1; b < G; ++b) acc = dfma(K[b], dual_shfl(prop, c.base + b * S + c.s), acc);
    }
    D beta_t = p.beta;
    if constexpr (SEASONAL) {
      // beta*(1 + amp*sin(2*pi*t/period + phase))   (seirs_seasonal_forcing.py:34-37)
      D seas;
      seas.v = fma(p.amp.v, sn, 1.0);
      if constexpr (P > 0) {
#pragma unroll
        for (int k = 0; k < P; ++k) seas.d[k] = fma(p.amp.d[k], sn, p.amp.v * cs * p.phase.d[k]);
      }
      beta_t = p.beta * seas;
    }
    const D foi = beta_t * acc;
    const D newinf = foi * y[0];
    const D rec = p.gamma * y[II];
    if constexpr (WANING) {
      dy[0] = sum_strains(dmsub(p.omega, y[IR], newinf), c);  // ds_g = sum_s (omega_s r - newinf)
      dy[IR] = dnmadd(p.omega, y[IR], rec);                   // gamma i - omega r
    } else {
      dy[0] = sum_strains(dneg(newinf), c);
      dy[IR] = rec;
    }
    if constexpr (HAS_E) {
      dy[IE] = dnmadd(p.sigma, y[IE], newinf);  // newinf - sigma e
      dy[II] = dmsub(p.sigma, y[IE], rec);      // sigma e - gamma i
    } else {
      dy[II] = newinf - rec;
    }
    if constexpr (HAS_C) dy[IC] = newinf;
  }

  static DYN_DI double sq(double x) { return x * x; }
  // base + k rows of N doubles, written as byte arithmetic so that it folds into one IMAD.WIDE
  static DYN_DI double* row_ptr(double* base, int k) {
    return reinterpret_cast<double*>(reinterpret_cast<char*>(base) + (int64_t)k * (int64_t)(N * 8));
  }

  // ClipStepSizeController(jump_ts): a step [t0, t1] that would contain a discontinuity point ends just
  // before it.  i0 = #{jump <= t0}, i1 = #{jump <= t1}; a jump lies in (t0, t1] iff
  // i0 < i1, and then t1 <- prevbefore(jump[i0]).
  static DYN_DI double clip_to_jumps(const SolveArgs& a, double t0, double t1, bool& made_jump) {
    int i0 = 0, i1 = 0;
    for (int k = 0; k < a.n_jump; ++k) {
      const double j = __ldg(a.jump_ts + k);
      i0 += (j <= t0);
      i1 += (j <= t1);
    }
    made_jump = i0 < i1;
    return made_jump ? nextafter(__ldg(a.jump_ts + (i0 < a.n_jump ? i0 : a.n_jump - 1)), -CUDART_INF) : t1;
  }

  // ---- the kernel body ------------------------------------------------------------------
  static __device__ void run(const SolveArgs& a) {
    using namespace tsit5;
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int tw = lane / L;
    const int q = lane - tw * L;
    Geo c;
    c.g = q / S;
    c.s = q - c.g * S;
    c.base = tw * L;
    c.sbase = c.base + c.g * S;
    double Kr[G];  // contact[g][:]
#pragma unroll
    for (int b = 0; b < G; ++b)
      Kr[b] = a.prm.contact ? __ldg(a.prm.contact + c.g * G + b) : (b == c.g ? 1.0 : 0.0);
    __shared__ double sm_off[NOFF][kThreads];
    double* const my = &sm_off[0][threadIdx.x];  // value v of this thread sits at my[v * kThreads]
    if constexpr (OFFLOAD) {
#pragma unroll
      for (int b = 0; b < G; ++b) my[(OFF_K + b) * kThreads] = Kr[b];
    }
    // the parameters of my slot's trajectory as the step loop sees them
    auto cur_K = [&](double (&K)[G]) {
#pragma unroll
      for (int b = 0; b < G; ++b) K[b] = OFFLOAD ? lds_v(my + (OFF_K + b) * kThreads) : Kr[b];
    };
    const bool slot_ok = tw < TPW;   // lanes beyond the last whole slot never own a trajectory
    const bool lead = (c.s == 0);    // owner of the replicated S_g for norms and stores
    const bool head = slot_ok && q == 0;

    // this warp's contiguous chunk of the ensemble
    const int64_t chunk_begin = warp_global * a.chunk;
    const int64_t Bv = a.B * a.n_pass;  // work items: (trajectory, tangent group) pairs
    const int64_t chunk_end = (chunk_begin + a.chunk < Bv) ? chunk_begin + a.chunk : Bv;
    int64_t next = chunk_begin;  // warp-uniform: first trajectory not yet handed to a slot
    if (chunk_begin >= Bv) return;

    // ---- element offsets inside a full state row and inside a saved row
    int off_full[NE], off_save[NE];
    off_full[0] = c.g;
#pragma unroll
    for (int e = 1; e < NE; ++e) off_full[e] = G + (e - 1) * G * S + c.g * S + c.s;
    int n_saved = 0;
    {
      int run_off = 0;
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        const int sz = (e == 0) ? G : G * S;
        const bool on = (a.save_mask >> e) & 1u;
        off_save[e] = on ? run_off + (e == 0 ? c.g : c.g * S + c.s) : -1;
        if (on) run_off += sz;
      }
      n_saved = run_off;
    }
    if (!lead) off_save[0] = -1;  // S_g is stored once, by the strain-0 lane
    // the one-lane-per-row instances store rows as 16-byte pairs: a caller's ys that is only 8-byte aligned takes
    // the general path (8-byte stores) instead of faulting
    const bool full_save = a.write_primal && a.save_mask == ((1u << NE) - 1u) &&
                           (!(L == 1 && NE == 4) || (reinterpret_cast<uintptr_t>(a.ys) & 15u) == 0);
    // save time k: generated arithmetically for build_saveat's uniform grid, else loaded
    auto save_time = [&](int k) -> double {
      if (k >= a.T) return CUDART_INF;
      if (a.save_dt > 0.0) return (k == a.T - 1) ? a.t1 : fma((double)k, a.save_dt, a.t0);
      return __ldg(a.save_ts + k);
    };
    const double t1 = a.t1, rtol = a.rtol, atol = a.atol;
    const double inv_n = 1.0 / (double)N;
    const int obs_m = (a.obs_comp == 0) ? G : G * S;
    const int obs_q = (a.obs_comp == 0) ? c.g : c.g * S + c.s;
    const bool obs_owner = (a.obs_comp != 0) || lead;

    // ---- per-slot state (registers)
    Prm prm;
    D y[NE], f[7][NE], ys[NE];
    int64_t traj = chunk_begin / a.n_pass;  // trajectory of my slot (valid index even while idle)
    int p0s = 0;                            // first tangent direction my slot carries
    bool wp = a.write_primal != 0;          // my slot writes the primal outputs (tangent group 0)
    double tprev = t1, tnext = t1;
    int32_t n_acc = 0, n_rej = 0, n_steps = 0, save_i = 0;
    bool active = false;
    double sn0 = 0.0, cs0 = 1.0;  // seasonal sine / cosine at tprev of my slot (exact, libm)
    bool made_jump = false;  // the running step was clipped to end just before a discontinuity point
    double* out_s = a.ys;  // row 0 of my trajectory in ys (my S_g / my cell): the full-save fast path
    double* out_c = a.ys;
    D lp_acc = make_dual<P>(0.0), obs_prev = make_dual<P>(0.0);
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      y[e] = make_dual<P>(1.0);
      f[0][e] = make_dual<P>(0.0);
    }
    prm.beta = prm.gamma = prm.sigma = prm.omega = prm.amp = prm.phase = make_dual<P>(0.0);
    prm.period = 1.0;

    bool had_traj = false;
    // ================= refill idle slots from the warp's chunk =================
    auto refill = [&]() {
      {
        const unsigned idle_heads = __ballot_sync(0xffffffffu, head && !active);
        const int n_idle = __popc(idle_heads);
        const bool any_active = __any_sync(0xffffffffu, active);
        if (next < chunk_end && n_idle > 0 && (n_idle >= a.refill_min || !any_active)) {
          // candidates: the next 32 work items of the chunk minus the rows the caller masked out (a.only);
          // the idle slot of rank r takes the (r+1)-th candidate
          const int64_t ci = next + lane;
          bool cok = ci < chunk_end;
          if (a.only != nullptr && cok) cok = __ldg(a.only + ((a.n_pass == 1) ? ci : ci / a.n_pass)) != 0;
          const unsigned cmask = __ballot_sync(0xffffffffu, cok);
          const int rank = __popc(idle_heads & ((1u << c.base) - 1u));  // idle slots below mine
          const unsigned pos = __fns(cmask, 0u, rank + 1);              // 0xffffffff: no candidate left for me
          const bool take = slot_ok && !active && pos < 32u;
          const int64_t cand = next + (take ? (int64_t)pos : 0);
          // consumed: everything up to the last candidate handed out, or all 32 scanned items
          const unsigned used = (__popc(cmask) >= n_idle) ? __fns(cmask, 0u, n_idle) + 1u : 32u;
          next = (next + used < chunk_end) ? next + used : chunk_end;
          const int64_t cand_tr = (a.n_pass == 1) ? cand : cand / a.n_pass;
          const int cand_p0 = (a.n_pass == 1) ? 0 : (int)(cand - cand_tr * a.n_pass) * P;
          const int64_t tr = take ? cand_tr : traj;
          const int p0n = take ? cand_p0 : p0s;
          // ---- parameters and initial state of the new trajectory.  Only the slots that take one load anything:
          // lanes that keep integrating their trajectory go through the masked initialisation below on placeholder
          // values (all 1.0: finite everywhere) whose results are discarded.  Letting them re-load the rows of their
          // current trajectory instead cost 4.6x the algorithmic DRAM reads on the 32-slot seasonal kernel (ncu r1:
          // 403 MB against 88 MB -- three quarters of every refill's loads were scattered re-reads).
          auto ld = [&](const DynodeArray& arr, int k, double dflt) -> double {
            if (!arr.ptr) return dflt;
            return take ? __ldg(arr.ptr + tr * arr.batch_stride + k) : 1.0;
          };
          Prm pn;
          pn.beta = make_dual<P>(ld(a.prm.beta, c.s, 0.0));
          pn.gamma = make_dual<P>(ld(a.prm.gamma, c.s, 0.0));
          pn.sigma = make_dual<P>(HAS_E ? ld(a.prm.sigma, c.s, 0.0) : 0.0);
          pn.omega = make_dual<P>(WANING ? ld(a.prm.omega, c.s, 0.0) : 0.0);
          pn.amp = make_dual<P>(SEASONAL ? ld(a.prm.season_amp, 0, 0.0) : 0.0);
          pn.phase = make_dual<P>(SEASONAL ? ld(a.prm.season_phase, 0, 0.0) : 0.0);
          pn.period = SEASONAL ? ld(a.prm.season_period, 0, 1.0) : 1.0;
          if constexpr (P > 0) {
#pragma unroll
            for (int k = 0; k < P; ++k) {
              const int w = (p0n + k < a.P_total) ? a.wrt[p0n + k] : -1;
              if (w >= 0 && (w & 15) == c.s) {
                const int kind = w >> 4;
                if (kind == DYNODE_P_BETA) pn.beta.d[k] = 1.0;
                if (kind == DYNODE_P_GAMMA) pn.gamma.d[k] = 1.0;
                if (kind == DYNODE_P_SIGMA) pn.sigma.d[k] = 1.0;
                if (kind == DYNODE_P_OMEGA) pn.omega.d[k] = 1.0;
              }
              if (w >= 0 && (w >> 4) == DYNODE_P_SEASON_AMP) pn.amp.d[k] = 1.0;
              if (w >= 0 && (w >> 4) == DYNODE_P_SEASON_PHASE) pn.phase.d[k] = 1.0;
            }
          }
          D yn[NE], fn[NE], f1[NE];
#pragma unroll
          for (int e = 0; e < NE; ++e) {
            yn[e] = make_dual<P>(take ? __ldg(a.y0.ptr + tr * a.y0.batch_stride + off_full[e]) : 1.0);
            if constexpr (P > 0) {
              if (a.dy0 && take) {
#pragma unroll
                for (int k = 0; k < P; ++k)
                  if (p0n + k < a.P_total)
                    yn[e].d[k] = __ldg(a.dy0 + ((tr * a.P_total + p0n + k) * (int64_t)N) + off_full[e]);
              }
            }
          }
          const D invN0 = inv_population(yn, c);
          double Kl[G];
          cur_K(Kl);
          rhs(a.t0, yn, fn, c, Kl, pn, invN0);  // FSAL f0 (solver.init)
          double tn;
          if (a.const_dt > 0.0) {
            tn = a.t0 + a.const_dt;  // ConstantStepSize (odes.py:115-118)
          } else {
            // Hairer-Wanner initial step (PIDController._select_initial_step).  It runs masked
            // inside the step loop of the persistent-slot instances, for the whole warp at every refill, so it is
            // kept short: one reciprocal per scale instead of three IEEE divisions, quotients by reciprocal +
            // correction, and the probe RHS on the primal alone (the result is under stop_gradient).
            double p0 = 0.0, p1 = 0.0;
            double inv_scale[NE];
#pragma unroll
            for (int e = 0; e < NE; ++e) {
              inv_scale[e] = rcp_fast(fma(fabs(yn[e].v), rtol, atol));
              const double w = (e == 0 && !lead) ? 0.0 : 1.0;
              p0 += w * sq(yn[e].v * inv_scale[e]);
              p1 += w * sq(fn[e].v * inv_scale[e]);
            }
            const double d0 = sqrt(traj_sum(p0, c) * inv_n);
            const double d1 = sqrt(traj_sum(p1, c) * inv_n);
            const bool small = (d0 < 1e-5) || (d1 < 1e-5);
            const double h0 = small ? 1e-6 : 0.01 * div_fast(d0, small ? 1.0 : d1);
            double df[NE];  // f(t0 + h0, y0 + h0 f0) - f0, primal
            if constexpr (P == 0) {
#pragma unroll
              for (int e = 0; e < NE; ++e) ys[e] = dfma(h0, fn[e], yn[e]);
              rhs(a.t0 + h0, ys, f1, c, Kl, pn, invN0);
#pragma unroll
              for (int e = 0; e < NE; ++e) df[e] = f1[e].v - fn[e].v;
            } else {
              using L0 = LaneSolver<FLOW, FLAGS, G, S, 0, MODE>;
              typename L0::Geo c0;
              c0.base = c.base; c0.sbase = c.sbase; c0.g = c.g; c0.s = c.s;
              typename L0::Prm q0;
              q0.beta.v = pn.beta.v; q0.gamma.v = pn.gamma.v; q0.sigma.v = pn.sigma.v; q0.omega.v = pn.omega.v;
              q0.amp.v = pn.amp.v; q0.phase.v = pn.phase.v; q0.period = pn.period;
              Dual<0> yv[NE], fv[NE], iN;
              iN.v = invN0.v;
#pragma unroll
              for (int e = 0; e < NE; ++e) yv[e].v = fma(h0, fn[e].v, yn[e].v);
              L0::rhs(a.t0 + h0, yv, fv, c0, Kl, q0, iN);
#pragma unroll
              for (int e = 0; e < NE; ++e) df[e] = fv[e].v - fn[e].v;
            }
            double p2 = 0.0;
#pragma unroll
            for (int e = 0; e < NE; ++e) {
              const double w = (e == 0 && !lead) ? 0.0 : 1.0;
              p2 += w * sq(df[e] * inv_scale[e]);
            }
            const double d2 = div_fast(sqrt(traj_sum(p2, c) * inv_n), h0);
            const double md = fmax(d1, d2);
            const double h1 = (md <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(div_fast(0.01, md), 0.2);
            tn = a.t0 + fmin(100.0 * h0, h1);
          }
          if (take) {
            if constexpr (OFFLOAD) {
              my[(OFF_PRM + 0) * kThreads] = pn.beta.v;
              my[(OFF_PRM + 1) * kThreads] = pn.gamma.v;
              my[(OFF_PRM + 2) * kThreads] = pn.sigma.v;
              my[(OFF_PRM + 3) * kThreads] = pn.omega.v;
              if constexpr (SEASONAL) {
                my[(OFF_SEAS + 0) * kThreads] = pn.amp.v;
                my[(OFF_SEAS + 1) * kThreads] = pn.phase.v;
                my[(OFF_SEAS + 2) * kThreads] = pn.period;
              }
            } else {
              prm = pn;
            }
#pragma unroll
            for (int e = 0; e < NE; ++e) { y[e] = yn[e]; f[0][e] = fn[e]; }
            traj = cand_tr;
            p0s = cand_p0;
            wp = a.write_primal && cand_p0 == 0;
            tprev = a.t0;
            if constexpr (SEASONAL) season_angle(a.t0, pn, sn0, cs0);
            if constexpr (JUMPS) tn = clip_to_jumps(a, a.t0, tn, made_jump);
            tnext = fmin(tn, t1);
            n_acc = n_rej = n_steps = 0;
            save_i = 0;
            out_s = a.ys + cand_tr * (int64_t)a.T * N + c.g;
            out_c = a.ys + cand_tr * (int64_t)a.T * N + G + c.g * S + c.s;
            lp_acc = make_dual<P>(0.0);
            obs_prev = make_dual<P>(0.0);
            active = true;
            had_traj = true;
          }
        }
      }
    };
    // ================= retire slots whose trajectory is complete (or ran out of steps) ==========
    auto retire = [&](bool fin) {
      if (__any_sync(0xffffffffu, fin)) {
        if constexpr (IS_SAVE) {
          if (fin && wp) {
            // slots never reached keep diffrax's +inf fill
            for (int k = save_i; k < a.T; ++k) {
              const int64_t row = (traj * a.T + k) * (int64_t)n_saved;
#pragma unroll
              for (int e = 0; e < NE; ++e)
                if (off_save[e] >= 0) a.ys[row + off_save[e]] = CUDART_INF;
            }
          }
        } else {
          // a trajectory that ran out of max_steps has a truncated sum: the reference raises (diffrax throw=True),
          // the asynchronous kernel reports NaN for lp and its gradient -- as the adjoint kernel does -- so that a
          // sampler sees a non-finite potential instead of a plausible number (stats carry the result code)
          const bool failed = tprev < t1;
          const double tot = traj_sum(obs_owner ? lp_acc.v : 0.0, c);
          if (fin && q == 0 && wp) a.lp[traj] = failed ? CUDART_NAN : tot + a.lp_const;
          if constexpr (P > 0) {
#pragma unroll
            for (int k = 0; k < P; ++k) {
              const double gp = traj_sum(obs_owner ? lp_acc.d[k] : 0.0, c);
              if (fin && q == 0 && p0s + k < a.P_total) a.grad[traj * a.P_total + p0s + k] = failed ? CUDART_NAN : gp;
            }
          }
        }
        if (fin && q == 0 && wp) {
          int32_t* st = a.stats + traj * 4;
          st[DYNODE_STAT_RESULT] = (tprev < t1) ? DYNODE_RESULT_MAX_STEPS : DYNODE_RESULT_OK;
          st[DYNODE_STAT_ACCEPTED] = n_acc;
          st[DYNODE_STAT_REJECTED] = n_rej;
          st[DYNODE_STAT_STEPS] = n_steps;
        }
        if (fin) {
          active = false;
          tprev = tnext = t1;  // idle slots step with h = 0 (harmless) until refilled
        }
      }
    };

    if constexpr (!PERSIST) refill();  // one generation: every slot takes its trajectory up front
    while (true) {
      if constexpr (PERSIST) refill();
      // a new trajectory whose horizon is empty (t0 == t1) or max_steps == 0 finishes in the commit
      // block below after one masked pass; nothing to do when no slot is active
      if (!__any_sync(0xffffffffu, active)) {
        if (PERSIST && next < chunk_end) continue;  // every candidate of this pass was masked out: scan on
        break;
      }
      const bool stepping = active && (tprev < t1) && (n_steps < a.max_steps);

      const double h = tnext - tprev;
      const D invN = inv_population(y, c);
      // Rates and contact row of my cell for this step.  Offloaded kernels re-read them from shared memory
      // ONCE per step through an index the compiler cannot prove loop-invariant, so they occupy registers
      // only while the stages run (not during the Q / error / save phases where pressure peaks) and the
      // loads issue early, under the 1/N_g computation.
      Prm pl;
      double Kl[G];
      if constexpr (OFFLOAD) {
        int o = 0;
        asm volatile("" : "+r"(o));  // opaque zero
        const double* src = my + o;
        pl.beta.v = src[(OFF_PRM + 0) * kThreads];
        pl.gamma.v = src[(OFF_PRM + 1) * kThreads];
        pl.sigma.v = src[(OFF_PRM + 2) * kThreads];
        pl.omega.v = src[(OFF_PRM + 3) * kThreads];
        if constexpr (SEASONAL) {
          pl.amp.v = src[(OFF_SEAS + 0) * kThreads];
          pl.phase.v = src[(OFF_SEAS + 1) * kThreads];
          pl.period = src[(OFF_SEAS + 2) * kThreads];
        } else {
          pl.amp.v = pl.phase.v = 0.0;
          pl.period = 1.0;
        }
#pragma unroll
        for (int b = 0; b < G; ++b) Kl[b] = src[(OFF_K + b) * kThreads];
      } else {
        pl = prm;
#pragma unroll
        for (int b = 0; b < G; ++b) Kl[b] = Kr[b];
      }
      // Seasonal angles of the step: the angle at tnext exactly (libm, shared by the two c = 1 stages and
      // -- through FSAL -- the next step's first stage), the four inner stage times by angle addition from
      // the exact angle at tprev with Taylor sin/cos of the increment 2*pi*c_i*h/period (<= 0.5 rad, else libm)
      double sn1 = 0.0, cs1 = 1.0, wstep = 0.0;
      bool small_step = true;
      if constexpr (SEASONAL) {
        season_angle(tnext, pl, sn1, cs1);
        wstep = div_fast((2.0 * CUDART_PI) * h, pl.period);
        // per lane, not per warp: a trajectory's arithmetic must not depend on which trajectories share its warp
        // (a permutation or a different sharding of the ensemble has to reproduce it bit for bit); all lanes of
        // one trajectory share h and the period, hence the decision
        small_step = !(fabs(wstep) > 0.5);
      }
      auto stage_rhs = [&](double t, double ci, D (&out)[NE]) {  // inner stage at t = tprev + ci*h
        if constexpr (SEASONAL) {
          double sn, cs;
          if (small_step) {
            double sd, cd;
            sincos_small(ci * wstep, sd, cd);
            sn = fma(sn0, cd, cs0 * sd);
            cs = fma(cs0, cd, -(sn0 * sd));
          } else {
            season_angle(t, pl, sn, cs);
          }
          rhs_sc(sn, cs, ys, out, c, Kl, pl, invN);
        } else {
          rhs(t, ys, out, c, Kl, pl, invN);
        }
      };
      auto end_rhs = [&](D (&out)[NE]) {  // the two stages at t = tnext
        if constexpr (SEASONAL) {
          rhs_sc(sn1, cs1, ys, out, c, Kl, pl, invN);
        } else {
          rhs(tnext, ys, out, c, Kl, pl, invN);
        }
      };
      // ---- Tsit5 stages 2..7 (6 new RHS evaluations; stage 7 = y1 (SSAL) and next f0 (FSAL))
#pragma unroll
      for (int e = 0; e < NE; ++e) ys[e] = dfma(h, T5_a21 * f[0][e], y[e]);
      stage_rhs(fma(T5_c2, h, tprev), T5_c2, f[1]);
#pragma unroll
      for (int e = 0; e < NE; ++e) ys[e] = dfma(h, dfma(T5_a32, f[1][e], T5_a31 * f[0][e]), y[e]);
      stage_rhs(fma(T5_c3, h, tprev), T5_c3, f[2]);
#pragma unroll
      for (int e = 0; e < NE; ++e)
        ys[e] = dfma(h, dfma(T5_a43, f[2][e], dfma(T5_a42, f[1][e], T5_a41 * f[0][e])), y[e]);
      stage_rhs(fma(T5_c4, h, tprev), T5_c4, f[3]);
#pragma unroll
      for (int e = 0; e < NE; ++e)
        ys[e] = dfma(h, dfma(T5_a54, f[3][e], dfma(T5_a53, f[2][e], dfma(T5_a52, f[1][e], T5_a51 * f[0][e]))), y[e]);
      stage_rhs(fma(T5_c5, h, tprev), T5_c5, f[4]);
#pragma unroll
      for (int e = 0; e < NE; ++e)
        ys[e] = dfma(h, dfma(T5_a65, f[4][e], dfma(T5_a64, f[3][e], dfma(T5_a63, f[2][e],
                     dfma(T5_a62, f[1][e], T5_a61 * f[0][e])))), y[e]);
      end_rhs(f[5]);
#pragma unroll
      for (int e = 0; e < NE; ++e)
        ys[e] = dfma(h, dfma(T5_a76, f[5][e], dfma(T5_a75, f[4][e], dfma(T5_a74, f[3][e], dfma(T5_a73, f[2][e],
                     dfma(T5_a72, f[1][e], T5_a71 * f[0][e]))))), y[e]);
      end_rhs(f[6]);  // ys is y1

      // ---- dense-output coefficients, formed unconditionally right after the last stage so their
      // independent FMAs overlap the latency-bound error-norm / controller chain below.
      // Monomial form of the Tsit5 interpolant (tsit5.cuh kDense):
      //   y(th) = y + (h th w11) f1 + (h th^2) (Q2 + th (Q3 + th Q4)),  Q_m = sum_i w_im f_i
      // The fused log-likelihood reads ONE compartment: its coefficients, y and f1 are gathered here through the
      // warp-uniform branch chain (picked through an index the compiler cannot see -- written as a dynamically
      // indexed read it put y, f and Q in local memory), so that neither the other compartments' 21 FMAs each
      // (x (1 + P) with tangents) nor the chain itself are paid per step / per save pass.
      D Q[3][IS_SAVE ? NE : 1];
      D y_obs = make_dual<P>(0.0), f_obs = make_dual<P>(0.0);
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        bool wanted = true;
        if constexpr (!IS_SAVE) {
          int ee = e;
          asm volatile("" : "+r"(ee));
          wanted = ee == a.obs_comp;
        }
        if (wanted) {
#pragma unroll
          for (int m = 0; m < 3; ++m) {
            D acc = kDense[0][m + 1] * f[0][e];
#pragma unroll
            for (int i = 1; i < 7; ++i) acc = dfma(kDense[i][m + 1], f[i][e], acc);
            if constexpr (OFFLOAD_Q) {
              my[(OFF_Q + m * NE + e) * kThreads] = acc.v;
            } else {
              Q[m][IS_SAVE ? e : 0] = acc;
            }
          }
          if constexpr (!IS_SAVE) { y_obs = y[e]; f_obs = f[0][e]; }
        }
      }

      // ---- embedded error, scaled RMS norm over the whole state (PIDController.adapt_step_size).
      // A NaN error estimate propagates to err and is rejected with factor 0.2, which is what
      // diffrax's NaN->inf substitution yields.
      bool keep;
      double dt_next;
      if (a.const_dt > 0.0) {
        keep = true;
        dt_next = a.const_dt;
      } else {
        double part = 0.0;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          double er = T5_e7 * f[6][e].v;
          er = fma(T5_e6, f[5][e].v, er);
          er = fma(T5_e5, f[4][e].v, er);
          er = fma(T5_e4, f[3][e].v, er);
          er = fma(T5_e3, f[2][e].v, er);
          er = fma(T5_e2, f[1][e].v, er);
          er = fma(T5_e1, f[0][e].v, er);
          er *= h;
          const double sc = fma(fmax(fabs(y[e].v), fabs(ys[e].v)), rtol, atol);
          const double r = er * rcp_fast1(sc);  // 2^-46 relative: only the RMS norm sees it
          if (e == 0) part = lead ? r * r : 0.0; else part = fma(r, r, part);
        }
        const double err2 = traj_sum(part, c) * inv_n;  // err^2 (RMS norm squared)
        keep = err2 < 1.0;                               // err < 1
        dt_next = h * controller_factor_sq(err2, keep);
      }
      double ntprev = keep ? tnext : tprev;
      bool next_made_jump = false;
      if constexpr (JUMPS) {
        // a kept step that ended just before a jump restarts exactly at the jump (nextafter), see a8
        if (keep && made_jump) ntprev = nextafter(tnext, CUDART_INF);
      }
      double ntnext = ntprev + dt_next;
      if constexpr (JUMPS) ntnext = clip_to_jumps(a, ntprev, ntnext, next_made_jump);
      ntprev = fmin(ntprev, t1);
      if (ntnext > t1 - 1e-10) ntnext = keep ? t1 : fma(0.5, t1 - ntprev, ntprev);  // _clip_to_end

      // ---- SaveAt(ts): dense output over [tprev, tnext] for every ts[k] <= tnext
      const bool do_save = stepping && keep;
      double ts_next = do_save ? save_time(save_i) : CUDART_INF;
      if (__any_sync(0xffffffffu, ts_next <= tnext)) {
        const double inv_h = rcp_fast((tnext == tprev) ? 1.0 : h);
        const double hw = h * kDense[0][0];
        // offloaded kernels bring Q back into registers ONCE per step (the stage derivatives f1..f5 are
        // dead by now); read inside the save loop it was re-fetched after every global store (12 LDS per pass,
        // 9 passes per step for the 32-slot warps: profiles/r1/ncu_full_c3_lane_solver.md)
        D Qs[3][NE];
        if constexpr (OFFLOAD_Q) {
#pragma unroll
          for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int e = 0; e < NE; ++e) Qs[m][e] = make_dual<P>(my[(OFF_Q + m * NE + e) * kThreads]);
        }
        auto dense = [&](int e, double th, double hthw, double hth2) -> D {
          D q0, q1, q2;
          if constexpr (OFFLOAD_Q) {
            q0 = Qs[0][e]; q1 = Qs[1][e]; q2 = Qs[2][e];
          } else {
            q0 = Q[0][IS_SAVE ? e : 0]; q1 = Q[1][IS_SAVE ? e : 0]; q2 = Q[2][IS_SAVE ? e : 0];
          }
          D u = dfma(th, q2, q1);
          u = dfma(th, u, q0);
          return dfma(hth2, u, dfma(hthw, f[0][e], y[e]));
        };
        if (IS_SAVE && P == 0 && full_save) {
          // fast path: every compartment saved -> compile-time offsets off the trajectory's two row pointers.
          // One pass of this loop is the unit the warp pays for every slot whenever ANY slot has a save
          // pending (profiles/r1/lockstep_probe.md), so it is kept short: row address = base + save_i * N (one
          // IMAD.WIDE each, no loop-carried pointers), pending-ness evaluated once per pass, and the
          // uniform-grid / loaded-grid choice made outside the loop.
          auto passes = [&](auto next_time) {
            bool pend = ts_next <= tnext;
            do {
              if (pend) {
                // the time of the save after this one decides whether the loop goes on: it is the loop-carried
                // chain (constant loads, I2F, FMA, compare, vote -- the top stall site of the r1 captures), so it is
                // issued first and overlaps the dense evaluation
                const double th = (ts_next - tprev) * inv_h;
                ts_next = next_time(save_i + 1);
                const double hthw = hw * th, hth2 = (h * th) * th;
                const D v0 = dense(0, th, hthw, hth2);
                if constexpr (L == 1 && NE == 4) {
                  // one lane owns the whole 32-byte row (s, e, i, r): two 16-byte stores instead of four 8-byte ones
                  // halve the store wavefronts of a pass whose 32 lanes hit 32 different sectors (ncu r2: the pass's
                  // entry branch waited on the previous pass's stores for 9 % of all stall samples)
                  double2* const pr = reinterpret_cast<double2*>(row_ptr(out_s, save_i));
                  pr[0] = make_double2(v0.v, dense(1, th, hthw, hth2).v);
                  pr[1] = make_double2(dense(2, th, hthw, hth2).v, dense(3, th, hthw, hth2).v);
                } else {
                  if (lead) *row_ptr(out_s, save_i) = v0.v;
                  double* const pc = row_ptr(out_c, save_i);
#pragma unroll
                  for (int e = 1; e < NE; ++e) pc[(e - 1) * G * S] = dense(e, th, hthw, hth2).v;
                }
                ++save_i;
              }
              pend = ts_next <= tnext;
            } while (__any_sync(0xffffffffu, pend));
          };
          if (a.save_dt > 0.0) {
            // t0 + k*dt exceeds t1 >= tnext for k >= T, so "no save left" needs no case of its own
            passes([&](int k) -> double { return (k == a.T - 1) ? t1 : fma((double)k, a.save_dt, a.t0); });
          } else {
            passes([&](int k) -> double { return (k < a.T) ? __ldg(a.save_ts + k) : CUDART_INF; });
          }
        } else {
          // general path (partial save masks, tangents, the fused log-likelihood): same loop shape as above
          auto passes = [&](auto next_time) {
            bool pend = ts_next <= tnext;
            do {
              if (pend) {
                const double th = (ts_next - tprev) * inv_h;
                ts_next = next_time(save_i + 1);  // issued early, see the fast path
                const double hthw = hw * th, hth2 = (h * th) * th;
                if constexpr (IS_SAVE) {
                  const int64_t row = (traj * a.T + save_i) * (int64_t)n_saved;
#pragma unroll
                  for (int e = 0; e < NE; ++e) {
                    const D v = dense(e, th, hthw, hth2);
                    if (off_save[e] >= 0) {
                      if (wp) a.ys[row + off_save[e]] = v.v;
                      if constexpr (P > 0) {
#pragma unroll
                        for (int k = 0; k < P; ++k)
                          if (p0s + k < a.P_total) a.dys[(row + off_save[e]) * a.P_total + p0s + k] = v.d[k];
                      }
                    }
                  }
                } else {
                  // Poisson(max(diff(comp), 1e-6)).log_prob(obs) accumulated on the fly
                  // (examples/sir_infer_parameters.py:30-38); lgamma(obs+1) arrives in lp_const.
                  D u = dfma(th, Q[2][0], Q[1][0]);
                  u = dfma(th, u, Q[0][0]);
                  const D v = dfma(hth2, u, dfma(hthw, f_obs, y_obs));  // the observed compartment at ts_next
                  if (save_i > 0 && obs_owner) {
                    D inc = v - obs_prev;
                    const double o = __ldg(a.obs + (int64_t)(save_i - 1) * obs_m + obs_q);
                    if (inc.v > kLikC[0]) {
                      const D lg = dual_log(inc);
                      lp_acc = lp_acc + (o * lg - inc);
                    } else {
                      lp_acc.v += fma(o, kLikC[1], -kLikC[0]);  // clamped at 1e-6: zero gradient (jnp.maximum)
                    }
                  }
                  obs_prev = v;
                }
                ++save_i;
              }
              pend = ts_next <= tnext;
            } while (__any_sync(0xffffffffu, pend));
          };
          if (a.save_dt > 0.0) {
            passes([&](int k) -> double { return (k == a.T - 1) ? t1 : fma((double)k, a.save_dt, a.t0); });
          } else {
            passes([&](int k) -> double { return (k < a.T) ? __ldg(a.save_ts + k) : CUDART_INF; });
          }
        }
      }

      // ---- commit
      if (stepping) {
        ++n_steps;
        if (keep) {
          ++n_acc;
#pragma unroll
          for (int e = 0; e < NE; ++e) { y[e] = ys[e]; f[0][e] = f[6][e]; }
          if constexpr (SEASONAL) { sn0 = sn1; cs0 = cs1; }
        } else {
          ++n_rej;
        }
        tprev = ntprev;
        tnext = ntnext;
      }
      if constexpr (JUMPS) {
        // no FSAL across a discontinuity: f0 is re-evaluated at (jump, y1) for the slots that crossed one
        const bool crossed = stepping && keep && made_jump;
        if (__any_sync(0xffffffffu, crossed)) {
          D fj[NE];
          const D invNj = inv_population(y, c);
          rhs(tprev, y, fj, c, Kl, pl, invNj);
          if (crossed) {
#pragma unroll
            for (int e = 0; e < NE; ++e) f[0][e] = fj[e];
          }
        }
        if (stepping) made_jump = next_made_jump;
      }
      const bool done = !((tprev < t1) && (n_steps < a.max_steps));
      if constexpr (PERSIST) {
        retire(active && done);
      } else {
        active = active && !done;
      }
    }
    if constexpr (!PERSIST) retire(had_traj);
  }
};

// Resident CTAs per SM asked of ptxas (64-thread CTAs).  Measured on B200 (profiles/r1_tuning.md):
// 5-element lanes (SEIRS+C) run best at 6 CTAs = 12 warps with 168 registers and no spills; forcing
// 128 registers (16 warps) spills and loses 14%.  3/4-element lanes fit 128 registers -> 16 warps.
// Tangent-carrying kernels keep every register they can get.
constexpr int min_blocks(int flow, int p) {
#ifdef DYN_MINBLOCKS
  return DYN_MINBLOCKS;
#else
#ifdef DYN_MINBLOCKS_TANGENT
  return p > 0 ? DYN_MINBLOCKS_TANGENT : (flow == DYNODE_FLOW_SEIRS_C ? 6 : 8);
#else
#ifdef DYN_MINBLOCKS_P1
  return p == 1 ? DYN_MINBLOCKS_P1 : (p > 0 ? 1 : (flow == DYNODE_FLOW_SEIRS_C ? 6 : 8));
#else
  return p > 0 ? 1 : (flow == DYNODE_FLOW_SEIRS_C ? 6 : 8);
#endif
#endif
#endif
}
template <int FLOW, int FLAGS, int G, int S, int P, int MODE>
__global__ void __launch_bounds__(kThreads, min_blocks(FLOW, P)) lane_solver_kernel(const SolveArgs a) {
  LaneSolver<FLOW, FLAGS, G, S, P, MODE>::run(a);
}

// Host launcher.  The grid is sized to the number of warps the GPU keeps resident (occupancy x SMs,
// times a small oversubscription that evens out the tail); every warp integrates a contiguous
// chunk of ceil(B / warps) trajectories through its 32/(G*S) slots.
template <int FLOW, int FLAGS, int G, int S, int P, int MODE>
cudaError_t launch_lane_solver(const SolveArgs& a_in, cudaStream_t stream) {
  using LS = LaneSolver<FLOW, FLAGS, G, S, P, MODE>;
  auto kern = lane_solver_kernel<FLOW, FLAGS, G, S, P, MODE>;
  if (a_in.B <= 0) return cudaSuccess;
  static thread_local int cached_dev = -1, cached_warps = 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev != cached_dev) {
    int sms = 0, per_sm = 0;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0)) != cudaSuccess) return e;
    cached_warps = sms * (per_sm > 0 ? per_sm : 1) * (kThreads / 32);
    cached_dev = dev;
  }
  constexpr int wpc = kThreads / 32;
  SolveArgs a = a_in;
  // non-persistent instances: one generation per warp (chunk = TPW)
  int64_t warps = LS::PERSIST ? (int64_t)cached_warps * kOversubscribe : (int64_t)1 << 40;
  const int64_t min_chunk = LS::TPW;                               // at least one trajectory per slot
  if (a.n_pass < 1) a.n_pass = 1;
  const int64_t Bv = a.B * a.n_pass;
  const int64_t max_warps = (Bv + min_chunk - 1) / min_chunk;
  if (warps > max_warps) warps = max_warps;
  a.chunk = (Bv + warps - 1) / warps;
  warps = (Bv + a.chunk - 1) / a.chunk;
  const int64_t grid = (warps + wpc - 1) / wpc;
  // tuning knob: DYNODE_REFILL=<n> overrides the refill threshold of the persistent-slot instances
  static const int env_refill = [] { const char* v = getenv("DYNODE_REFILL"); return v ? atoi(v) : 0; }();
  a.refill_min = env_refill > 0 ? (env_refill < LS::TPW ? env_refill : LS::TPW) : LS::REFILL;
  // tuning knob: DYNODE_DEBUG_SMEM=<bytes> of unused dynamic shared memory per CTA lowers the number of
  // resident CTAs (occupancy experiments, profiles/r1/occupancy.md); 0 in production
  static const int debug_smem = [] { const char* v = getenv("DYNODE_DEBUG_SMEM"); return v ? atoi(v) : 0; }();
  if (debug_smem > 48 * 1024)
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, debug_smem);
  kern<<<(unsigned)grid, kThreads, debug_smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace dynode
