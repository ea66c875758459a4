// adjoint_solver.cuh -- fused Poisson-incidence log-likelihood with its gradient w.r.t. EVERY rate and the
// initial state by a discrete adjoint (reverse sweep over the accepted Tsit5 steps).
//
// The reference differentiates diffeqsolve in reverse mode (diffrax RecursiveCheckpointAdjoint, implicit in
// reference src/dynode/simulation/odes.py:133-144): the exact derivative of the discrete
// scheme with the step sequence frozen.  Forward sensitivities (lane_solver.cuh, P > 0) cost one extra solve
// per `tangent_chunk` directions; this kernel costs ~4.5 solves whatever the number of parameters, so the
// host picks it once more than a few directions are asked for (engine.poisson_loglik_grad).
//
// One trajectory per lane group exactly as in lane_solver.cuh (same lane geometry, same RHS code).
//   forward sweep : Tsit5 + I-controller; every accepted step checkpoints (tprev, tnext, y_k) to global
//                   scratch, every save time stores the observed compartment's value;
//   cotangents    : lp = sum_s obs_s*log(inc_s) - inc_s, inc_s = max(v_s - v_{s-1}, 1e-6)  ->  d lp / d v_s;
//   reverse sweep : for each step, last to first: recompute the 7 stages from the checkpoint, seed the stage
//                   cotangents with the dense-output weights of the saves that fall into the step, and pull
//                   them back through the stages with the hand-written vector-Jacobian product of the flow
//                   family (transposed contact contraction by warp shuffles), accumulating d lp/d rates.
#pragma once
#include <type_traits>

#include "lane_solver.cuh"

namespace dynode {

// JUMPS: SolverParams.discontinuity_points (ClipStepSizeController) in the forward sweep, as in
// lane_solver.cuh: a step that would contain a point ends just before it, the next starts exactly at it with a fresh
// f0.  The reverse sweep needs nothing for it: it rebuilds every step from its checkpoint (tprev, tnext, y_k) and
// recomputes f0 there anyway, and y is carried unchanged across the one-ulp gap.
template <int FLOW, int FLAGS, int G, int S, bool JUMPS = false>
struct AdjointSolver {
  using LS = LaneSolver<FLOW, FLAGS, G, S, 0, MODE_LOGLIK>;
  using D = typename LS::D;
  using Geo = typename LS::Geo;
  using Prm = typename LS::Prm;
  static constexpr int NE = LS::NE, L = LS::L, TPW = LS::TPW, N = LS::N;
  static constexpr int IE = LS::IE, II = LS::II, IR = LS::IR, IC = LS::IC;
  static constexpr bool HAS_E = LS::HAS_E, HAS_C = LS::HAS_C, WANING = LS::WANING, SEASONAL = LS::SEASONAL,
                        DENSITY = LS::DENSITY;

  struct ParamGrad {
    double beta, gamma, sigma, omega, amp, phase;
  };

  // Vector-Jacobian product of the RHS at (t, y): given cotangents m[] of dy[], returns cotangents cy[] of y[]
  // (cy[0] replicated over the strain lanes like S_g itself), adds the rate gradients to pg and the cotangent
  // of 1/N_g (replicated) to c_invN.  K = contact[g][:], Kc = contact[:][g].
  static DYN_DI void rhs_vjp(double t, const D (&y)[NE], const double (&m)[NE], double (&cy)[NE], const Geo& c,
                             const double (&K)[G], const double (&Kc)[G], const Prm& p, double invN,
                             ParamGrad& pg, double& c_invN) {
    const double I = y[II].v, Sg = y[0].v;
    const double prop = DENSITY ? I : I * invN;
    double acc;
    if constexpr (G == 1) {
      acc = K[0] * prop;
    } else {
      acc = K[0] * __shfl_sync(0xffffffffu, prop, c.base + c.s);
#pragma unroll
      for (int b = 1; b < G; ++b) acc = fma(K[b], __shfl_sync(0xffffffffu, prop, c.base + b * S + c.s), acc);
    }
    double seas = 1.0, sn = 0.0, cs = 0.0;
    if constexpr (SEASONAL) {
      const double arg = ((2.0 * CUDART_PI) * t) / p.period + p.phase.v;
      sincos(arg, &sn, &cs);
      seas = fma(p.amp.v, sn, 1.0);
    }
    const double beta_t = p.beta.v * seas;
    const double foi = beta_t * acc;
    // cotangent of newinf: it enters dS (-), dE or dI (+) and dC (+)
    double w = -m[0] + (HAS_E ? m[IE] : m[II]);
    if constexpr (HAS_C) w += m[IC];
    cy[0] = LS::sum_strains(foi * w, c);
    const double cfoi = w * Sg;
    const double cbeta_t = cfoi * acc;
    pg.beta = fma(cbeta_t, seas, pg.beta);
    if constexpr (SEASONAL) {
      const double cseas = cbeta_t * p.beta.v;
      pg.amp = fma(cseas, sn, pg.amp);
      pg.phase = fma(cseas * p.amp.v, cs, pg.phase);
    }
    const double cacc = cfoi * beta_t;
    // transposed contact contraction: lane (b, s) collects K[g][b] * cacc(g, s) over g
    double cprop;
    if constexpr (G == 1) {
      cprop = Kc[0] * cacc;
    } else {
      cprop = Kc[0] * __shfl_sync(0xffffffffu, cacc, c.base + c.s);
#pragma unroll
      for (int g2 = 1; g2 < G; ++g2) cprop = fma(Kc[g2], __shfl_sync(0xffffffffu, cacc, c.base + g2 * S + c.s), cprop);
    }
    double cI;
    if constexpr (DENSITY) {
      cI = cprop;
    } else {
      cI = cprop * invN;
      c_invN += LS::sum_strains(cprop * I, c);
    }
    const double crec = m[IR] - m[II];  // rec = gamma*I enters dI (-) and dR (+)
    cI = fma(p.gamma.v, crec, cI);
    pg.gamma = fma(I, crec, pg.gamma);
    cy[II] = cI;
    if constexpr (WANING) {
      const double cw = m[0] - m[IR];  // omega*R enters dS (+) and dR (-)
      cy[IR] = p.omega.v * cw;
      pg.omega = fma(y[IR].v, cw, pg.omega);
    } else {
      cy[IR] = 0.0;
    }
    if constexpr (HAS_E) {
      const double ce = m[II] - m[IE];  // sigma*E enters dE (-) and dI (+)
      cy[IE] = p.sigma.v * ce;
      pg.sigma = fma(y[IE].v, ce, pg.sigma);
    }
    if constexpr (HAS_C) cy[IC] = 0.0;
  }

  // Tsit5 dense-output weights b_i(theta) from the monomial table (same numbers the forward sweep uses)
  static DYN_DI void dense_b(double th, double (&b)[7]) {
    using namespace tsit5;
#pragma unroll
    for (int i = 0; i < 7; ++i)
      b[i] = th * fma(th, fma(th, fma(th, kDense[i][3], kDense[i][2]), kDense[i][1]), kDense[i][0]);
  }

  static __device__ void run(const AdjointArgs& aa) {
    using namespace tsit5;
    const SolveArgs& a = aa.s;
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int tw = lane / L;
    const int q = lane - tw * L;
    Geo c;
    c.g = q / S;
    c.s = q - c.g * S;
    c.base = tw * L;
    c.sbase = c.base + c.g * S;
    double K[G], Kc[G];
#pragma unroll
    for (int b = 0; b < G; ++b) {
      K[b] = a.prm.contact ? __ldg(a.prm.contact + c.g * G + b) : (b == c.g ? 1.0 : 0.0);
      Kc[b] = a.prm.contact ? __ldg(a.prm.contact + b * G + c.g) : (b == c.g ? 1.0 : 0.0);
    }
    const bool lead = c.s == 0;
    const int64_t traj = warp_global * TPW + tw;
    const bool have = tw < TPW && traj < a.B && (a.only == nullptr || __ldg(a.only + traj) != 0);  // row mask
    if (!__any_sync(0xffffffffu, have)) return;
    const int64_t tr = have ? traj : 0;  // idle lanes shadow trajectory 0 (nothing of theirs is stored)

    int off_full[NE];
    off_full[0] = c.g;
#pragma unroll
    for (int e = 1; e < NE; ++e) off_full[e] = G + (e - 1) * G * S + c.g * S + c.s;
    const double t1 = a.t1, rtol = a.rtol, atol = a.atol;
    const double inv_n = 1.0 / (double)N;
    const int obs_m = (a.obs_comp == 0) ? G : G * S;
    const int obs_q = (a.obs_comp == 0) ? c.g : c.g * S + c.s;
    const bool obs_owner = have && ((a.obs_comp != 0) || lead);
    double* const vs = aa.vsave + (tr * a.T) * (int64_t)obs_m + obs_q;      // my observed element, stride obs_m
    double* const ck = aa.ckpt + tr * (int64_t)aa.cap * (N + 2);            // my trajectory's checkpoints
    auto save_time = [&](int k) -> double {
      if (k >= a.T) return CUDART_INF;
      if (a.save_dt > 0.0) return (k == a.T - 1) ? a.t1 : fma((double)k, a.save_dt, a.t0);
      return __ldg(a.save_ts + k);
    };

    // ---- parameters, initial state
    auto ld = [&](const DynodeArray& arr, int k, double dflt) -> double {
      return arr.ptr ? __ldg(arr.ptr + tr * arr.batch_stride + k) : dflt;
    };
    Prm prm;
    prm.beta.v = ld(a.prm.beta, c.s, 0.0);
    prm.gamma.v = ld(a.prm.gamma, c.s, 0.0);
    prm.sigma.v = HAS_E ? ld(a.prm.sigma, c.s, 0.0) : 0.0;
    prm.omega.v = WANING ? ld(a.prm.omega, c.s, 0.0) : 0.0;
    prm.amp.v = SEASONAL ? ld(a.prm.season_amp, 0, 0.0) : 0.0;
    prm.phase.v = SEASONAL ? ld(a.prm.season_phase, 0, 0.0) : 0.0;
    prm.period = SEASONAL ? ld(a.prm.season_period, 0, 1.0) : 1.0;
    D y[NE], f[7][NE], ys[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) y[e].v = __ldg(a.y0.ptr + tr * a.y0.batch_stride + off_full[e]);

    // the 7 stages of one step from (tprev, tnext, y): fills f[0..6] and ys = y_{k+1}
    auto stages = [&](double tprev, double tnext, const D& invN, bool have_f0) {
      const double h = tnext - tprev;
      if (!have_f0) LS::rhs(tprev, y, f[0], c, K, prm, invN);
#pragma unroll
      for (int e = 0; e < NE; ++e) ys[e] = dfma(h, T5_a21 * f[0][e], y[e]);
      LS::rhs(fma(T5_c2, h, tprev), ys, f[1], c, K, prm, invN);
#pragma unroll
      for (int e = 0; e < NE; ++e) ys[e] = dfma(h, dfma(T5_a32, f[1][e], T5_a31 * f[0][e]), y[e]);
      LS::rhs(fma(T5_c3, h, tprev), ys, f[2], c, K, prm, invN);
#pragma unroll
      for (int e = 0; e < NE; ++e)
        ys[e] = dfma(h, dfma(T5_a43, f[2][e], dfma(T5_a42, f[1][e], T5_a41 * f[0][e])), y[e]);
      LS::rhs(fma(T5_c4, h, tprev), ys, f[3], c, K, prm, invN);
#pragma unroll
      for (int e = 0; e < NE; ++e)
        ys[e] = dfma(h, dfma(T5_a54, f[3][e], dfma(T5_a53, f[2][e], dfma(T5_a52, f[1][e], T5_a51 * f[0][e]))), y[e]);
      LS::rhs(fma(T5_c5, h, tprev), ys, f[4], c, K, prm, invN);
#pragma unroll
      for (int e = 0; e < NE; ++e)
        ys[e] = dfma(h, dfma(T5_a65, f[4][e], dfma(T5_a64, f[3][e], dfma(T5_a63, f[2][e],
                     dfma(T5_a62, f[1][e], T5_a61 * f[0][e])))), y[e]);
      LS::rhs(tnext, ys, f[5], c, K, prm, invN);
#pragma unroll
      for (int e = 0; e < NE; ++e)
        ys[e] = dfma(h, dfma(T5_a76, f[5][e], dfma(T5_a75, f[4][e], dfma(T5_a74, f[3][e], dfma(T5_a73, f[2][e],
                     dfma(T5_a72, f[1][e], T5_a71 * f[0][e]))))), y[e]);
      LS::rhs(tnext, ys, f[6], c, K, prm, invN);
    };

    // =============================== forward sweep ===============================
    double tprev = a.t0, tnext;
    bool made_jump = false;  // the running step was clipped to end just before a discontinuity point
    {
      const D invN0 = LS::inv_population(y, c);
      LS::rhs(a.t0, y, f[0], c, K, prm, invN0);
      if (a.const_dt > 0.0) {
        tnext = a.t0 + a.const_dt;
      } else {  // Hairer-Wanner initial step (same arithmetic as lane_solver.cuh)
        double p0 = 0.0, p1 = 0.0, scale[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          scale[e] = fma(fabs(y[e].v), rtol, atol);
          const double wgt = (e == 0 && !lead) ? 0.0 : 1.0;
          p0 += wgt * LS::sq(y[e].v / scale[e]);
          p1 += wgt * LS::sq(f[0][e].v / scale[e]);
        }
        const double d0 = sqrt(LS::traj_sum(p0, c) * inv_n);
        const double d1 = sqrt(LS::traj_sum(p1, c) * inv_n);
        const bool small = (d0 < 1e-5) || (d1 < 1e-5);
        const double h0 = small ? 1e-6 : 0.01 * (d0 / d1);
#pragma unroll
        for (int e = 0; e < NE; ++e) ys[e] = dfma(h0, f[0][e], y[e]);
        LS::rhs(a.t0 + h0, ys, f[1], c, K, prm, invN0);
        double p2 = 0.0;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          const double wgt = (e == 0 && !lead) ? 0.0 : 1.0;
          p2 += wgt * LS::sq((f[1][e].v - f[0][e].v) / scale[e]);
        }
        const double d2 = sqrt(LS::traj_sum(p2, c) * inv_n) / h0;
        const double md = fmax(d1, d2);
        const double h1 = (md <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / md, 0.2);
        tnext = a.t0 + fmin(100.0 * h0, h1);
      }
      if constexpr (JUMPS) tnext = LS::clip_to_jumps(a, a.t0, tnext, made_jump);
      tnext = fmin(tnext, t1);
    }
    int32_t n_acc = 0, n_rej = 0, n_steps = 0, save_i = 0;
    bool active = have;
    while (__any_sync(0xffffffffu, active)) {
      const bool stepping = active && (tprev < t1) && (n_steps < a.max_steps);
      const double h = tnext - tprev;
      const D invN = LS::inv_population(y, c);
      stages(tprev, tnext, invN, /*have_f0=*/true);
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
          const double r = er * rcp_fast1(sc);
          if (e == 0) part = lead ? r * r : 0.0; else part = fma(r, r, part);
        }
        const double err2 = LS::traj_sum(part, c) * inv_n;
        keep = err2 < 1.0;
        dt_next = h * controller_factor_sq(err2, keep);
      }
      double ntprev = keep ? tnext : tprev;
      bool next_made_jump = false;
      if constexpr (JUMPS) {
        if (keep && made_jump) ntprev = nextafter(tnext, CUDART_INF);  // restart exactly at the jump
      }
      double ntnext = ntprev + dt_next;
      if constexpr (JUMPS) ntnext = LS::clip_to_jumps(a, ntprev, ntnext, next_made_jump);
      ntprev = fmin(ntprev, t1);
      if (ntnext > t1 - 1e-10) ntnext = keep ? t1 : fma(0.5, t1 - ntprev, ntprev);
      if (stepping && keep) {
        // checkpoint the step, then the observed compartment at every save time inside it
        if (n_acc < aa.cap) {
          double* row = ck + (int64_t)n_acc * (N + 2);
          if (q == 0) { row[0] = tprev; row[1] = tnext; }
#pragma unroll
          for (int e = 0; e < NE; ++e)
            if (e > 0 || lead) row[2 + off_full[e]] = y[e].v;
        }
        const double inv_h = rcp_fast((tnext == tprev) ? 1.0 : h);
        while (save_i < a.T && save_time(save_i) <= tnext) {
          const double th = (save_time(save_i) - tprev) * inv_h;
          double b[7];
          dense_b(th, b);
          double v = 0.0;
#pragma unroll
          for (int e = 0; e < NE; ++e) {
            int ee = e;
            asm volatile("" : "+r"(ee));
            if (ee == a.obs_comp) {
              double acc = b[0] * f[0][e].v;
#pragma unroll
              for (int i = 1; i < 7; ++i) acc = fma(b[i], f[i][e].v, acc);
              v = fma(h, acc, y[e].v);
            }
          }
          if (obs_owner) vs[(int64_t)save_i * obs_m] = v;
          ++save_i;
        }
      }
      if (stepping) {
        ++n_steps;
        if (keep) {
          ++n_acc;
#pragma unroll
          for (int e = 0; e < NE; ++e) { y[e] = ys[e]; f[0][e] = f[6][e]; }
        } else {
          ++n_rej;
        }
        tprev = ntprev;
        tnext = ntnext;
      }
      if constexpr (JUMPS) {
        // no FSAL across a discontinuity: f0 is re-evaluated at (jump, y1) for the trajectories that crossed one
        const bool crossed = stepping && keep && made_jump;
        if (__any_sync(0xffffffffu, crossed)) {
          const D invNj = LS::inv_population(y, c);
          LS::rhs(tprev, y, ys, c, K, prm, invNj);
          if (crossed) {
#pragma unroll
            for (int e = 0; e < NE; ++e) f[0][e] = ys[e];
          }
        }
        if (stepping) made_jump = next_made_jump;
      }
      active = active && (tprev < t1) && (n_steps < a.max_steps);
    }
    const bool complete = have && !(tprev < t1) && n_acc <= aa.cap && save_i >= a.T;

    // =============================== cotangents of the saved values ===============================
    // lp = sum_s obs_{s-1} log(inc_s) - inc_s, inc_s = max(v_s - v_{s-1}, 1e-6); vs[s] <- d lp / d v_s
    double lp_part = 0.0;
    if (obs_owner && complete) {
      double prev = vs[0], carry = 0.0;
      for (int s2 = 1; s2 < a.T; ++s2) {
        const double cur = vs[(int64_t)s2 * obs_m];
        const double inc = cur - prev;
        const double o = __ldg(a.obs + (int64_t)(s2 - 1) * obs_m + obs_q);
        double wgt = 0.0;
        if (inc > kLikC[0]) {
          lp_part += o * log_fast(inc) - inc;
          wgt = div_fast(o, inc) - 1.0;
        } else {
          lp_part += fma(o, kLikC[1], -kLikC[0]);
        }
        vs[(int64_t)(s2 - 1) * obs_m] = carry - wgt;
        carry = wgt;
        prev = cur;
      }
      vs[(int64_t)(a.T - 1) * obs_m] = carry;
    }
    const double lp_tot = LS::traj_sum(lp_part, c);
    if (have && q == 0) {
      a.lp[traj] = complete ? lp_tot + a.lp_const : CUDART_NAN;
      int32_t* st = a.stats + traj * 4;
      st[DYNODE_STAT_RESULT] = (tprev < t1) ? DYNODE_RESULT_MAX_STEPS : (n_acc > aa.cap ? DYNODE_RESULT_ADJOINT_CAPACITY : DYNODE_RESULT_OK);
      st[DYNODE_STAT_ACCEPTED] = n_acc;
      st[DYNODE_STAT_REJECTED] = n_rej;
      st[DYNODE_STAT_STEPS] = n_steps;
    }

    // =============================== reverse sweep ===============================
    double lam[NE];  // cotangent of y_{k+1}; lam[0] replicated over the strain lanes
#pragma unroll
    for (int e = 0; e < NE; ++e) lam[e] = 0.0;
    ParamGrad pg = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    int k = complete ? n_acc - 1 : -1;
    int sj = a.T - 1;  // next save (from the end) whose cotangent has not been consumed
    while (__any_sync(0xffffffffu, k >= 0)) {
      const bool on = k >= 0;
      const double* row = ck + (int64_t)(on ? k : 0) * (N + 2);
      const double tp = on ? row[0] : a.t0, tn = on ? row[1] : a.t0;
#pragma unroll
      for (int e = 0; e < NE; ++e) y[e].v = on ? row[2 + off_full[e]] : 1.0;
      const double h = tn - tp;
      const D invN = LS::inv_population(y, c);
      stages(tp, tn, invN, /*have_f0=*/false);  // f[0] = f(tp, y_k) recomputed (equals the FSAL value)
      // ---- stage cotangents mu_i of f_i: dense-output saves inside (tp, tn] (and ts == t0 for the first step)
      double mu[7][NE];
#pragma unroll
      for (int i = 0; i < 7; ++i)
#pragma unroll
        for (int e = 0; e < NE; ++e) mu[i][e] = 0.0;
      double cyk[NE];  // cotangent of y_k from the identity paths (saves, stage states)
#pragma unroll
      for (int e = 0; e < NE; ++e) cyk[e] = 0.0;
      const double inv_h = rcp_fast((tn == tp) ? 1.0 : h);
      while (true) {
        const bool mine = on && sj >= 0 && (save_time(sj) > tp || k == 0);
        if (!__any_sync(0xffffffffu, mine)) break;
        if (mine) {
          const double th = (save_time(sj) - tp) * inv_h;
          double b[7];
          dense_b(th, b);
          const double cv = obs_owner ? vs[(int64_t)sj * obs_m] : 0.0;
#pragma unroll
          for (int e = 0; e < NE; ++e) {
            int ee = e;
            asm volatile("" : "+r"(ee));
            if (ee == a.obs_comp) {
              cyk[e] += cv;
#pragma unroll
              for (int i = 0; i < 7; ++i) mu[i][e] = fma(h * b[i], cv, mu[i][e]);
            }
          }
          --sj;
        }
      }
      // S_g is replicated: a cotangent that arrived on the lead lane only is shared with its group
      if constexpr (S > 1) {
        cyk[0] = __shfl_sync(0xffffffffu, cyk[0], c.sbase);
#pragma unroll
        for (int i = 0; i < 7; ++i) mu[i][0] = __shfl_sync(0xffffffffu, mu[i][0], c.sbase);
      }
      // ---- y_{k+1} = Y_7 (the stage-7 state): its cotangent is lam plus the pull-back of mu_7
      double c_invN = 0.0, cY[NE];
      rhs_vjp(tn, ys, mu[6], cY, c, K, Kc, prm, invN.v, pg, c_invN);
#pragma unroll
      for (int e = 0; e < NE; ++e) cY[e] += lam[e];
      // Y_7 = y_k + h*sum_j a_7j f_j
      const double a7[6] = {T5_a71, T5_a72, T5_a73, T5_a74, T5_a75, T5_a76};
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        cyk[e] += cY[e];
#pragma unroll
        for (int j = 0; j < 6; ++j) mu[j][e] = fma(h * a7[j], cY[e], mu[j][e]);
      }
      // stages 6 .. 2: rebuild the stage state from the f's, pull mu_i back, spread over earlier stages
      auto back_stage = [&](auto idx, double ti, const double* arow) {
        constexpr int i = decltype(idx)::value;  // f index of the stage (stage i+1 of the tableau)
        D Y[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          double acc = arow[0] * f[0][e].v;
#pragma unroll
          for (int j = 1; j < i; ++j) acc = fma(arow[j], f[j][e].v, acc);
          Y[e].v = fma(h, acc, y[e].v);
        }
        rhs_vjp(ti, Y, mu[i], cY, c, K, Kc, prm, invN.v, pg, c_invN);
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          cyk[e] += cY[e];
#pragma unroll
          for (int j = 0; j < i; ++j) mu[j][e] = fma(h * arow[j], cY[e], mu[j][e]);
        }
      };
      {
        const double r6[5] = {T5_a61, T5_a62, T5_a63, T5_a64, T5_a65};
        back_stage(std::integral_constant<int, 5>{}, tn, r6);
        const double r5[4] = {T5_a51, T5_a52, T5_a53, T5_a54};
        back_stage(std::integral_constant<int, 4>{}, fma(T5_c5, h, tp), r5);
        const double r4[3] = {T5_a41, T5_a42, T5_a43};
        back_stage(std::integral_constant<int, 3>{}, fma(T5_c4, h, tp), r4);
        const double r3[2] = {T5_a31, T5_a32};
        back_stage(std::integral_constant<int, 2>{}, fma(T5_c3, h, tp), r3);
        const double r2[1] = {T5_a21};
        back_stage(std::integral_constant<int, 1>{}, fma(T5_c2, h, tp), r2);
      }
      // stage 1: Y_1 = y_k
      rhs_vjp(tp, y, mu[0], cY, c, K, Kc, prm, invN.v, pg, c_invN);
#pragma unroll
      for (int e = 0; e < NE; ++e) cyk[e] += cY[e];
      // 1/N_g was formed from y_k: N_g = S_g + sum_s (E + I + R)
      if constexpr (!DENSITY) {
        const double cN = -(invN.v * invN.v) * c_invN;
        cyk[0] += cN;
        cyk[II] += cN;
        cyk[IR] += cN;
        if constexpr (HAS_E) cyk[IE] += cN;
      }
      if (on) {
#pragma unroll
        for (int e = 0; e < NE; ++e) lam[e] = cyk[e];
        --k;
      }
    }

    // =============================== write the gradients ===============================
    const double gb = LS::sum_groups(pg.beta, c), gg = LS::sum_groups(pg.gamma, c);
    const double gs = LS::sum_groups(pg.sigma, c), go = LS::sum_groups(pg.omega, c);
    const double ga = LS::traj_sum(pg.amp, c), gp = LS::traj_sum(pg.phase, c);
    if (have) {
      double* gr = aa.grad + traj * (4 * S + 2);
      const double bad = complete ? 0.0 : CUDART_NAN;
      if (c.g == 0) {
        gr[0 * S + c.s] = gb + bad;
        gr[1 * S + c.s] = gg + bad;
        gr[2 * S + c.s] = gs + bad;
        gr[3 * S + c.s] = go + bad;
      }
      if (q == 0) {
        gr[4 * S + 0] = ga + bad;
        gr[4 * S + 1] = gp + bad;
      }
      if (aa.grad_y0) {
        double* g0 = aa.grad_y0 + traj * (int64_t)N;
#pragma unroll
        for (int e = 0; e < NE; ++e)
          if (e > 0 || lead) g0[off_full[e]] = lam[e] + bad;
      }
    }
  }
};

template <int FLOW, int FLAGS, int G, int S, bool JUMPS>
__global__ void __launch_bounds__(kThreads, 1) adjoint_solver_kernel(const AdjointArgs a) {
  AdjointSolver<FLOW, FLAGS, G, S, JUMPS>::run(a);
}

template <int FLOW, int FLAGS, int G, int S, bool JUMPS>
cudaError_t launch_adjoint_solver(const AdjointArgs& a, cudaStream_t stream) {
  using AS = AdjointSolver<FLOW, FLAGS, G, S, JUMPS>;
  if (a.s.B <= 0) return cudaSuccess;
  constexpr int wpc = kThreads / 32;
  const int64_t warps = (a.s.B + AS::TPW - 1) / AS::TPW;
  const int64_t grid = (warps + wpc - 1) / wpc;
  adjoint_solver_kernel<FLOW, FLAGS, G, S, JUMPS><<<(unsigned)grid, kThreads, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace dynode
