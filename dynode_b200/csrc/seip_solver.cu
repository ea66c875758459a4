// seip_solver.cu -- one thread block per trajectory for the immune-history / waning family
// (include/dynode_b200_seip.h).  State, stage state and the 7 stage derivatives live in shared memory
// (9*n doubles), each thread owns the elements e = tid, tid + blockDim, ...; the infectious totals per
// (age, strain), the force of infection through the contact matrix and the RMS error norm are block-wide
// steps separated by __syncthreads.  The integrator is the same restatement of
// diffeqsolve(Tsit5, PIDController(rtol, atol), SaveAt(ts)) as lane_solver.cuh (reference
// src/dynode/simulation/odes.py:107-144); the right-hand side follows
// oracle/dynode_oracle.cpp FAM_SEIP term by term, in the same summation order.
#include <cuda_runtime.h>
#include <math_constants.h>

#include <type_traits>

#include "../../include/dynode_b200.h"
#include "../../include/dynode_b200_seip.h"
#include "tsit5.cuh"

namespace dynode {

int fail_msg(const char* fmt, ...);  // capi.cu

namespace {

constexpr int kSeipThreads = 128;

struct SeipArgs {
  int A, K, W, H, n;
  int V, NK;  // vaccination tiers, spline knots
  int64_t B;
  DynodeArray y0, beta, sigma, gamma, omega;
  DynodeArray itime, iscale, ipct;  // external introductions per strain (ptr NULL = none)
  const double* contact;
  const double* pop;
  const double* imm;
  const double *vbase, *vknot, *vcoef;  // vaccination-rate splines [A][V][4 | NK | NK] (vbase NULL = no vaccination)
  const double* iages;                  // [K][A]
  double season_tau, season_on;
  double t0, t1, rtol, atol, const_dt, save_dt;
  const double* save_ts;
  int T;
  int max_steps;
  double* ys;
  int32_t* stats;
  const double* jump_ts;  // SolverParams.discontinuity_points, sorted, device memory (NULL / 0 = none)
  int n_jump;
  uint32_t save_mask;  // bit c = compartment c (S, E, I, C) is saved (sub_save_indices, reference odes.py:182-193)
  int n_saved;         // doubles per saved row
};

// ClipStepSizeController(jump_ts), as lane_solver.cuh: a step [t0, t1] that would contain a
// discontinuity point ends just before it; the next one restarts exactly at it with a fresh f0.
__device__ __forceinline__ double seip_clip_to_jumps(const SeipArgs& a, double t0, double t1, bool& made_jump) {
  int i0 = 0, i1 = 0;
  for (int k = 0; k < a.n_jump; ++k) {
    const double j = a.jump_ts[k];
    i0 += (j <= t0);
    i1 += (j <= t1);
  }
  made_jump = i0 < i1;
  return made_jump ? nextafter(a.jump_ts[i0 < a.n_jump ? i0 : a.n_jump - 1], -CUDART_INF) : t1;
}

struct Smem {
  double *ys, *dx;
  double *itot, *foi, *beta, *sigma, *gamma, *omega, *contact, *pop, *imm, *red;
  double *rate, *vbase, *vknot, *vcoef, *iages, *itime, *iscale, *ipct;
};

__device__ __forceinline__ double block_sum(double v, double* red) {
  // sum over the block, returned to every thread (4 warps)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5;
  __syncthreads();  // red[] may still be read from the previous reduction
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  double t = red[0];
  for (int k = 1; k < kSeipThreads / 32; ++k) t += red[k];
  return t;
}

// dx = f(t, x): all threads call it; x and dx are shared-memory arrays of n doubles.  KT / WT > 0 fix the number
// of strains / waning stages at compile time (loops unrolled, index arithmetic folded); 0 = runtime value.
// Term by term oracle/dynode_oracle.cpp FAM_SEIPV (V = 1 without splines / introductions is FAM_SEIP), same
// summation order.
// EXT = false compiles the vaccination tiers, splines, introductions and the seasonal reset out (V == 1, no tables):
// the lean right-hand side of the first version, at its speed.
template <int KT, int WT, bool EXT>
__device__ __forceinline__ void seip_rhs(const SeipArgs& a, const Smem& sm, double t, const double* x, double* dx) {
  const int A = a.A, K = KT ? KT : a.K, W = WT ? WT : a.W, H = KT ? (1 << KT) : a.H, V = EXT ? a.V : 1,
            NK = EXT ? a.NK : 0;
  const int nS = A * H * V * W, nX = A * H * V * K;
  const double* xS = x;
  const double* xE = x + nS;
  const double* xI = xE + nX;
  for (int q = threadIdx.x; q < A * K; q += blockDim.x) {
    const int ag = q / K, k = q - ag * K;
    double acc = 0.0;
    for (int jv = 0; jv < H * V; ++jv) acc += xI[(ag * H * V + jv) * K + k];
    double fr = acc / sm.pop[ag];  // infectious fraction of age group ag for strain k (divided once, not per target)
    if (EXT && a.ipct.ptr && sm.ipct[k] != 0.0) {  // external introductions: Gaussian in time (ode_model.md:183)
      const double zs = (t - sm.itime[k]) / sm.iscale[k];
      const double pdf = exp(-0.5 * zs * zs) / (sm.iscale[k] * 2.5066282746310002);
      fr += pdf * sm.ipct[k] * sm.iages[k * A + ag];
    }
    sm.itot[q] = fr;
  }
  if constexpr (EXT)
  for (int q = threadIdx.x; q < A * V; q += blockDim.x) {
    // vaccination rate out of tier v of age group ag: min(nu(t) pop / sum_{j,w} S, 1)  (ode_model.md:19-29)
    const int ag = q / V, v = q - ag * V;
    double r = 0.0;
    if (a.vbase) {
      const double* bs = sm.vbase + q * 4;
      double nu = bs[0] + bs[1] * t + bs[2] * t * t + bs[3] * t * t * t;
      for (int i = 0; i < NK; ++i) {
        const double d = t - sm.vknot[q * NK + i];
        if (d > 0.0) nu += sm.vcoef[q * NK + i] * d * d * d;
      }
      if (!(nu > 0.0)) nu = 0.0;
      double tot = 0.0;
      for (int j = 0; j < H; ++j)
        for (int w = 0; w < W; ++w) tot += xS[((ag * H + j) * V + v) * W + w];
      if (nu > 0.0 && tot > 0.0) {
        r = (nu * sm.pop[ag]) / tot;
        if (!(r < 1.0)) r = 1.0;
      }
    }
    sm.rate[q] = r;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < A * K; q += blockDim.x) {
    const int ag = q / K, k = q - ag * K;
    double acc = sm.contact[ag * A + 0] * sm.itot[0 * K + k];
    for (int b = 1; b < A; ++b) acc += sm.contact[ag * A + b] * sm.itot[b * K + k];
    sm.foi[q] = sm.beta[k] * acc;
  }
  __syncthreads();
  double phi = 0.0;  // seasonal reset of the top tier (ode_model.md:72-75)
  if (EXT && a.season_on != 0.0 && V >= 2) {
    const double sn = sin(2.0 * 3.14159265358979323846 * (t + a.season_tau) / 730.0);
    phi = a.season_on * pow(sn * sn, 500.0);
  }
  // one thread per (age, history, tier) cell group: the W x K exposure terms are formed once and feed dS (summed over
  // strains), dE and dC (summed over waning stages)
  constexpr int KMAX = KT ? KT : DYNODE_SEIP_MAX_STRAINS;
  double* dS = dx;
  double* dE = dx + nS;
  double* dI = dE + nX;
  double* dC = dI + nX;
  for (int cell = threadIdx.x; cell < A * H * V; cell += blockDim.x) {
    const int ag = cell / (H * V), jv = cell - ag * H * V, j = jv / V, v = jv - j * V;
    const bool top = v == V - 1;
    const double rv = EXT ? sm.rate[ag * V + v] : 0.0;
    double expo[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) expo[k] = 0.0;
    double vin = 0.0;  // vaccinated into this cell's stage 0: from the tier below, and boosters within the top tier
    if constexpr (EXT) {
    if (v >= 1) {
      double below = 0.0;
#pragma unroll
      for (int w = 0; w < W; ++w) below += xS[(cell - 1) * W + w];
      vin += sm.rate[ag * V + v - 1] * below;
    }
    if (top) {
      double older = 0.0;
#pragma unroll
      for (int w = 1; w < W; ++w) older += xS[cell * W + w];
      vin += rv * older;
    }
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const double s = xS[cell * W + w];
      double out = 0.0;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
          const double xk = sm.foi[ag * K + k] * (1.0 - sm.imm[(jv * W + w) * K + k]) * s;
          expo[k] += xk;
          out += xk;
        }
      }
      double d = -out;
      if (w > 0) d += sm.omega[w - 1] * xS[cell * W + w - 1];
      if (w < W - 1) d -= sm.omega[w] * s;
      if (w == 0) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K && ((j >> k) & 1))
            d += sm.gamma[k] * (xI[cell * K + k] + xI[((ag * H + (j ^ (1 << k))) * V + v) * K + k]);
      }
      if constexpr (EXT) {
        if (!(top && w == 0)) d -= rv * s;
        if (w == 0) d += vin;
        if (phi != 0.0) {
          if (top) d -= phi * s;
          if (v == V - 2) d += phi * xS[(cell + 1) * W + w];
        }
      }
      dS[cell * W + w] = d;
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < K) {
        const int q = cell * K + k;
        double de = expo[k] - sm.sigma[k] * xE[q];
        double di = sm.sigma[k] * xE[q] - sm.gamma[k] * xI[q];
        if (EXT && phi != 0.0) {
          if (top) { de -= phi * xE[q]; di -= phi * xI[q]; }
          if (v == V - 2) { de += phi * xE[q + K]; di += phi * xI[q + K]; }
        }
        dE[q] = de;
        dI[q] = di;
        dC[q] = expo[k];
      }
    }
  }
  __syncthreads();
}

// One thread block per trajectory.  Each thread owns the elements e = tid + 128 i (i < EPT) and keeps THEIR y, the
// stage state and the 7 stage derivatives in registers (9 EPT doubles); shared memory holds only what the block
// exchanges: the stage state the right-hand side reads (n doubles), its output (n doubles) and the tables.  With
// 2 n instead of 9 n doubles per trajectory an SM keeps 16 trajectories in flight instead of 7 (n = 416), the stage
// sums / error estimate / dense output run out of registers, and saves are coalesced straight from them.
template <int KT, int WT, bool EXT, int EPT>
__global__ void __launch_bounds__(kSeipThreads) seip_solver_kernel(const SeipArgs a) {
  using namespace tsit5;
  extern __shared__ double smem_raw[];
  const int A = a.A, K = KT ? KT : a.K, W = WT ? WT : a.W, H = KT ? (1 << KT) : a.H, n = a.n;
  Smem sm;
  double* p = smem_raw;
  sm.ys = p; p += n;   // stage state, read by the right-hand side
  sm.dx = p; p += n;   // its output
  sm.itot = p; p += A * K;
  sm.foi = p; p += A * K;
  sm.beta = p; p += K;
  sm.sigma = p; p += K;
  sm.gamma = p; p += K;
  sm.omega = p; p += W;
  sm.contact = p; p += A * A;
  sm.pop = p; p += A;
  sm.imm = p; p += H * a.V * W * K;
  sm.red = p; p += 8;
  sm.rate = p; p += A * a.V;
  sm.vbase = p; p += A * a.V * 4;
  sm.vknot = p; p += A * a.V * a.NK;
  sm.vcoef = p; p += A * a.V * a.NK;
  sm.iages = p; p += K * A;
  sm.itime = p; p += K;
  sm.iscale = p; p += K;
  sm.ipct = p; p += K;

  const int64_t traj = blockIdx.x;
  const int tid = threadIdx.x;
  // ---- stage the shared tables and this trajectory's rates
  for (int q = tid; q < K; q += blockDim.x) {
    sm.beta[q] = a.beta.ptr[traj * a.beta.batch_stride + q];
    sm.sigma[q] = a.sigma.ptr[traj * a.sigma.batch_stride + q];
    sm.gamma[q] = a.gamma.ptr[traj * a.gamma.batch_stride + q];
  }
  for (int q = tid; q < W; q += blockDim.x) sm.omega[q] = a.omega.ptr[traj * a.omega.batch_stride + q];
  for (int q = tid; q < A * A; q += blockDim.x) sm.contact[q] = a.contact[q];
  for (int q = tid; q < A; q += blockDim.x) sm.pop[q] = a.pop[q];
  for (int q = tid; q < H * a.V * W * K; q += blockDim.x) sm.imm[q] = a.imm[q];
  if (a.vbase) {
    for (int q = tid; q < A * a.V * 4; q += blockDim.x) sm.vbase[q] = a.vbase[q];
    for (int q = tid; q < A * a.V * a.NK; q += blockDim.x) { sm.vknot[q] = a.vknot[q]; sm.vcoef[q] = a.vcoef[q]; }
  }
  if (a.ipct.ptr) {
    for (int q = tid; q < K; q += blockDim.x) {
      sm.itime[q] = a.itime.ptr[traj * a.itime.batch_stride + q];
      sm.iscale[q] = a.iscale.ptr[traj * a.iscale.batch_stride + q];
      sm.ipct[q] = a.ipct.ptr[traj * a.ipct.batch_stride + q];
    }
    for (int q = tid; q < K * A; q += blockDim.x) sm.iages[q] = a.iages[q];
  }
  // ---- my elements: y, stage state, stage derivatives (registers)
  double y[EPT], yst[EPT], f[7][EPT];
  bool own[EPT];
  int soff[EXT ? EPT : 1];  // general kernel: my element's offset inside a saved row, -1 when its compartment is not saved
  const int nS_ = A * H * a.V * W, nX_ = A * H * a.V * K;
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
    const int e = tid + i * kSeipThreads;
    if constexpr (EXT) {  // the general kernel honours a compartment mask; the plain one saves whole rows
      const int comp = e < nS_ ? 0 : 1 + (e - nS_) / nX_;
      int skipped = 0;  // doubles of the unsaved compartments in front of mine
      if (comp > 0 && !(a.save_mask & 1u)) skipped += nS_;
      for (int c = 1; c < comp; ++c)
        if (!((a.save_mask >> c) & 1u)) skipped += nX_;
      soff[i] = (e < n && ((a.save_mask >> comp) & 1u)) ? e - skipped : -1;
    }
    own[i] = e < n;
    y[i] = own[i] ? a.y0.ptr[traj * a.y0.batch_stride + e] : 0.0;
    yst[i] = y[i];
#pragma unroll
    for (int j = 0; j < 7; ++j) f[j][i] = 0.0;
  }
  auto save_off = [&](int i) -> int {
    if constexpr (EXT) return soff[i]; else return own[i] ? tid + i * kSeipThreads : -1;
  };
  // right-hand side of the block's stage state `yst` at time t into f[slot]
  auto eval = [&](double t, double (&out)[EPT]) {
#pragma unroll
    for (int i = 0; i < EPT; ++i)
      if (own[i]) sm.ys[tid + i * kSeipThreads] = yst[i];
    __syncthreads();
    seip_rhs<KT, WT, EXT>(a, sm, t, sm.ys, sm.dx);  // ends with a barrier: dx is complete, ys free to overwrite
#pragma unroll
    for (int i = 0; i < EPT; ++i) out[i] = own[i] ? sm.dx[tid + i * kSeipThreads] : 0.0;
  };

  const double t1 = a.t1, rtol = a.rtol, atol = a.atol;
  const double inv_n = 1.0 / (double)n;
  auto save_time = [&](int k) -> double {
    if (k >= a.T) return CUDART_INF;
    if (a.save_dt > 0.0) return (k == a.T - 1) ? a.t1 : fma((double)k, a.save_dt, a.t0);
    return a.save_ts[k];
  };

  // ---- FSAL f0 and the initial step (Hairer-Wanner, PIDController._select_initial_step)
  eval(a.t0, f[0]);
  double tprev = a.t0, tnext;
  if (a.const_dt > 0.0) {
    tnext = a.t0 + a.const_dt;
  } else {
    double p0 = 0.0, p1 = 0.0;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      if (own[i]) {
        const double sc = atol + fabs(y[i]) * rtol;
        const double u = y[i] / sc, v = f[0][i] / sc;
        p0 += u * u;
        p1 += v * v;
      }
    }
    const double d0 = sqrt(block_sum(p0, sm.red) * inv_n);
    const double d1 = sqrt(block_sum(p1, sm.red) * inv_n);
    const bool small = (d0 < 1e-5) || (d1 < 1e-5);
    const double h0 = small ? 1e-6 : 0.01 * (d0 / d1);
#pragma unroll
    for (int i = 0; i < EPT; ++i) yst[i] = y[i] + h0 * f[0][i];
    eval(a.t0 + h0, f[1]);
    double p2 = 0.0;
#pragma unroll
    for (int i = 0; i < EPT; ++i) {
      if (own[i]) {
        const double sc = atol + fabs(y[i]) * rtol;
        const double u = (f[1][i] - f[0][i]) / sc;
        p2 += u * u;
      }
    }
    const double d2 = sqrt(block_sum(p2, sm.red) * inv_n) / h0;
    const double md = fmax(d1, d2);
    const double h1 = (md <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / md, 0.2);
    tnext = a.t0 + fmin(100.0 * h0, h1);
  }
  bool made_jump = false;  // the running step was clipped to end just before a discontinuity point
  if (EXT && a.n_jump > 0 && !(a.const_dt > 0.0)) tnext = seip_clip_to_jumps(a, a.t0, tnext, made_jump);
  tnext = fmin(tnext, t1);

  int32_t n_acc = 0, n_rej = 0, n_steps = 0, save_i = 0;
  const int ns = EXT ? a.n_saved : n;
  double* const out = a.ys + traj * (int64_t)a.T * ns;

  while (tprev < t1 && n_steps < a.max_steps) {
    const double h = tnext - tprev;
    // ---- Tsit5 stages 2..7 (stage index at compile time: the tableau row is static, the f_j are registers)
    auto stage = [&](auto sc) {
      constexpr int s = decltype(sc)::value;
      constexpr int base = s == 1 ? I_a21 : s == 2 ? I_a31 : s == 3 ? I_a41 : s == 4 ? I_a51 : s == 5 ? I_a61 : I_a71;
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        double acc = kTab[base] * f[0][i];
#pragma unroll
        for (int j = 1; j < s; ++j) acc = fma(kTab[base + j], f[j][i], acc);
        yst[i] = fma(h, acc, y[i]);
      }
      // stage times tprev + c_s h; the two c = 1 stages use tnext itself
      const double ts = s >= 5 ? tnext : fma(kTab[I_c2 + (s - 1)], h, tprev);
      eval(ts, f[s]);
    };
    stage(std::integral_constant<int, 1>{});
    stage(std::integral_constant<int, 2>{});
    stage(std::integral_constant<int, 3>{});
    stage(std::integral_constant<int, 4>{});
    stage(std::integral_constant<int, 5>{});
    stage(std::integral_constant<int, 6>{});
    // ---- embedded error, scaled RMS norm, I-controller
    bool keep;
    double dt_next;
    if (a.const_dt > 0.0) {
      keep = true;
      dt_next = a.const_dt;
    } else {
      double part = 0.0;
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        if (own[i]) {
          double er = kTab[I_e1] * f[0][i];
#pragma unroll
          for (int j = 1; j < 7; ++j) er = fma(kTab[I_e1 + j], f[j][i], er);
          er *= h;
          const double sc = fma(fmax(fabs(y[i]), fabs(yst[i])), rtol, atol);
          const double r = er / sc;
          part = fma(r, r, part);
        }
      }
      const double err2 = block_sum(part, sm.red) * inv_n;
      keep = err2 < 1.0;
      dt_next = h * controller_factor_sq(err2, keep);
    }
    double ntprev = keep ? tnext : tprev;
    bool next_made_jump = false;
    const bool jumps = EXT && a.n_jump > 0 && !(a.const_dt > 0.0);
    if (jumps && keep && made_jump) ntprev = nextafter(tnext, CUDART_INF);  // restart exactly at the jump
    double ntnext = ntprev + dt_next;
    if (jumps) ntnext = seip_clip_to_jumps(a, ntprev, ntnext, next_made_jump);
    ntprev = fmin(ntprev, t1);
    if (ntnext > t1 - 1e-10) ntnext = keep ? t1 : fma(0.5, t1 - ntprev, ntprev);
    ++n_steps;
    if (keep) {
      ++n_acc;
      // ---- SaveAt(ts): dense output for every ts[k] <= tnext, straight from the registers; consecutive threads
      // write consecutive elements
      const double inv_h = 1.0 / ((tnext == tprev) ? 1.0 : h);
      while (save_i < a.T && save_time(save_i) <= tnext) {
        const double th = (save_time(save_i) - tprev) * inv_h;
        double b[7];
#pragma unroll
        for (int j = 0; j < 7; ++j)
          b[j] = th * fma(th, fma(th, fma(th, kDense[j][3], kDense[j][2]), kDense[j][1]), kDense[j][0]);
        double* row = out + (int64_t)save_i * ns;
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
          const int so = save_off(i);
          if (so >= 0) {
            double acc = b[0] * f[0][i];
#pragma unroll
            for (int j = 1; j < 7; ++j) acc = fma(b[j], f[j][i], acc);
            row[so] = fma(h, acc, y[i]);
          }
        }
        ++save_i;
      }
#pragma unroll
      for (int i = 0; i < EPT; ++i) {
        y[i] = yst[i];
        f[0][i] = f[6][i];
      }
      if (jumps && made_jump) {  // no FSAL across a discontinuity: f0 is re-evaluated at (jump, y1)
        tprev = ntprev;
        eval(tprev, f[0]);
      }
    } else {
      ++n_rej;
    }
    if (jumps) made_jump = next_made_jump;
    tprev = ntprev;
    tnext = ntnext;
  }
  // slots never reached keep diffrax's +inf fill
  for (int k = save_i; k < a.T; ++k)
#pragma unroll
    for (int i = 0; i < EPT; ++i)
      if (save_off(i) >= 0) out[(int64_t)k * ns + save_off(i)] = CUDART_INF;
  if (tid == 0) {
    int32_t* st = a.stats + traj * 4;
    st[DYNODE_STAT_RESULT] = (tprev < t1) ? DYNODE_RESULT_MAX_STEPS : DYNODE_RESULT_OK;
    st[DYNODE_STAT_ACCEPTED] = n_acc;
    st[DYNODE_STAT_REJECTED] = n_rej;
    st[DYNODE_STAT_STEPS] = n_steps;
  }
}

size_t seip_smem_bytes(int A, int K, int W, int H, int V, int NK, int n) {
  const size_t doubles = (size_t)2 * n + 2 * A * K + 3 * K + W + (size_t)A * A + A + (size_t)H * V * W * K + 8 +
                         (size_t)A * V * (5 + 2 * NK) + (size_t)K * A + 3 * K;
  return doubles * sizeof(double);
}

}  // namespace

// This file is compiled four times (dynode_b200/_build.py): SEIP_EXT_UNIT = 0 holds the plain kernels and the C ABI;
// SEIP_EXT_UNIT = 1 with SEIP_EPT = 4 / 8 / 12 the kernels with the vaccination / introduction / seasonal-reset terms,
// one translation unit per elements-per-thread width -- the instantiations build in parallel (as one unit they took
// 190 s, three quarters of the whole library's build).
#ifndef SEIP_EXT_UNIT
#define SEIP_EXT_UNIT 0
#endif
cudaError_t seip_launch_plain(const void* args, size_t smem, int n, cudaStream_t stream);
cudaError_t seip_launch_ext(const void* args, size_t smem, int n, cudaStream_t stream);
cudaError_t seip_launch_ext_4(const void* args, size_t smem, cudaStream_t stream);
cudaError_t seip_launch_ext_8(const void* args, size_t smem, cudaStream_t stream);
cudaError_t seip_launch_ext_12(const void* args, size_t smem, cudaStream_t stream);

namespace {
cudaError_t seip_launch_kernel(void (*kern)(const SeipArgs), const SeipArgs& a, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaSuccess;
  if (smem > 48 * 1024) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<(unsigned)a.B, kSeipThreads, smem, stream>>>(a);
  return cudaGetLastError();
}
}  // namespace

// kernels specialised for the common (strains, waning stages) pairs (loops unrolled, index arithmetic folded); any
// other shape runs the generic one
#define SEIP_PICK_KW(EXT, EPT)                                          \
  kern = seip_solver_kernel<0, 0, EXT, EPT>;                            \
  if (a.K == 2 && a.W == 3) kern = seip_solver_kernel<2, 3, EXT, EPT>;  \
  if (a.K == 2 && a.W == 4) kern = seip_solver_kernel<2, 4, EXT, EPT>;  \
  if (a.K == 3 && a.W == 3) kern = seip_solver_kernel<3, 3, EXT, EPT>;  \
  if (a.K == 3 && a.W == 4) kern = seip_solver_kernel<3, 4, EXT, EPT>;

#if SEIP_EXT_UNIT
#define SEIP_CAT_(a, b) a##b
#define SEIP_CAT(a, b) SEIP_CAT_(a, b)
cudaError_t SEIP_CAT(seip_launch_ext_, SEIP_EPT)(const void* args, size_t smem, cudaStream_t stream) {
  const SeipArgs& a = *static_cast<const SeipArgs*>(args);
  void (*kern)(const SeipArgs) = nullptr;
  SEIP_PICK_KW(true, SEIP_EPT)
  return seip_launch_kernel(kern, a, smem, stream);
}
#else
// elements per thread, rounded up to a compiled width
cudaError_t seip_launch_ext(const void* args, size_t smem, int n, cudaStream_t stream) {
  const int ept = (n + kSeipThreads - 1) / kSeipThreads;
  return ept <= 4 ? seip_launch_ext_4(args, smem, stream)
                  : (ept <= 8 ? seip_launch_ext_8(args, smem, stream) : seip_launch_ext_12(args, smem, stream));
}
cudaError_t seip_launch_plain(const void* args, size_t smem, int n, cudaStream_t stream) {
  const SeipArgs& a = *static_cast<const SeipArgs*>(args);
  const int ept = (n + kSeipThreads - 1) / kSeipThreads;
  void (*kern)(const SeipArgs) = nullptr;
  if (ept <= 4) { SEIP_PICK_KW(false, 4) }
  else if (ept <= 8) { SEIP_PICK_KW(false, 8) }
  else { SEIP_PICK_KW(false, 12) }
  return seip_launch_kernel(kern, a, smem, stream);
}
#endif
#undef SEIP_PICK_KW

}  // namespace dynode

#if !SEIP_EXT_UNIT
using namespace dynode;

extern "C" {

int dynode_seip_state_size(const DynodeSeipDesc* m) {
  if (!m || m->n_ages < 1 || m->n_strains < 1 || m->n_strains > DYNODE_SEIP_MAX_STRAINS || m->n_wane < 1) return -1;
  if (m->n_vax < 0 || m->n_vax > 64 || m->n_knots < 0 || m->n_knots > 64) return -1;
  const int H = 1 << m->n_strains, V = m->n_vax > 0 ? m->n_vax : 1;
  return m->n_ages * H * V * (m->n_wane + 3 * m->n_strains);
}

int dynode_seip_solve_f64(const DynodeSeipDesc* model, const DynodeSolverDesc* sv, int64_t B, DynodeArray y0,
                          const DynodeSeipParams* p, const double* save_ts, int32_t T, double* ys, int32_t* stats,
                          void* stream) {
  if (!model || !sv || !p) return fail_msg("null model/solver/params descriptor");
  const int n = dynode_seip_state_size(model);
  if (n < 0)
    return fail_msg("unsupported ODE: SEIP dims ages=%d strains=%d wane=%d (1..%d strains); there is no CPU fallback",
                    model->n_ages, model->n_strains, model->n_wane, DYNODE_SEIP_MAX_STRAINS);
  if (n > DYNODE_SEIP_MAX_STATE || model->n_ages > 1023 || model->n_wane > 1023)
    return fail_msg("unsupported ODE: SEIP state of %d doubles exceeds the %d a thread block can stage in shared "
                    "memory; there is no CPU fallback", n, DYNODE_SEIP_MAX_STATE);
  if (B < 0) return fail_msg("negative ensemble size");
  if (!y0.ptr || !p->beta.ptr || !p->sigma.ptr || !p->gamma.ptr || !p->omega.ptr)
    return fail_msg("y0/beta/sigma/gamma/omega are required");
  if (!p->contact || !p->pop || !p->immunity) return fail_msg("contact/pop/immunity tables are required");
  if (!save_ts || T <= 0 || !ys || !stats) return fail_msg("save_ts (T >= 1), ys and stats are required");
  if (!(sv->t1 >= sv->t0)) return fail_msg("t1 must be >= t0");
  if (!(sv->const_dt > 0.0) && !(sv->rtol > 0.0 && sv->atol > 0.0)) return fail_msg("rtol/atol must be positive");
  if (sv->max_steps <= 0) return fail_msg("max_steps must be positive");
  if (sv->n_jump < 0 || sv->n_jump > 32) return fail_msg("n_jump must be in [0, 32]");
  if (sv->n_jump > 0 && !sv->jump_ts) return fail_msg("jump_ts is null");
  if (model->save_mask > 15u) return fail_msg("save_mask has bits beyond the four compartments (S, E, I, C)");
  if (B == 0) return 0;
  SeipArgs a;
  a.A = model->n_ages; a.K = model->n_strains; a.W = model->n_wane; a.H = 1 << a.K; a.n = n;
  a.V = model->n_vax > 0 ? model->n_vax : 1;
  a.NK = model->n_knots;
  a.B = B;
  a.y0 = y0; a.beta = p->beta; a.sigma = p->sigma; a.gamma = p->gamma; a.omega = p->omega;
  a.contact = p->contact; a.pop = p->pop; a.imm = p->immunity;
  a.vbase = p->vax_base; a.vknot = p->vax_knots; a.vcoef = p->vax_coef;
  if (a.vbase && a.NK > 0 && (!a.vknot || !a.vcoef)) return fail_msg("vax_knots / vax_coef are required with n_knots > 0");
  a.itime = p->intro_time; a.iscale = p->intro_scale; a.ipct = p->intro_pct; a.iages = p->intro_ages;
  if (a.ipct.ptr && (!a.itime.ptr || !a.iscale.ptr || !a.iages))
    return fail_msg("intro_time / intro_scale / intro_ages are required with intro_pct");
  a.season_tau = p->season_tau; a.season_on = p->season_on;
  a.t0 = sv->t0; a.t1 = sv->t1; a.rtol = sv->rtol; a.atol = sv->atol; a.const_dt = sv->const_dt;
  a.save_dt = sv->save_dt > 0.0 ? sv->save_dt : 0.0;
  a.save_ts = save_ts; a.T = T;
  a.max_steps = (int)(sv->max_steps > 0x7fffffff ? 0x7fffffff : sv->max_steps);
  a.ys = ys; a.stats = stats;
  // the reference ignores discontinuity points in constant-step mode (odes.py:113-131)
  a.jump_ts = (sv->n_jump > 0 && !(sv->const_dt > 0.0)) ? sv->jump_ts : nullptr;
  a.n_jump = a.jump_ts ? sv->n_jump : 0;
  a.save_mask = model->save_mask ? model->save_mask : 15u;
  {
    const int nS = a.A * a.H * a.V * a.W, nX = a.A * a.H * a.V * a.K;
    a.n_saved = ((a.save_mask & 1u) ? nS : 0) + nX * (((a.save_mask >> 1) & 1u) + ((a.save_mask >> 2) & 1u) +
                                                       ((a.save_mask >> 3) & 1u));
  }
  const size_t smem = seip_smem_bytes(a.A, a.K, a.W, a.H, a.V, a.NK, n);
  // kernels specialised for the common (strains, waning stages) pairs; any other shape runs the generic one
  // the general kernels also carry the solver options the plain ones leave out (discontinuity points, sub-save)
  const bool ext = a.V > 1 || a.vbase || a.ipct.ptr || a.season_on != 0.0 || a.n_jump > 0 || a.save_mask != 15u;
  const cudaError_t e = ext ? seip_launch_ext(&a, smem, n, (cudaStream_t)stream)
                            : seip_launch_plain(&a, smem, n, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : fail_msg("kernel launch failed: %s", cudaGetErrorString(e));
}

}  // extern "C"
#endif  // !SEIP_EXT_UNIT
