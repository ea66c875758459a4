// nuts_round.cu -- one CUDA thread per chain: the tree bookkeeping of a NUTS round
// (include/dynode_b200_nuts.h).  Algorithm: numpyro's iterative NUTS (hmc_util.py: _iterative_build_subtree,
// _is_iterative_turning, _combine_tree / _biased_transition_kernel, _double_tree; warmup_adapter with
// dual_averaging and welford_covariance), which is what the reference's MCMCProcess runs
// (reference src/dynode/infer/inference.py:149-163).  dynode_b200/infer/nuts.py holds the same round as
// masked torch tensor operations (CPU path and cross-check); this file replaces ~290 small kernels per
// round by two.
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/dynode_b200.h"
#include "../../include/dynode_b200_nuts.h"

namespace dynode {

int fail_msg(const char* fmt, ...);  // capi.cu: sets the thread-local message, returns 1

namespace {

constexpr int MD = DYNODE_NUTS_MAX_DIM;
constexpr double kMaxDeltaEnergy = 1000.0;

struct Vec {
  double v[MD];
};

// The kernels below are instantiated with the dimension at compile time for the common sizes (DT > 0): D is then a
// constant in these helpers, the loops unroll, and the Vec's live in registers instead of local memory (ncu: at 128
// chains the run-time-D post kernel took 70 us of a 280 us round, one thread walking ~10 D x D products out of local
// and global memory).  Same operations in the same order either way.
__device__ __forceinline__ void load(Vec& x, const double* p, int D) {
#pragma unroll
  for (int i = 0; i < D; ++i) x.v[i] = p[i];
}
__device__ __forceinline__ void store(double* p, const Vec& x, int D) {
#pragma unroll
  for (int i = 0; i < D; ++i) p[i] = x.v[i];
}
__device__ __forceinline__ void copy(double* dst, const double* src, int D) {
#pragma unroll
  for (int i = 0; i < D; ++i) dst[i] = src[i];
}
// y = M x, M row-major [D][D]
__device__ __forceinline__ void matvec(Vec& y, const double* M, const Vec& x, int D) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) acc += M[i * D + j] * x.v[j];
    y.v[i] = acc;
  }
}
__device__ __forceinline__ double dot(const Vec& a, const Vec& b, int D) {
  double acc = 0.0;
#pragma unroll
  for (int i = 0; i < D; ++i) acc += a.v[i] * b.v[i];
  return acc;
}
// the inverse mass matrix of a chain: a register copy when the dimension is a small compile-time constant
template <int DT>
struct MassCache {
  static constexpr bool CACHED = DT > 0 && DT <= 8;
  double m[CACHED ? DT * DT : 1];
  const double* p;
  __device__ __forceinline__ explicit MassCache(const double* g) : p(g) {
    if constexpr (CACHED) {
#pragma unroll
      for (int i = 0; i < DT * DT; ++i) m[i] = g[i];
    }
  }
  __device__ __forceinline__ const double* get() const { return CACHED ? m : p; }
};
__device__ __forceinline__ double logaddexp(double a, double b) {
  const double m = fmax(a, b);
  if (isinf(m)) return m;  // (-inf, -inf) -> -inf ; (+inf, .) -> +inf
  return m + log1p(exp(-fabs(a - b)));
}
// U-turn criterion (numpyro _is_turning): either end's velocity points against the centred momentum sum
__device__ __forceinline__ bool is_turning(const double* imm, const Vec& r_left, const Vec& r_right,
                                           const Vec& r_sum, int D) {
  Vec vl, vr, rc;
  matvec(vl, imm, r_left, D);
  matvec(vr, imm, r_right, D);
#pragma unroll
  for (int i = 0; i < D; ++i) rc.v[i] = r_sum.v[i] - 0.5 * (r_left.v[i] + r_right.v[i]);
  return (dot(vl, rc, D) <= 0.0) || (dot(vr, rc, D) <= 0.0);
}

template <int DT>
__global__ void __launch_bounds__(128) nuts_pre_kernel(const DynodeNutsState s, const double* __restrict__ rnd_n,
                                                       const double* __restrict__ rnd_u) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0) *s.any_active = 0;  // _post of this round sets it again if a chain is still running
  if (c >= s.C) return;
  const int D = DT > 0 ? DT : s.D;
  const int64_t o = (int64_t)c * D;
  const MassCache<DT> mass(s.imm + o * D);
  const double* imm = mass.get();
  const bool act = s.active[c] != 0;
  const bool probe = act && s.searching[c] != 0;
  if (probe) {
    // find_reasonable_step_size: fresh momentum, one leapfrog forward from the current state with eps * 2^dir
    const int64_t dir = s.fr_dir[c];
    const double e = s.eps[c] * (dir > 0 ? 2.0 : (dir < 0 ? 0.5 : 1.0));
    s.eps[c] = e;
    Vec xi, r0, v;
    load(xi, rnd_n + o, D);
    matvec(r0, s.msqrt + o * D, xi, D);
    matvec(v, imm, r0, D);
    s.energy0[c] = s.U[c] + 0.5 * dot(r0, v, D);
    copy(s.s_z + o, s.z + o, D);
    store(s.s_r + o, r0, D);
    copy(s.s_g + o, s.g + o, D);
    s.s_right[c] = 1;
  }
  if (act && !probe && s.need_tree[c]) {
    // fresh momentum r ~ N(0, M) and a one-node tree at the current state
    Vec xi, r0, v;
    load(xi, rnd_n + o, D);
    matvec(r0, s.msqrt + o * D, xi, D);
    matvec(v, imm, r0, D);
    s.energy0[c] = s.U[c] + 0.5 * dot(r0, v, D);
    copy(s.zL + o, s.z + o, D); copy(s.zR + o, s.z + o, D); copy(s.zP + o, s.z + o, D);
    copy(s.gL + o, s.g + o, D); copy(s.gR + o, s.g + o, D); copy(s.gP + o, s.g + o, D);
    store(s.rL + o, r0, D); store(s.rR + o, r0, D); store(s.r_sum + o, r0, D);
    s.UP[c] = s.U[c];
    s.weight[c] = 0.0; s.sum_acc[c] = 0.0;
    s.depth[c] = 0; s.nprop[c] = 0; s.s_n[c] = 0;
    s.turning[c] = 0; s.diverging[c] = 0;
    s.need_tree[c] = 0;
  }
  if (act && !probe && s.s_n[c] == 0) {
    // a new doubling: pick a direction and start from that edge of the tree
    const bool right = rnd_u[(int64_t)c * 3 + 0] < 0.5;
    s.s_right[c] = right;
    copy(s.s_z + o, (right ? s.zR : s.zL) + o, D);
    copy(s.s_r + o, (right ? s.rR : s.rL) + o, D);
    copy(s.s_g + o, (right ? s.gR : s.gL) + o, D);
  }
  // first half of the leapfrog (idle chains too: their result is discarded)
  const double h = s.s_right[c] ? s.eps[c] : -s.eps[c];
  Vec rh, v;
#pragma unroll
  for (int i = 0; i < D; ++i) rh.v[i] = s.s_r[o + i] - 0.5 * h * s.s_g[o + i];
  matvec(v, imm, rh, D);
#pragma unroll
  for (int i = 0; i < D; ++i) s.z_new[o + i] = s.s_z[o + i] + h * v.v[i];
  store(s.r_half + o, rh, D);
}

template <int DT>
__global__ void __launch_bounds__(128) nuts_post_kernel(const DynodeNutsState s, const double* __restrict__ U_in,
                                                        const double* __restrict__ g_in,
                                                        const double* __restrict__ rnd_u) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= s.C) return;
  const int D = DT > 0 ? DT : s.D, md = s.max_depth;
  const int64_t o = (int64_t)c * D;
  const bool act = s.active[c] != 0;
  if (!act) return;  // nothing of an idle chain changes in a round
  const MassCache<DT> mass(s.imm + o * D);
  const double* imm = mass.get();
  // "some chain is still running": set by every chain that entered this round active (one round late for the last
  // chain to finish, which costs the driver at most one extra host check)
  *s.any_active = 1;

  // ---- second half of the leapfrog; a non-finite potential or gradient makes a divergent leaf
  double U_new = U_in[c];
  Vec g_new, r_new, v;
  bool bad = !isfinite(U_new);
#pragma unroll
  for (int i = 0; i < D; ++i) { g_new.v[i] = g_in[o + i]; bad = bad || !isfinite(g_new.v[i]); }
  if (bad) { U_new = CUDART_INF; for (int i = 0; i < D; ++i) g_new.v[i] = 0.0; }
  const double h = s.s_right[c] ? s.eps[c] : -s.eps[c];
#pragma unroll
  for (int i = 0; i < D; ++i) r_new.v[i] = s.r_half[o + i] - 0.5 * h * g_new.v[i];
  matvec(v, imm, r_new, D);
  double delta = U_new + 0.5 * dot(r_new, v, D) - s.energy0[c];
  if (isnan(delta)) delta = CUDART_INF;
  if (s.searching[c]) {
    // ---- a step-size probe (numpyro find_reasonable_step_size: _body_fn / _cond_fn)
    const double de = U_new + 0.5 * dot(r_new, v, D) - s.energy0[c];  // NaN compares false: direction -1
    const int64_t dir_used = s.fr_dir[c];
    const int64_t dir_new = (log(0.8) < -de) ? 1 : -1;
    const double e = s.eps[c];
    const bool not_extreme = (e > 2.2250738585072014e-308 || dir_new >= 0) && (e < 1.7976931348623157e308 || dir_new <= 0);
    const bool go_on = not_extreme && (dir_used == 0 || dir_new == dir_used);
    s.n_leap[c] += 1;
    if (go_on) {
      s.fr_last[c] = dir_used;
      s.fr_dir[c] = dir_new;
    } else {  // the direction flipped: keep this step size, restart dual averaging around 10 x it
      s.searching[c] = 0;
      s.fr_dir[c] = 0; s.fr_last[c] = 0;
      s.da_prox[c] = log(10.0 * e);
      s.da_x[c] = 0.0; s.da_xavg[c] = 0.0; s.da_gavg[c] = 0.0; s.da_t[c] = 0.0;
      s.need_tree[c] = 1;
    }
    return;
  }
  const double leaf_w = -delta;
  const bool leaf_div = delta > kMaxDeltaEnergy;
  const double leaf_acc = fmin(exp(-delta), 1.0);

  // ---- fold the leaf into the subtree (uniform / multinomial transition inside a subtree)
  const int64_t n = s.s_n[c];
  const bool first = n == 0;
  const double s_w_old = s.s_w[c];
  const double new_w = first ? leaf_w : logaddexp(s_w_old, leaf_w);
  const double p_take = first ? 1.0 : 1.0 / (1.0 + exp(-(leaf_w - s_w_old)));
  if (rnd_u[(int64_t)c * 3 + 1] < p_take) {
    copy(s.s_zP + o, s.z_new + o, D);
    s.s_UP[c] = U_new;
    store(s.s_gP + o, g_new, D);
  }
  Vec rsum;
#pragma unroll
  for (int i = 0; i < D; ++i) rsum.v[i] = first ? r_new.v[i] : s.s_rsum[o + i] + r_new.v[i];
  store(s.s_rsum + o, rsum, D);
  copy(s.s_z + o, s.z_new + o, D);
  store(s.s_r + o, r_new, D);
  store(s.s_g + o, g_new, D);
  s.s_w[c] = new_w;
  const double s_acc = first ? leaf_acc : s.s_acc[c] + leaf_acc;
  s.s_acc[c] = s_acc;

  // ---- checkpointed U-turn tests over the sub-subtrees this leaf completes (_leaf_idx_to_ckpt_idxs)
  const int i_max = __popcll((unsigned long long)(n >> 1));
  int ones = 0;
  for (int64_t m = n; m & 1; m >>= 1) ++ones;
  const int i_min = i_max - ones + 1;
  double* rck = s.r_ck + (int64_t)c * md * D;
  double* rsck = s.rs_ck + (int64_t)c * md * D;
  if ((n & 1) == 0 && i_max < md) {
    store(rck + i_max * D, r_new, D);
    store(rsck + i_max * D, rsum, D);
  }
  bool s_turn = false;
  if (!first) {
    Vec vr;
    matvec(vr, imm, r_new, D);
    for (int i = i_min; i <= i_max && i < md; ++i) {
      if (i < 0) continue;
      Vec rl, vl, rc;
      load(rl, rck + i * D, D);
      matvec(vl, imm, rl, D);
#pragma unroll
      for (int j = 0; j < D; ++j)
        rc.v[j] = (rsum.v[j] - rsck[i * D + j] + rl.v[j]) - 0.5 * (rl.v[j] + r_new.v[j]);
      s_turn = s_turn || (dot(vl, rc, D) <= 0.0) || (dot(vr, rc, D) <= 0.0);
    }
  }
  const int64_t s_n = n + 1;

  // ---- subtree finished (full size, U-turn or divergence): double the tree (biased transition)
  int64_t depth = s.depth[c];
  const bool done_sub = (s_n >= ((int64_t)1 << depth)) || s_turn || leaf_div;
  s.s_turn[c] = s_turn;
  s.s_div[c] = leaf_div;
  s.s_n[c] = done_sub ? 0 : s_n;
  s.n_leap[c] += 1;
  if (!done_sub) return;

  const double weight = s.weight[c];
  double p_bias = fmin(exp(new_w - weight), 1.0);
  if (s_turn || leaf_div) p_bias = 0.0;
  if (rnd_u[(int64_t)c * 3 + 2] < p_bias) {
    copy(s.zP + o, s.s_zP + o, D);
    s.UP[c] = s.s_UP[c];
    copy(s.gP + o, s.s_gP + o, D);
  }
  const bool right = s.s_right[c] != 0;
  copy((right ? s.zR : s.zL) + o, s.z_new + o, D);
  store((right ? s.rR : s.rL) + o, r_new, D);
  store((right ? s.gR : s.gL) + o, g_new, D);
  s.weight[c] = logaddexp(weight, new_w);
  Vec tsum, rL, rR;
#pragma unroll
  for (int i = 0; i < D; ++i) tsum.v[i] = s.r_sum[o + i] + rsum.v[i];
  store(s.r_sum + o, tsum, D);
  load(rL, s.rL + o, D);
  load(rR, s.rR + o, D);
  const bool turning = s_turn || is_turning(imm, rL, rR, tsum, D);
  s.turning[c] = turning;
  s.diverging[c] = leaf_div;
  const double sum_acc = s.sum_acc[c] + s_acc;
  s.sum_acc[c] = sum_acc;
  const int64_t nprop = s.nprop[c] + s_n;
  s.nprop[c] = nprop;
  depth += 1;
  s.depth[c] = depth;

  // ---- tree complete: commit the transition and let the chain start its next tree in the next round
  if (!((depth >= md) || turning || leaf_div)) return;
  const double accept = sum_acc / (double)(nprop > 0 ? nprop : 1);
  Vec z;
  load(z, s.zP + o, D);
  store(s.z + o, z, D);
  const double U = s.UP[c];
  s.U[c] = U;
  copy(s.g + o, s.gP + o, D);
  s.last_accept[c] = accept;
  s.last_steps[c] = (double)nprop;
  const int64_t k = s.k[c];
  const int64_t nwin = *s.nwin;
  const unsigned fl = s.sched[k < nwin ? k : nwin - 1];
  if (fl & DYNODE_NUTS_ADAPT) {  // dual averaging of the log step size (t0 = 10, kappa = 0.75, gamma = 0.05)
    const double tt = s.da_t[c] + 1.0;
    const double gavg = (1.0 - 1.0 / (tt + 10.0)) * s.da_gavg[c] + (s.target_accept - accept) / (tt + 10.0);
    const double x = s.da_prox[c] - sqrt(tt) / 0.05 * gavg;
    const double wt = pow(tt, -0.75);
    s.da_xavg[c] = (1.0 - wt) * s.da_xavg[c] + wt * x;
    s.da_t[c] = tt;
    s.da_gavg[c] = gavg;
    s.da_x[c] = x;
    s.eps[c] = exp(fmin(fmax(x, -700.0), 700.0));
  }
  if (fl & DYNODE_NUTS_WELFORD) {  // Welford covariance of the positions
    const double n1 = s.wf_n[c] + 1.0;
    double* mean = s.wf_mean + o;
    double* m2 = s.wf_m2 + o * D;
    Vec d1;
#pragma unroll
    for (int i = 0; i < D; ++i) { d1.v[i] = z.v[i] - mean[i]; mean[i] += d1.v[i] / n1; }
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int j = 0; j < D; ++j) m2[i * D + j] += d1.v[i] * (z.v[j] - mean[j]);
    s.wf_n[c] = n1;
  }
  if (fl & DYNODE_NUTS_END_SLOW) {
    // End of a slow window (Stan / numpyro warmup_adapter): the chain's own window-end update, so that no
    // chain waits for another one.
    if (fl & DYNODE_NUTS_WELFORD) {
      const double n = s.sched_n[k < nwin ? k : nwin - 1];
      double* m2 = s.wf_m2 + o * D;
      if (n > 1.0) {
        // inverse mass matrix <- (n/(n+5)) cov + 1e-3 (5/(n+5)) I ; msqrt <- L^-T with imm = L L^T
        double* A = s.imm + o * D;
        double* Ms = s.msqrt + o * D;
        const double a = n / (n + 5.0), bshr = 1e-3 * (5.0 / (n + 5.0));
        for (int i = 0; i < D; ++i)
          for (int j = 0; j < D; ++j) {
            double v = a * (m2[i * D + j] / (n - 1.0));
            if (i == j) v += bshr;
            else if (!s.dense) v = 0.0;
            A[i * D + j] = v;
          }
        // Cholesky A = L L^T with L held in the Welford buffer (reset below), L^-1 by forward substitution
        // written transposed straight into msqrt: no per-thread scratch
        double* Lm = m2;
        for (int i = 0; i < D; ++i)
          for (int j = 0; j <= i; ++j) {
            double acc = A[i * D + j];
            for (int q = 0; q < j; ++q) acc -= Lm[i * D + q] * Lm[j * D + q];
            Lm[i * D + j] = (i == j) ? sqrt(acc) : acc / Lm[j * D + j];
          }
        for (int j = 0; j < D; ++j)  // column j of L^-1 = row j of msqrt
          for (int i = 0; i < D; ++i) {
            if (i < j) { Ms[j * D + i] = 0.0; continue; }
            double acc = (i == j) ? 1.0 : 0.0;
            for (int q = j; q < i; ++q) acc -= Lm[i * D + q] * Ms[j * D + q];
            Ms[j * D + i] = acc / Lm[i * D + i];
          }
      }
      s.wf_n[c] = 0.0;
      for (int i = 0; i < D; ++i) s.wf_mean[o + i] = 0.0;
      for (int i = 0; i < D * D; ++i) m2[i] = 0.0;
    }
    if (fl & DYNODE_NUTS_ADAPT) {  // step-size search from the current step size (new metric), see the probe branch
      s.searching[c] = 1;
      s.fr_dir[c] = 0; s.fr_last[c] = 0;
    }
  }
  if ((fl & DYNODE_NUTS_END_WARMUP) && (fl & DYNODE_NUTS_ADAPT))  // final step size = averaged iterate
    s.eps[c] = exp(fmin(fmax(s.da_xavg[c], -700.0), 700.0));
  if (fl & DYNODE_NUTS_SAMPLING) {
    int64_t kk = k - s.n_warmup;
    kk = kk < 0 ? 0 : (kk < s.N ? kk : s.N - 1);
    store(s.out_z + ((int64_t)c * s.N + kk) * D, z, D);
    const int64_t q = (int64_t)c * s.N + kk;
    s.out_accept[q] = accept;
    s.out_steps[q] = (double)nprop;
    s.out_div[q] = leaf_div ? 1.0 : 0.0;
    s.out_energy[q] = U;
    s.out_depth[q] = (double)depth;
  }
  s.k[c] = k + 1;
  s.need_tree[c] = 1;
  if (k + 1 >= nwin) s.active[c] = 0;
}

int check(const DynodeNutsState* st) {
  if (!st) return fail_msg("null NUTS state");
  if (st->C < 0 || st->D < 1 || st->D > DYNODE_NUTS_MAX_DIM)
    return fail_msg("NUTS dimension %d outside [1, %d]", st->D, DYNODE_NUTS_MAX_DIM);
  if (st->max_depth < 1 || st->max_depth > DYNODE_NUTS_MAX_DEPTH)
    return fail_msg("NUTS max_tree_depth %d outside [1, %d]", st->max_depth, DYNODE_NUTS_MAX_DEPTH);
  if (st->N < 1) return fail_msg("NUTS output capacity N must be >= 1");
  return 0;
}

}  // namespace
}  // namespace dynode

using namespace dynode;

extern "C" {

int dynode_nuts_round_pre(const DynodeNutsState* st, const double* rnd_n, const double* rnd_u, void* stream) {
  if (int rc = check(st)) return rc;
  if (!rnd_n || !rnd_u) return fail_msg("null random-number buffers");
  if (st->C == 0) return 0;
  const dim3 grid((st->C + 127) / 128);
  switch (st->D) {
#define DYN_NUTS_CASE(d) case d: nuts_pre_kernel<d><<<grid, 128, 0, (cudaStream_t)stream>>>(*st, rnd_n, rnd_u); break;
    DYN_NUTS_CASE(1) DYN_NUTS_CASE(2) DYN_NUTS_CASE(3) DYN_NUTS_CASE(4) DYN_NUTS_CASE(5) DYN_NUTS_CASE(6)
    DYN_NUTS_CASE(7) DYN_NUTS_CASE(8)
#undef DYN_NUTS_CASE
    default: nuts_pre_kernel<0><<<grid, 128, 0, (cudaStream_t)stream>>>(*st, rnd_n, rnd_u);
  }
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail_msg("nuts_pre launch failed: %s", cudaGetErrorString(e));
}

int dynode_nuts_round_post(const DynodeNutsState* st, const double* U_new, const double* g_new,
                           const double* rnd_u, void* stream) {
  if (int rc = check(st)) return rc;
  if (!U_new || !g_new || !rnd_u) return fail_msg("null potential / gradient / random-number buffers");
  if (st->C == 0) return 0;
  const dim3 grid((st->C + 127) / 128);
  switch (st->D) {
#define DYN_NUTS_CASE(d) \
  case d: nuts_post_kernel<d><<<grid, 128, 0, (cudaStream_t)stream>>>(*st, U_new, g_new, rnd_u); break;
    DYN_NUTS_CASE(1) DYN_NUTS_CASE(2) DYN_NUTS_CASE(3) DYN_NUTS_CASE(4) DYN_NUTS_CASE(5) DYN_NUTS_CASE(6)
    DYN_NUTS_CASE(7) DYN_NUTS_CASE(8)
#undef DYN_NUTS_CASE
    default: nuts_post_kernel<0><<<grid, 128, 0, (cudaStream_t)stream>>>(*st, U_new, g_new, rnd_u);
  }
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail_msg("nuts_post launch failed: %s", cudaGetErrorString(e));
}

}  // extern "C"
