// capi.cu -- the C ABI of include/dynode_b200.h: argument validation, dispatch to the compiled
// flow-family instances (instances.def), tangent-direction chunking.  No CPU fallback anywhere:
// an unsupported model returns an error.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/dynode_b200.h"
#include "solve_args.h"

namespace dynode {

static thread_local char g_err[512] = "";

static int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

// shared with nuts_round.cu
int fail_msg(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

typedef cudaError_t (*LaunchFn)(const SolveArgs&, cudaStream_t);
typedef cudaError_t (*AdjointFn)(const AdjointArgs&, cudaStream_t);

struct Instance {
  int flow, flags, g, s, chunk;
  LaunchFn save0, saveP, lik0, likP, saveJ, saveJP, likJ0, likJP;
  AdjointFn adjoint, adjointJ;
  LaunchFn lik1;  // fused log-likelihood with one direction per work item (latency regime of chunk-2 flows)
};

#define X(IDX, FLOW, FLAGS, G, S)                                                        \
  {FLOW, FLAGS, G, S, tangent_chunk(FLOW),                                               \
   &launch_lane_solver<FLOW, FLAGS, G, S, 0, MODE_SAVE>,                                 \
   &launch_lane_solver<FLOW, FLAGS, G, S, tangent_chunk(FLOW), MODE_SAVE>,               \
   &launch_lane_solver<FLOW, FLAGS, G, S, 0, MODE_LOGLIK>,                               \
   &launch_lane_solver<FLOW, FLAGS, G, S, tangent_chunk(FLOW), MODE_LOGLIK>,             \
   &launch_lane_solver<FLOW, FLAGS, G, S, 0, MODE_SAVE_JUMPS>,                           \
   &launch_lane_solver<FLOW, FLAGS, G, S, tangent_chunk(FLOW), MODE_SAVE_JUMPS>,         \
   &launch_lane_solver<FLOW, FLAGS, G, S, 0, MODE_LOGLIK_JUMPS>,                         \
   &launch_lane_solver<FLOW, FLAGS, G, S, tangent_chunk(FLOW), MODE_LOGLIK_JUMPS>,       \
   &launch_adjoint_solver<FLOW, FLAGS, G, S, false>,                                     \
   &launch_adjoint_solver<FLOW, FLAGS, G, S, true>,                                      \
   &launch_loglik_single_direction<FLOW, FLAGS, G, S>},
static const Instance kInstances[] = {
#include "instances.def"
};
#undef X

static const Instance* find_instance(const DynodeModelDesc* m) {
  if (!m) return nullptr;
  for (const Instance& i : kInstances)
    if (i.flow == m->flow && i.flags == m->flags && i.g == m->n_groups && i.s == m->n_strains) return &i;
  return nullptr;
}

static int ncomp(int flow) {
  switch (flow) {
    case DYNODE_FLOW_SIR: return 3;
    case DYNODE_FLOW_SEIRS: return 4;
    case DYNODE_FLOW_SEIRS_C: return 5;
  }
  return -1;
}

static int check_common(const DynodeModelDesc* model, const DynodeSolverDesc* sv, int64_t B,
                        DynodeArray y0, const DynodeParams* p, const double* save_ts, int32_t T,
                        const Instance** inst) {
  if (!model || !sv || !p) return fail("null model/solver/params descriptor");
  *inst = find_instance(model);
  if (!*inst)
    return fail("unsupported ODE: flow=%d flags=%d groups=%d strains=%d is not in the compiled flow "
                "family (dynode_b200/csrc/instances.def); there is no CPU fallback",
                model->flow, model->flags, model->n_groups, model->n_strains);
  if (B < 0) return fail("negative ensemble size");
  if (!y0.ptr) return fail("y0 is null");
  if (!p->beta.ptr || !p->gamma.ptr) return fail("beta/gamma are required");
  if (model->flow != DYNODE_FLOW_SIR && !p->sigma.ptr) return fail("sigma is required for flows with an exposed compartment");
  if ((model->flags & DYNODE_FLAG_SEASONAL) &&
      (!p->season_amp.ptr || !p->season_phase.ptr || !p->season_period.ptr))
    return fail("seasonal flow needs season_amp/phase/period");
  if (!save_ts || T <= 0) return fail("save_ts is required (T >= 1)");
  if (!(sv->t1 >= sv->t0)) return fail("t1 must be >= t0");
  if (!(sv->const_dt > 0.0) && !(sv->rtol > 0.0 && sv->atol > 0.0)) return fail("rtol/atol must be positive");
  if (sv->max_steps <= 0) return fail("max_steps must be positive");
  if (sv->n_jump < 0 || sv->n_jump > kMaxJumps) return fail("n_jump must be in [0, %d]", kMaxJumps);
  if (sv->n_jump > 0 && !sv->jump_ts) return fail("jump_ts is null");
  if (sv->n_jump > 0 && sv->const_dt > 0.0)
    return fail("discontinuity points apply to the adaptive controller only (odes.py:115-131)");
  return 0;
}

static void fill_common(SolveArgs& a, const DynodeSolverDesc* sv, int64_t B, DynodeArray y0,
                        const DynodeParams* p, const double* save_ts, int32_t T) {
  memset(&a, 0, sizeof(a));
  a.B = B;
  a.y0 = y0;
  a.prm = *p;
  a.save_ts = save_ts;
  a.T = T;
  a.t0 = sv->t0; a.t1 = sv->t1; a.rtol = sv->rtol; a.atol = sv->atol; a.const_dt = sv->const_dt;
  a.save_dt = sv->save_dt > 0.0 ? sv->save_dt : 0.0;
  a.jump_ts = sv->n_jump > 0 ? sv->jump_ts : nullptr;
  a.n_jump = sv->n_jump > 0 ? sv->n_jump : 0;
  a.max_steps = (int32_t)(sv->max_steps > 0x7fffffff ? 0x7fffffff : sv->max_steps);
  a.write_primal = 1;
  a.n_pass = 1;
  a.only = sv->only;
  for (int k = 0; k < kMaxWrt; ++k) a.wrt[k] = -1;
}

static int check_wrt(const DynodeModelDesc* m, int32_t n_wrt, const int32_t* wrt) {
  if (n_wrt < 0 || n_wrt > 64) return fail("n_wrt out of range");
  if (n_wrt > 0 && !wrt) return fail("wrt is null");
  for (int k = 0; k < n_wrt; ++k) {
    if (wrt[k] < 0) continue;
    const int kind = wrt[k] >> 4, strain = wrt[k] & 15;
    if (kind > DYNODE_P_SEASON_PHASE || strain >= m->n_strains) return fail("bad wrt id %d", wrt[k]);
  }
  return 0;
}

static int run_passes(const Instance* inst, SolveArgs& a, int32_t n_wrt, const int32_t* wrt, bool loglik,
                      cudaStream_t stream) {
  if (a.B == 0) return 0;
  cudaError_t e;
  const bool jumps = a.n_jump > 0;
  if (n_wrt == 0) {
    a.P_total = 0;
    a.n_pass = 1;
    e = (loglik ? (jumps ? inst->likJ0 : inst->lik0) : (jumps ? inst->saveJ : inst->save0))(a, stream);
    if (e != cudaSuccess) return fail("kernel launch failed: %s", cudaGetErrorString(e));
    return 0;
  }
  // tangent directions ride the same step sequence in groups of `chunk` directions (registers); all groups
  // of all trajectories are work items of ONE launch (solve_args.h), so a many-parameter gradient for a few
  // hundred NUTS chains still fills the GPU instead of serialising ceil(P/chunk) small launches
  a.P_total = n_wrt;
  a.n_pass = (n_wrt + inst->chunk - 1) / inst->chunk;
  for (int k = 0; k < kMaxWrt; ++k) a.wrt[k] = k < n_wrt ? wrt[k] : -1;
  a.write_primal = 1;
  if (loglik && !jumps && inst->chunk > 1 && n_wrt > 1) {
    // Few chains: the launch lasts as long as one warp's instruction stream (measured: a lone warp of the config-2
    // kernel issues 0.28 instructions per cycle, 98 us for 34 steps).  One direction per work item shortens that
    // stream by a third; the repeated primal is free while the work items do not fill the GPU's resident warps.
    static const int64_t few = [] { const char* v = getenv("DYNODE_B200_SINGLE_DIRECTION_BELOW"); return v ? atoll(v) : 1184 * 16; }();
    if (a.B * n_wrt <= few) {
      SolveArgs a1 = a;
      a1.n_pass = n_wrt;
      e = inst->lik1(a1, stream);
      if (e == cudaSuccess) return 0;
      if (e != cudaErrorNotSupported) return fail("kernel launch failed: %s", cudaGetErrorString(e));
    }
  }
  e = (loglik ? (jumps ? inst->likJP : inst->likP) : (jumps ? inst->saveJP : inst->saveP))(a, stream);
  if (e != cudaSuccess) return fail("kernel launch failed: %s", cudaGetErrorString(e));
  return 0;
}

// ---- roofline probes ------------------------------------------------------------------------
constexpr int kProbeIlp = 8;
__global__ void __launch_bounds__(256) probe_dfma_kernel(double* sink, int iters) {
  double x[kProbeIlp];
#pragma unroll
  for (int k = 0; k < kProbeIlp; ++k) x[k] = 1.0 + 1e-9 * (threadIdx.x + k);
  const double a = 1.0000001, b = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < kProbeIlp; ++k) x[k] = fma(x[k], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < kProbeIlp; ++k) s += x[k];
  if (s == 123.456) sink[0] = s;  // never true; keeps the loop alive
}
__global__ void __launch_bounds__(256) probe_write_kernel(double2* dst, int64_t n2) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride)
    dst[i] = make_double2((double)i, 1.0);
}

}  // namespace dynode

using namespace dynode;

extern "C" {

int dynode_version(void) { return DYNODE_B200_VERSION; }
const char* dynode_last_error(void) { return g_err; }

int dynode_num_compartments(const DynodeModelDesc* m) { return m ? ncomp(m->flow) : -1; }

int dynode_state_size(const DynodeModelDesc* m) {
  if (!m || ncomp(m->flow) < 0) return -1;
  return m->n_groups + (ncomp(m->flow) - 1) * m->n_groups * m->n_strains;
}

int dynode_saved_size(const DynodeModelDesc* m, uint32_t mask) {
  if (!m || ncomp(m->flow) < 0) return -1;
  int n = 0;
  for (int c = 0; c < ncomp(m->flow); ++c)
    if ((mask >> c) & 1u) n += (c == 0) ? m->n_groups : m->n_groups * m->n_strains;
  return n;
}

int dynode_is_supported(const DynodeModelDesc* m) {
  if (find_instance(m)) return 1;
  if (m) fail("flow=%d flags=%d groups=%d strains=%d is not compiled in", m->flow, m->flags, m->n_groups, m->n_strains);
  return 0;
}

int dynode_solve_f64(const DynodeModelDesc* model, const DynodeSolverDesc* solver, int64_t B,
                     DynodeArray y0, const DynodeParams* params, const double* save_ts, int32_t T,
                     uint32_t save_comp_mask, double* ys, int32_t* stats, void* stream) {
  return dynode_solve_sens_f64(model, solver, B, y0, params, save_ts, T, save_comp_mask, 0, nullptr,
                               nullptr, ys, nullptr, stats, stream);
}

int dynode_solve_sens_f64(const DynodeModelDesc* model, const DynodeSolverDesc* solver, int64_t B,
                          DynodeArray y0, const DynodeParams* params, const double* save_ts, int32_t T,
                          uint32_t save_comp_mask, int32_t n_wrt, const int32_t* wrt, const double* dy0,
                          double* ys, double* dys, int32_t* stats, void* stream) {
  const Instance* inst = nullptr;
  if (int rc = check_common(model, solver, B, y0, params, save_ts, T, &inst)) return rc;
  if (int rc = check_wrt(model, n_wrt, wrt)) return rc;
  if (!stats) return fail("stats is null");
  const int ns = dynode_saved_size(model, save_comp_mask);
  if (ns > 0 && !ys) return fail("ys is null");
  if (n_wrt > 0 && ns > 0 && !dys) return fail("dys is null");
  SolveArgs a;
  fill_common(a, solver, B, y0, params, save_ts, T);
  a.save_mask = save_comp_mask;
  a.ys = ys;
  a.stats = stats;
  a.dy0 = dy0;
  a.dys = dys;
  return run_passes(inst, a, n_wrt, wrt, /*loglik=*/false, (cudaStream_t)stream);
}

int dynode_poisson_loglik_grad_f64(const DynodeModelDesc* model, const DynodeSolverDesc* solver, int64_t B,
                                   DynodeArray y0, const DynodeParams* params, const double* save_ts,
                                   int32_t T, int32_t obs_comp, const double* obs, double lp_const,
                                   int32_t n_wrt, const int32_t* wrt, const double* dy0, double* lp,
                                   double* grad, int32_t* stats, void* stream) {
  const Instance* inst = nullptr;
  if (int rc = check_common(model, solver, B, y0, params, save_ts, T, &inst)) return rc;
  if (int rc = check_wrt(model, n_wrt, wrt)) return rc;
  if (obs_comp < 0 || obs_comp >= ncomp(model->flow)) return fail("obs_comp out of range");
  if (!obs || !lp || !stats) return fail("obs/lp/stats must not be null");
  if (n_wrt > 0 && !grad) return fail("grad is null");
  if (T < 2) return fail("need at least two save times to form increments");
  SolveArgs a;
  fill_common(a, solver, B, y0, params, save_ts, T);
  a.stats = stats;
  a.dy0 = dy0;
  a.obs_comp = obs_comp;
  a.obs = obs;
  a.lp_const = lp_const;
  a.lp = lp;
  a.grad = grad;
  return run_passes(inst, a, n_wrt, wrt, /*loglik=*/true, (cudaStream_t)stream);
}

int dynode_poisson_loglik_adjoint_f64(const DynodeModelDesc* model, const DynodeSolverDesc* solver, int64_t B,
                                      DynodeArray y0, const DynodeParams* params, const double* save_ts,
                                      int32_t T, int32_t obs_comp, const double* obs, double lp_const,
                                      double* lp, double* grad, double* grad_y0, int32_t* stats, double* ckpt,
                                      int32_t cap, double* vsave, void* stream) {
  const Instance* inst = nullptr;
  if (int rc = check_common(model, solver, B, y0, params, save_ts, T, &inst)) return rc;
  if (obs_comp < 0 || obs_comp >= ncomp(model->flow)) return fail("obs_comp out of range");
  if (!obs || !lp || !grad || !stats) return fail("obs/lp/grad/stats must not be null");
  if (!ckpt || !vsave || cap < 1) return fail("the adjoint needs checkpoint scratch (ckpt, vsave, cap >= 1)");
  if (T < 2) return fail("need at least two save times to form increments");
  AdjointArgs a;
  fill_common(a.s, solver, B, y0, params, save_ts, T);
  a.s.stats = stats;
  a.s.obs_comp = obs_comp;
  a.s.obs = obs;
  a.s.lp_const = lp_const;
  a.s.lp = lp;
  a.grad = grad;
  a.grad_y0 = grad_y0;
  a.ckpt = ckpt;
  a.vsave = vsave;
  a.cap = cap;
  if (B == 0) return 0;
  const cudaError_t e = (a.s.n_jump > 0 ? inst->adjointJ : inst->adjoint)(a, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail("kernel launch failed: %s", cudaGetErrorString(e));
  return 0;
}

int64_t dynode_probe_dfma(double* sink, int32_t iters, void* stream) {
  const int grid = 148 * 8, block = 256;
  probe_dfma_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(sink, iters);
  if (cudaGetLastError() != cudaSuccess) return -1;
  return (int64_t)grid * block * kProbeIlp * 2 * (int64_t)iters;
}

int dynode_probe_hbm_write(double* dst, int64_t n, void* stream) {
  if (dst == nullptr || n < 0 || (reinterpret_cast<uintptr_t>(dst) & 15u) != 0)
    return fail("probe: dst must be a 16-byte aligned device pointer");
  probe_write_kernel<<<148 * 16, 256, 0, (cudaStream_t)stream>>>((double2*)dst, n / 2);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // extern "C"
