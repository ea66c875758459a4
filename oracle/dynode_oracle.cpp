// dynode_oracle.cpp -- CPU restatement of DynODE's hot path.  TEST INFRASTRUCTURE ONLY.
//
// PARITY UNPINNED: the arithmetic of the reference's hot path lives in third-party
// diffrax 0.7.* on jax>=0.6.1,<0.7 (reference pyproject.toml:12-14), which is NOT under
// /root/reference and cannot be installed in this image (no network).  The solver loop,
// Tsit5 tableau, dense output, PID controller and initial-step rule below are restated from
// the published diffrax algorithm; they are verified only
// indirectly (order conditions, interpolant identities, the reference's physics pins, a
// DOP853 truth check).  The right-hand sides ARE pinned: tests/golden/rhs_golden.npz holds
// outputs of the reference's own RHS functions (examples/*.py) executed in the build
// container by tests/golden/make_rhs_golden.py.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product path (dynode_b200/) never links or calls it.
//
// Reference call sites this file follows:
//   src/dynode/simulation/odes.py:107-144   ODETerm, t0=0, dt0=None, PIDController(rtol,atol),
//                                           ConstantStepSize branch, diffeqsolve, max_steps
//   src/dynode/simulation/odes.py:148-198   SaveAt(ts=linspace(...)) (host side, see oracle.py)
//   src/dynode/config/params.py:24-67       Tsit5, rtol=1e-5, atol=1e-6, max_steps=1e6
//   examples/sir.py:78-84, seirs.py:88-95, seirs_seasonal_forcing.py:34-55,
//   sir_age_stratified.py:127-142, sir_age_risk_stratified.py:157-173,
//   seirs_multi_strain_age_stratified.py:213-243, tests/test_simulation/test_odes.py:17-28  (RHS)
//   examples/sir_infer_parameters.py:21-39  Poisson likelihood on diff(R)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ----------------------------------------------------------------------------------------
// Forward-mode dual number: value + P tangents.  P = 0 degenerates to a plain double.
// Gradients of the reference are those of the *discrete* scheme with the step sequence
// frozen (diffrax stop_gradient on the controller factor and on the automatic dt0), which
// is exactly what carrying tangents through the accepted steps computes.
// ----------------------------------------------------------------------------------------
template <int P>
struct Dual {
  double v;
  double d[P > 0 ? P : 1];
  Dual() : v(0.0) { for (int p = 0; p < P; ++p) d[p] = 0.0; }
  Dual(double x) : v(x) { for (int p = 0; p < P; ++p) d[p] = 0.0; }
};
template <int P> inline Dual<P> operator+(const Dual<P>& a, const Dual<P>& b) {
  Dual<P> r; r.v = a.v + b.v; for (int p = 0; p < P; ++p) r.d[p] = a.d[p] + b.d[p]; return r; }
template <int P> inline Dual<P> operator-(const Dual<P>& a, const Dual<P>& b) {
  Dual<P> r; r.v = a.v - b.v; for (int p = 0; p < P; ++p) r.d[p] = a.d[p] - b.d[p]; return r; }
template <int P> inline Dual<P> operator-(const Dual<P>& a) {
  Dual<P> r; r.v = -a.v; for (int p = 0; p < P; ++p) r.d[p] = -a.d[p]; return r; }
template <int P> inline Dual<P> operator*(const Dual<P>& a, const Dual<P>& b) {
  Dual<P> r; r.v = a.v * b.v; for (int p = 0; p < P; ++p) r.d[p] = a.d[p] * b.v + a.v * b.d[p]; return r; }
template <int P> inline Dual<P> operator/(const Dual<P>& a, const Dual<P>& b) {
  Dual<P> r; r.v = a.v / b.v;
  for (int p = 0; p < P; ++p) r.d[p] = (a.d[p] - r.v * b.d[p]) / b.v; return r; }
template <int P> inline Dual<P> operator*(double a, const Dual<P>& b) {
  Dual<P> r; r.v = a * b.v; for (int p = 0; p < P; ++p) r.d[p] = a * b.d[p]; return r; }
template <int P> inline Dual<P> operator*(const Dual<P>& b, double a) { return a * b; }
template <int P> inline Dual<P> operator+(double a, const Dual<P>& b) {
  Dual<P> r = b; r.v = a + b.v; return r; }
template <int P> inline Dual<P> operator+(const Dual<P>& b, double a) { return a + b; }
template <int P> inline Dual<P> operator-(const Dual<P>& b, double a) {
  Dual<P> r = b; r.v = b.v - a; return r; }
template <int P> inline Dual<P> dsin(const Dual<P>& a) {
  Dual<P> r; r.v = std::sin(a.v); double c = std::cos(a.v);
  for (int p = 0; p < P; ++p) r.d[p] = c * a.d[p]; return r; }
template <int P> inline Dual<P> dexp(const Dual<P>& a) {
  Dual<P> r; r.v = std::exp(a.v); for (int p = 0; p < P; ++p) r.d[p] = r.v * a.d[p]; return r; }
template <int P> inline Dual<P> dlog(const Dual<P>& a) {
  Dual<P> r; r.v = std::log(a.v); for (int p = 0; p < P; ++p) r.d[p] = a.d[p] / a.v; return r; }

// ----------------------------------------------------------------------------------------
// Tsit5 tableau (diffrax _solver/tsit5.py).  Digits beyond double
// precision are kept so the literals round exactly as the published ones do.
// ----------------------------------------------------------------------------------------
const double C_[6] = {161.0 / 1000.0, 327.0 / 1000.0, 9.0 / 10.0,
                      0.9800255409045096857298102862870245954942137979563024768854764293221195950761080302604,
                      1.0, 1.0};
const double A2_[1] = {161.0 / 1000.0};
const double A3_[2] = {
    -0.8480655492356988544426874250230774675121177393430391537369234245294192976164141156943e-2,
    0.3354806554923569885444268742502307746751211773934303915373692342452941929761641411569};
const double A4_[3] = {
    2.897153057105493432130432594192938764924887287701866490314866693455023795137503079289,
    -6.359448489975074843148159912383825625952700647415626703305928850207288721235210244366,
    4.362295432869581411017727318190886861027813359713760212991062156752264926097707165077};
const double A5_[4] = {
    5.325864828439256604428877920840511317836476253097040101202360397727981648835607691791,
    -11.74888356406282787774717033978577296188744178259862899288666928009020615663593781589,
    7.495539342889836208304604784564358155658679161518186721010132816213648793440552049753,
    -0.9249506636175524925650207933207191611349983406029535244034750452930469056411389539635e-1};
const double A6_[5] = {
    5.861455442946420028659251486982647890394337666164814434818157239052507339770711679748,
    -12.92096931784710929170611868178335939541780751955743459166312250439928519268343184452,
    8.159367898576158643180400794539253485181918321135053305748355423955009222648673734986,
    -0.7158497328140099722453054252582973869127213147363544882721139659546372402303777878835e-1,
    -0.2826905039406838290900305721271224146717633626879770007617876201276764571291579142206e-1};
const double A7_[6] = {
    0.9646076681806522951816731316512876333711995238157997181903319145764851595234062815396e-1,
    1.0 / 100.0,
    0.4798896504144995747752495322905965199130404621990332488332634944254542060153074523509,
    1.379008574103741893192274821856872770756462643091360525934940067397245698027561293331,
    -3.290069515436080679901047585711363850115683290894936158531296799594813811049925401677,
    2.324710524099773982415355918398765796109060233222962411944060046314465391054716027841};
const double* const A_[6] = {A2_, A3_, A4_, A5_, A6_, A7_};
// b_error = b_sol - b_hat, formed in double exactly as the published tableau does.
const double BERR_[7] = {
    0.9646076681806522951816731316512876333711995238157997181903319145764851595234062815396e-1 -
        0.9468075576583945807478876255758922856117527357724631226139574065785592789071067303271e-1,
    1.0 / 100.0 -
        0.9183565540343253096776363936645313759813746240984095238905939532922955247253608687270e-2,
    0.4798896504144995747752495322905965199130404621990332488332634944254542060153074523509 -
        0.4877705284247615707855642599631228241516691959761363774365216240304071651579571959813,
    1.379008574103741893192274821856872770756462643091360525934940067397245698027561293331 -
        1.234297566930478985655109673884237654035539930748192848315425833500484878378061439761,
    -3.290069515436080679901047585711363850115683290894936158531296799594813811049925401677 -
        -2.707712349983525454881109975059321670689605166938197378763992255714444407154902012702,
    2.324710524099773982415355918398765796109060233222962411944060046314465391054716027841 -
        1.866628418170587035753719399566211498666255505244122593996591602841258328965767580089,
    0.0 - 1.0 / 66.0};

// Dense-output weights (diffrax _Tsit5Interpolation.evaluate).
inline void dense_weights(double t, double b[7]) {
  b[0] = -1.0530884977290216 * t * (t - 1.3299890189751412) *
         (t * t - 1.4364028541716351 * t + 0.7139816917074209);
  b[1] = 0.1017 * (t * t) * (t * t - 2.1966568338249754 * t + 1.2949852507374631);
  b[2] = 2.490627285651252793 * (t * t) * (t * t - 2.38535645472061657 * t + 1.57803468208092486);
  b[3] = -16.54810288924490272 * (t - 1.21712927295533244) * (t - 0.61620406037800089) * (t * t);
  b[4] = 47.37952196281928122 * (t - 1.203071208372362603) * (t - 0.658047292653547382) * (t * t);
  b[5] = -34.87065786149660974 * (t - 1.2) * (t - 0.666666666666666667) * (t * t);
  b[6] = 2.5 * (t - 1.0) * (t - 0.6) * (t * t);
}

// ----------------------------------------------------------------------------------------
// Right-hand sides.  Each function restates one reference callable line by line, in the
// reference's own operation order.  State is the compartments concatenated, each compartment
// flattened in C order.  theta is the per-trajectory parameter vector, shared the
// un-batched tensor (contact matrix).
// ----------------------------------------------------------------------------------------
enum Family {
  FAM_SIR_1BIN = 0,          // examples/sir.py:78-84            theta = [beta, gamma]
  FAM_SIR_DENSITY = 1,       // tests/test_simulation/test_odes.py:17-28   theta = [beta, gamma]
  FAM_SEIRS_1BIN = 2,        // examples/seirs.py:88-95          theta = [beta, gamma, sigma, omega]
  FAM_SEIRS_SEASONAL = 3,    // examples/seirs_seasonal_forcing.py:34-55
                             //   theta = [beta, gamma, sigma, omega, amp, phase, period]
  FAM_SIR_AGE = 4,           // examples/sir_age_stratified.py:127-142   theta=[beta,gamma], shared=C[A][A]
  FAM_SIR_AGE_RISK = 5,      // examples/sir_age_risk_stratified.py:157-173 shared=CM[A][R][A][R]
  FAM_SEIRS_MULTISTRAIN = 6, // examples/seirs_multi_strain_age_stratified.py:213-243
                             //   theta = [beta[S], gamma[S], sigma[S], omega[S]], shared = C[A][A]
  FAM_SEIPV = 8,             // FAM_SEIP + the vaccination dimension, spline vaccination rates, the seasonal tier
                             //   reset and external introductions of reference ode_model.md:15-53,72-75,179-190
                             //   (utils/splines.py:72-109, config/strains.py:59-109); see the case below
  FAM_SEIP = 7               // immune-history / waning family after reference ode_model.md:15-53,100-118,179-211
                             //   (no reference implementation exists; the equations below are this repo's
                             //   reading of the prose model, without the vaccination dimension):
                             //   dims A ages, R = W waning stages, S = K strains, H = 2^K immune histories
                             //   state  S[A][H][W], E[A][H][K], I[A][H][K], C[A][H][K]
                             //   theta  [beta[K], sigma[K], gamma[K], omega[W]]  (omega[W-1] unused: last stage absorbs)
                             //   shared [contact[A][A] (target, source), pop[A], immunity[H][W][K] in [0, 1]]
};

struct Dims { int A, R, S; int V = 1, NK = 0; };
// FAM_SEIPV carries two more sizes than the three ints of the C interface: they ride in the upper bytes of S
// (S = K | V << 8 | NK << 16; oracle.py::seipv_dims builds it).
inline Dims make_dims(int fam, int A, int R, int S) {
  Dims d{A, R, S};
  if (fam == 8) { d.S = S & 0xff; d.V = (S >> 8) & 0xff; d.NK = (S >> 16) & 0xff; }
  return d;
}
const int MAXG = 32;  // max groups (A*R)
const int MAXS = 8;   // max strains

inline int state_size(int fam, Dims d) {
  switch (fam) {
    case FAM_SIR_1BIN: case FAM_SIR_DENSITY: return 3;
    case FAM_SEIRS_1BIN: case FAM_SEIRS_SEASONAL: return 4;
    case FAM_SIR_AGE: return 3 * d.A;
    case FAM_SIR_AGE_RISK: return 3 * d.A * d.R;
    case FAM_SEIRS_MULTISTRAIN: return d.A + 4 * d.A * d.S;
    case FAM_SEIP: return d.A * (1 << d.S) * (d.R + 3 * d.S);
    case FAM_SEIPV: return d.A * (1 << d.S) * d.V * (d.R + 3 * d.S);
  }
  return -1;
}
inline int theta_size(int fam, Dims d) {
  switch (fam) {
    case FAM_SIR_1BIN: case FAM_SIR_DENSITY: case FAM_SIR_AGE: case FAM_SIR_AGE_RISK: return 2;
    case FAM_SEIRS_1BIN: return 4;
    case FAM_SEIRS_SEASONAL: return 7;
    case FAM_SEIRS_MULTISTRAIN: return 4 * d.S;
    case FAM_SEIP: return 3 * d.S + d.R;
    case FAM_SEIPV: return 6 * d.S + d.R;
  }
  return -1;
}

template <class T>
void rhs(int fam, Dims dm, double t, const T* y, const T* th, const double* sh, T* dy) {
  switch (fam) {
    case FAM_SIR_1BIN: {  // sir.py:78-84
      T s = y[0], i = y[1], r = y[2];
      T N = s + i + r;
      T beta = th[0], gamma = th[1];
      dy[0] = -beta * s * i / N;
      dy[1] = beta * s * i / N - gamma * i;
      dy[2] = gamma * i;
    } break;
    case FAM_SIR_DENSITY: {  // test_odes.py:17-28
      T s = y[0], i = y[1];
      T s_to_i = th[0] * s * i;
      T i_to_r = i * th[1];
      dy[0] = -s_to_i;
      dy[1] = s_to_i - i_to_r;
      dy[2] = i_to_r;
    } break;
    case FAM_SEIRS_1BIN:
    case FAM_SEIRS_SEASONAL: {  // seirs.py:88-95 ; seirs_seasonal_forcing.py:34-55
      T s = y[0], e = y[1], i = y[2], r = y[3];
      T N = s + e + i + r;
      T beta = th[0], gamma = th[1], sigma = th[2], omega = th[3];
      if (fam == FAM_SEIRS_SEASONAL) {
        // 1.0 + amp * sin(2 * pi * t / period + phase), evaluated left to right
        T two_pi_t = T((2.0 * M_PI) * t);
        T seas = 1.0 + th[4] * dsin(two_pi_t / th[6] + th[5]);
        beta = beta * seas;
      }
      dy[0] = -beta * s * i / N + omega * r;
      dy[1] = beta * s * i / N - sigma * e;
      dy[2] = sigma * e - gamma * i;
      dy[3] = gamma * i - omega * r;
    } break;
    case FAM_SIR_AGE: {  // sir_age_stratified.py:127-142
      const int A = dm.A;
      const T* s = y; const T* i = y + A; const T* r = y + 2 * A;
      T beta = th[0], gamma = th[1];
      T pop[MAXG];
      for (int a = 0; a < A; ++a) pop[a] = s[a] + i[a] + r[a];
      for (int a = 0; a < A; ++a) {
        // beta * sum_b (C[a,b] * i[b]) / pop[b]
        T acc = (sh[a * A + 0] * i[0]) / pop[0];
        for (int b = 1; b < A; ++b) acc = acc + (sh[a * A + b] * i[b]) / pop[b];
        T foi = beta * acc;
        T s_to_i = s[a] * foi;
        T i_to_r = i[a] * gamma;
        dy[a] = -s_to_i;
        dy[A + a] = s_to_i - i_to_r;
        dy[2 * A + a] = i_to_r;
      }
    } break;
    case FAM_SIR_AGE_RISK: {  // sir_age_risk_stratified.py:157-173
      const int A = dm.A, R = dm.R, G = A * R;
      const T* s = y; const T* i = y + G; const T* r = y + 2 * G;
      T beta = th[0], gamma = th[1];
      T prop[MAXG];
      for (int g = 0; g < G; ++g) prop[g] = i[g] / (s[g] + i[g] + r[g]);
      for (int k = 0; k < G; ++k) {  // k = (k_age, l_risk) target
        // einsum("ijkl,ij->kl", CM, i/pop): sum over source (i,j) of CM[i,j,k,l]*prop[i,j]
        T acc = sh[0 * G + k] * prop[0];
        for (int g = 1; g < G; ++g) acc = acc + sh[g * G + k] * prop[g];
        T foi = beta * acc;
        T s_to_i = s[k] * foi;
        T i_to_r = i[k] * gamma;
        dy[k] = -s_to_i;
        dy[G + k] = s_to_i - i_to_r;
        dy[2 * G + k] = i_to_r;
      }
    } break;
    case FAM_SEIRS_MULTISTRAIN: {  // seirs_multi_strain_age_stratified.py:213-243
      const int A = dm.A, S = dm.S, AS = A * S;
      const T* s = y; const T* e = y + A; const T* i = e + AS; const T* r = i + AS;
      const T* beta = th; const T* gamma = th + S; const T* sigma = th + 2 * S; const T* omega = th + 3 * S;
      T N[MAXG], fois[MAXG * MAXS];
      for (int a = 0; a < A; ++a) {
        T se = e[a * S], si = i[a * S], sr = r[a * S];
        for (int k = 1; k < S; ++k) { se = se + e[a * S + k]; si = si + i[a * S + k]; sr = sr + r[a * S + k]; }
        N[a] = s[a] + se + si + sr;
      }
      for (int k = 0; k < S; ++k)
        for (int a = 0; a < A; ++a) {
          T acc = sh[a * A + 0] * (i[0 * S + k] / N[0]);
          for (int b = 1; b < A; ++b) acc = acc + sh[a * A + b] * (i[b * S + k] / N[b]);
          fois[a * S + k] = beta[k] * acc;
        }
      T* ds = dy; T* de = dy + A; T* di = de + AS; T* dr = di + AS; T* dc = dr + AS;
      for (int a = 0; a < A; ++a) {
        T inf = fois[a * S] * s[a];
        T wan = omega[0] * r[a * S];
        for (int k = 1; k < S; ++k) { inf = inf + fois[a * S + k] * s[a]; wan = wan + omega[k] * r[a * S + k]; }
        ds[a] = -inf + wan;
        for (int k = 0; k < S; ++k) {
          const int q = a * S + k;
          de[q] = fois[q] * s[a] - sigma[k] * e[q];
          di[q] = sigma[k] * e[q] - gamma[k] * i[q];
          dr[q] = gamma[k] * i[q] - omega[k] * r[q];
          dc[q] = fois[q] * s[a];
        }
      }
    } break;
    case FAM_SEIP: {
      // force of infection per (age, strain) through the contact matrix; a susceptible cell (age a, immune
      // history j, waning stage w) is exposed to strain k at rate foi[a][k] * (1 - immunity[j][w][k]);
      // exposure keeps the history j while infected; recovery moves to history eta(j, k) = j | 2^k
      // (ode_model.md:100-118), waning stage 0; waning is a chain w -> w+1 at rate omega[w].
      const int A = dm.A, W = dm.R, K = dm.S, H = 1 << K;
      const T* Sx = y; const T* E = y + A * H * W; const T* I = E + A * H * K;
      const T* beta = th; const T* sigma = th + K; const T* gamma = th + 2 * K; const T* omega = th + 3 * K;
      const double* contact = sh; const double* pop = sh + A * A; const double* imm = pop + A;
      T itot[MAXG * MAXS], foi[MAXG * MAXS];
      for (int a = 0; a < A; ++a)
        for (int k = 0; k < K; ++k) {
          T acc = I[(a * H + 0) * K + k];
          for (int j = 1; j < H; ++j) acc = acc + I[(a * H + j) * K + k];
          itot[a * K + k] = acc;
        }
      for (int a = 0; a < A; ++a)
        for (int k = 0; k < K; ++k) {
          T acc = contact[a * A + 0] * (itot[0 * K + k] / T(pop[0]));
          for (int b = 1; b < A; ++b) acc = acc + contact[a * A + b] * (itot[b * K + k] / T(pop[b]));
          foi[a * K + k] = beta[k] * acc;
        }
      T* dS = dy; T* dE = dy + A * H * W; T* dI = dE + A * H * K; T* dC = dI + A * H * K;
      for (int a = 0; a < A; ++a)
        for (int j = 0; j < H; ++j) {
          T expo[MAXS];
          for (int k = 0; k < K; ++k) expo[k] = T(0.0);
          for (int w = 0; w < W; ++w) {
            const T s = Sx[(a * H + j) * W + w];
            T out = T(0.0);
            for (int k = 0; k < K; ++k) {
              T x = foi[a * K + k] * (1.0 - imm[(j * W + w) * K + k]) * s;
              expo[k] = expo[k] + x;
              out = out + x;
            }
            T d = -out;
            if (w > 0) d = d + omega[w - 1] * Sx[(a * H + j) * W + w - 1];
            if (w < W - 1) d = d - omega[w] * s;
            if (w == 0)
              for (int k = 0; k < K; ++k)
                if ((j >> k) & 1)
                  d = d + gamma[k] * (I[(a * H + j) * K + k] + I[(a * H + (j ^ (1 << k))) * K + k]);
            dS[(a * H + j) * W + w] = d;
          }
          for (int k = 0; k < K; ++k) {
            const int q = (a * H + j) * K + k;
            dE[q] = expo[k] - sigma[k] * E[q];
            dI[q] = sigma[k] * E[q] - gamma[k] * I[q];
            dC[q] = expo[k];
          }
        }
    } break;
    case FAM_SEIPV: {
      // FAM_SEIP with a vaccination tier v = 0..V-1 on every cell (ode_model.md:15-53; this repository's reading
      // where the prose is silent or does not conserve people):
      //   state  S[A][H][V][W], E[A][H][V][K], I[A][H][V][K], C[A][H][V][K]
      //   theta  [beta[K], sigma[K], gamma[K], omega[W], intro_time[K], intro_scale[K], intro_pct[K]]
      //   shared [contact[A][A], pop[A], immunity[H][V][W][K], vax_base[A][V][4], vax_knots[A][V][NK],
      //           vax_coef[A][V][NK], intro_ages[K][A], season_tau, season_on]
      //   external introductions (ode_model.md:183, strains.py:59-109): the infectious fraction of age b for strain k
      //     is (sum_{j,v} I[b][j][v][k] + N(t; intro_time_k, intro_scale_k) intro_pct_k intro_ages[k][b] pop[b]) / pop[b]
      //   vaccination (ode_model.md:19-29, splines.py:72-109): nu[a][v](t) = max(0, cubic spline), the rate out of
      //     tier v is r[a][v] = min(nu pop[a] / sum_{j,w} S[a][j][v][w], 1); S[a][j][v][w] -> S[a][j][v+1][0];
      //     in the top tier a dose moves waning stages w >= 1 back to stage 0 (the prose adds sum_w including w = 0 to
      //     stage 0 without removing it, which creates people; here the w = 0 term is left out of both sides)
      //   seasonal reset (ode_model.md:32,37,43,72-75): phi(t) = season_on sin(2 pi (t + tau) / 730)^1000 moves the top
      //     tier of S, E and I to the tier below
      const int A = dm.A, W = dm.R, K = dm.S, H = 1 << K, V = dm.V, NK = dm.NK;
      const int nS = A * H * V * W, nX = A * H * V * K;
      const T* Sx = y; const T* E = y + nS; const T* I = E + nX;
      const T* beta = th; const T* sigma = th + K; const T* gamma = th + 2 * K; const T* omega = th + 3 * K;
      const T* itime = omega + W; const T* iscale = itime + K; const T* ipct = iscale + K;
      const double* contact = sh; const double* pop = contact + A * A; const double* imm = pop + A;
      const double* vbase = imm + H * V * W * K; const double* vknot = vbase + A * V * 4;
      const double* vcoef = vknot + A * V * NK; const double* iages = vcoef + A * V * NK;
      const double tau = iages[K * A], season_on = iages[K * A + 1];
      std::vector<T> frac(A * K), foi(A * K), rate(A * V);
      for (int a = 0; a < A; ++a)
        for (int k = 0; k < K; ++k) {
          T acc = T(0.0);
          for (int j = 0; j < H; ++j)
            for (int v = 0; v < V; ++v) acc = acc + I[((a * H + j) * V + v) * K + k];
          T f = acc / T(pop[a]);
          if (val(ipct[k]) != 0.0) {
            const T zs = (T(t) - itime[k]) / iscale[k];
            const T pdf = dexp(T(-0.5) * zs * zs) / (iscale[k] * 2.5066282746310002);
            f = f + pdf * ipct[k] * iages[k * A + a];
          }
          frac[a * K + k] = f;
        }
      for (int a = 0; a < A; ++a)
        for (int k = 0; k < K; ++k) {
          T acc = contact[a * A + 0] * frac[0 * K + k];
          for (int b = 1; b < A; ++b) acc = acc + contact[a * A + b] * frac[b * K + k];
          foi[a * K + k] = beta[k] * acc;
        }
      for (int a = 0; a < A; ++a)
        for (int v = 0; v < V; ++v) {
          const double* bs = vbase + (a * V + v) * 4;
          double nu = bs[0] + bs[1] * t + bs[2] * t * t + bs[3] * t * t * t;
          for (int i = 0; i < NK; ++i) {
            const double d = t - vknot[(a * V + v) * NK + i];
            if (d > 0.0) nu += vcoef[(a * V + v) * NK + i] * d * d * d;
          }
          if (!(nu > 0.0)) nu = 0.0;
          T tot = T(0.0);
          for (int j = 0; j < H; ++j)
            for (int w = 0; w < W; ++w) tot = tot + Sx[((a * H + j) * V + v) * W + w];
          T r = T(0.0);
          if (nu > 0.0 && val(tot) > 0.0) {
            r = T(nu * pop[a]) / tot;
            if (!(val(r) < 1.0)) r = T(1.0);
          }
          rate[a * V + v] = r;
        }
      double phi = 0.0;
      if (season_on != 0.0 && V >= 2) {
        const double sn = std::sin(2.0 * 3.14159265358979323846 * (t + tau) / 730.0);
        phi = season_on * std::pow(sn * sn, 500.0);
      }
      T* dS = dy; T* dE = dy + nS; T* dI = dE + nX; T* dC = dI + nX;
      for (int a = 0; a < A; ++a)
        for (int j = 0; j < H; ++j)
          for (int v = 0; v < V; ++v) {
            const int cell = (a * H + j) * V + v;
            const bool top = v == V - 1;
            T expo[MAXS];
            for (int k = 0; k < K; ++k) expo[k] = T(0.0);
            // people vaccinated into this cell's stage 0: from the tier below, and boosters within the top tier
            T vin = T(0.0);
            if (v >= 1) {
              T below = T(0.0);
              for (int w = 0; w < W; ++w) below = below + Sx[(cell - 1) * W + w];
              vin = vin + rate[a * V + v - 1] * below;
            }
            if (top) {
              T older = T(0.0);
              for (int w = 1; w < W; ++w) older = older + Sx[cell * W + w];
              vin = vin + rate[a * V + v] * older;
            }
            for (int w = 0; w < W; ++w) {
              const T s = Sx[cell * W + w];
              T out = T(0.0);
              for (int k = 0; k < K; ++k) {
                T x = foi[a * K + k] * (1.0 - imm[((j * V + v) * W + w) * K + k]) * s;
                expo[k] = expo[k] + x;
                out = out + x;
              }
              T d = -out;
              if (w > 0) d = d + omega[w - 1] * Sx[cell * W + w - 1];
              if (w < W - 1) d = d - omega[w] * s;
              if (w == 0)
                for (int k = 0; k < K; ++k)
                  if ((j >> k) & 1)
                    d = d + gamma[k] * (I[cell * K + k] + I[((a * H + (j ^ (1 << k))) * V + v) * K + k]);
              if (!(top && w == 0)) d = d - rate[a * V + v] * s;
              if (w == 0) d = d + vin;
              if (phi != 0.0) {
                if (top) d = d - phi * s;
                if (v == V - 2) d = d + phi * Sx[(cell + 1) * W + w];
              }
              dS[cell * W + w] = d;
            }
            for (int k = 0; k < K; ++k) {
              const int q = cell * K + k;
              T de = expo[k] - sigma[k] * E[q];
              T di = sigma[k] * E[q] - gamma[k] * I[q];
              if (phi != 0.0) {
                if (top) { de = de - phi * E[q]; di = di - phi * I[q]; }
                if (v == V - 2) { de = de + phi * E[q + K]; di = di + phi * I[q + K]; }
              }
              dE[q] = de;
              dI[q] = di;
              dC[q] = expo[k];
            }
          }
    } break;
  }
}

template <class T> inline double val(const T& x) { return x.v; }

// rms_norm over the whole flattened state (diffrax rms_norm: sqrt(mean(x*x))).
inline double rms(const double* x, int n) {
  double acc = 0.0;
  for (int e = 0; e < n; ++e) acc += x[e] * x[e];
  return std::sqrt(acc / n);
}

struct SolveCfg {
  double t0, t1, rtol, atol, const_dt;
  int64_t max_steps;
  const double* save_ts; int T;
  const int32_t* save_idx; int n_saved;
  const double* jump_ts; int n_jump;  // SolverParams.discontinuity_points (odes.py:120-131), sorted
};

// ClipStepSizeController(jump_ts): with i0 = #{jump <= t0} and i1 = #{jump <= t1}
// (searchsorted side="right"), a jump lies in (t0, t1] iff i0 < i1; the step then ends at
// prevbefore(jump[i0]) and the NEXT accepted step starts at nextafter(that) == the jump itself, with the
// FSAL derivative re-evaluated there.
inline double clip_to_jumps(const SolveCfg& c, double t0, double t1, bool& made_jump) {
  int i0 = 0, i1 = 0;
  for (int k = 0; k < c.n_jump; ++k) { i0 += (c.jump_ts[k] <= t0); i1 += (c.jump_ts[k] <= t1); }
  made_jump = i0 < i1;
  if (!made_jump) return t1;
  return std::nextafter(c.jump_ts[std::min(i0, c.n_jump - 1)], -std::numeric_limits<double>::infinity());
}

// Hairer-Wanner initial step (diffrax PIDController._select_initial_step).
template <class T>
double select_initial_step(int fam, Dims dm, int n, const SolveCfg& c, const T* y0, const T* f0,
                           const T* th, const double* sh) {
  std::vector<double> scale(n), tmp(n);
  for (int e = 0; e < n; ++e) scale[e] = c.atol + std::fabs(val(y0[e])) * c.rtol;
  for (int e = 0; e < n; ++e) tmp[e] = val(y0[e]) / scale[e];
  double d0 = rms(tmp.data(), n);
  for (int e = 0; e < n; ++e) tmp[e] = val(f0[e]) / scale[e];
  double d1 = rms(tmp.data(), n);
  bool cond = (d0 < 1e-5) || (d1 < 1e-5);
  double d1s = cond ? 1.0 : d1;
  double h0 = cond ? 1e-6 : 0.01 * (d0 / d1s);
  double t1 = c.t0 + h0;
  std::vector<T> y1(n), f1(n);
  for (int e = 0; e < n; ++e) y1[e] = y0[e] + h0 * f0[e];
  rhs<T>(fam, dm, t1, y1.data(), th, sh, f1.data());
  for (int e = 0; e < n; ++e) tmp[e] = (val(f1[e]) - val(f0[e])) / scale[e];
  double d2 = rms(tmp.data(), n) / h0;
  double max_d = std::max(d1, d2);
  double h1 = (max_d <= 1e-15) ? std::max(1e-6, h0 * 1e-3) : std::pow(0.01 / max_d, 1.0 / 5.0);
  return std::min(100.0 * h0, h1);
}

// One trajectory.  out_y: [T][n_saved] values, out_dy: [T][n_saved][P] tangents (may be null),
// stats: {result, num_accepted, num_rejected, num_steps}.  If lp_out != null the Poisson
// log-likelihood on the daily increments of the saved elements is accumulated instead
// (sir_infer_parameters.py:30-38): obs is [T-1][n_saved].
template <int P>
void solve_one(int fam, Dims dm, const SolveCfg& c, const Dual<P>* y0, const Dual<P>* th,
               const double* sh, double* out_y, double* out_dy, int32_t* stats) {
  typedef Dual<P> T;
  const int n = state_size(fam, dm);
  std::vector<T> y(y0, y0 + n), y1(n), ys(n), yerr(n);
  std::vector<std::vector<T>> f(7, std::vector<T>(n));
  std::vector<double> tmp(n);
  const double inf = std::numeric_limits<double>::infinity();
  // output buffer pre-filled with inf (diffrax SaveState init)
  for (int64_t q = 0; q < (int64_t)c.T * c.n_saved; ++q) out_y[q] = inf;
  if (out_dy) for (int64_t q = 0; q < (int64_t)c.T * c.n_saved * P; ++q) out_dy[q] = inf;

  double tprev = c.t0, tnext;
  rhs<T>(fam, dm, c.t0, y.data(), th, sh, f[0].data());  // FSAL f0 (solver.init)
  if (c.const_dt > 0.0) {
    tnext = c.t0 + c.const_dt;
  } else {
    double dt0 = select_initial_step<T>(fam, dm, n, c, y.data(), f[0].data(), th, sh);
    tnext = c.t0 + dt0;
  }
  bool made_jump = false;
  if (c.n_jump > 0 && !(c.const_dt > 0.0)) tnext = clip_to_jumps(c, c.t0, tnext, made_jump);
  tnext = std::min(tnext, c.t1);
  int64_t num_steps = 0; int32_t n_acc = 0, n_rej = 0; int save_i = 0;

  while (tprev < c.t1 && num_steps < c.max_steps) {
    const double h = tnext - tprev;
    // --- Tsit5.step: 6 new stages; stage 7 is the solution (SSAL) and the next f0 (FSAL)
    for (int st = 1; st <= 6; ++st) {
      const double* a = A_[st - 1];
      for (int e = 0; e < n; ++e) {
        T acc = a[0] * f[0][e];
        for (int j = 1; j < st; ++j) acc = acc + a[j] * f[j][e];
        ys[e] = y[e] + h * acc;
      }
      const double ti = (C_[st - 1] == 1.0) ? tnext : tprev + C_[st - 1] * h;
      rhs<T>(fam, dm, ti, ys.data(), th, sh, f[st].data());
      if (st == 6) y1 = ys;
    }
    for (int e = 0; e < n; ++e) {
      T acc = BERR_[0] * f[0][e];
      for (int j = 1; j < 7; ++j) acc = acc + BERR_[j] * f[j][e];
      yerr[e] = h * acc;
    }
    // --- PIDController.adapt_step_size (I-controller) / ConstantStepSize
    bool keep; double dt_next;
    if (c.const_dt > 0.0) {
      keep = true; dt_next = c.const_dt;
    } else {
      bool any_nan = false;
      for (int e = 0; e < n; ++e) any_nan = any_nan || std::isnan(val(y1[e]));
      for (int e = 0; e < n; ++e) {
        double ye = val(yerr[e]); if (std::isnan(ye)) ye = inf;
        double y1c = any_nan ? val(y[e]) : val(y1[e]);
        double sc = std::max(std::fabs(val(y[e])), std::fabs(y1c));
        tmp[e] = ye / (c.atol + sc * c.rtol);
      }
      double err = rms(tmp.data(), n);
      keep = err < 1.0;
      double inv = 1.0 / err;
      double factor = 0.9 * std::pow(inv, 1.0 / 5.0);
      double fmin = keep ? 1.0 : 0.2;
      factor = std::min(std::max(factor, fmin), 10.0);
      dt_next = h * factor;
    }
    double ntprev = keep ? tnext : tprev;
    bool next_made_jump = false;
    const bool jumps = c.n_jump > 0 && !(c.const_dt > 0.0);
    if (jumps && keep && made_jump) ntprev = std::nextafter(tnext, inf);
    double ntnext = ntprev + dt_next;
    if (jumps) ntnext = clip_to_jumps(c, ntprev, ntnext, next_made_jump);
    ntprev = std::min(ntprev, c.t1);
    // _clip_to_end
    if (ntnext > c.t1 - 1e-10) ntnext = keep ? c.t1 : ntprev + 0.5 * (c.t1 - ntprev);
    num_steps += 1;
    if (keep) {
      n_acc += 1;
      // --- save by dense output over [tprev, tnext] (interpolator built from y, k=h*f)
      while (save_i < c.T && c.save_ts[save_i] <= tnext) {
        const double ts = c.save_ts[save_i];
        const double div = (tnext == tprev) ? 1.0 : (tnext - tprev);
        const double theta = (ts - tprev) / div;
        double b[7]; dense_weights(theta, b);
        for (int q = 0; q < c.n_saved; ++q) {
          const int e = c.save_idx[q];
          T acc = b[0] * (h * f[0][e]);
          for (int j = 1; j < 7; ++j) acc = acc + b[j] * (h * f[j][e]);
          T v = y[e] + acc;
          out_y[(int64_t)save_i * c.n_saved + q] = v.v;
          if (out_dy) for (int p = 0; p < P; ++p)
            out_dy[((int64_t)save_i * c.n_saved + q) * P + p] = v.d[p];
        }
        ++save_i;
      }
      y = y1; f[0] = f[6];
      if (jumps && made_jump) rhs<T>(fam, dm, ntprev, y.data(), th, sh, f[0].data());  // no FSAL across a jump
    } else {
      n_rej += 1;
    }
    if (jumps) made_jump = next_made_jump;
    tprev = ntprev; tnext = ntnext;
  }
  stats[0] = (tprev < c.t1) ? 1 : 0;  // 1 = max_steps reached
  stats[1] = n_acc; stats[2] = n_rej; stats[3] = (int32_t)std::min<int64_t>(num_steps, INT32_MAX);
}

template <int P>
int solve_batch(int fam, Dims dm, const SolveCfg& c, int64_t B, const double* y0, int64_t y0_bs,
                const double* theta, int64_t th_bs, const double* sh, const int32_t* wrt,
                const double* dy0, double* ys, double* dys, int32_t* stats, int nthreads) {
  const int n = state_size(fam, dm), nt = theta_size(fam, dm);
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 16)
#endif
  for (int64_t b = 0; b < B; ++b) {
    std::vector<Dual<P>> yy(n), th(nt);
    for (int e = 0; e < n; ++e) yy[e] = Dual<P>(y0[b * y0_bs + e]);
    for (int k = 0; k < nt; ++k) th[k] = Dual<P>(theta[b * th_bs + k]);
    for (int p = 0; p < P; ++p) {
      if (wrt[p] >= 0) th[wrt[p]].d[p] = 1.0;
      if (dy0) for (int e = 0; e < n; ++e) yy[e].d[p] = dy0[((int64_t)b * P + p) * n + e];
    }
    solve_one<P>(fam, dm, c, yy.data(), th.data(), sh, ys + (int64_t)b * c.T * c.n_saved,
                 dys ? dys + (int64_t)b * c.T * c.n_saved * P : nullptr, stats + 4 * b);
  }
  return 0;
}

}  // namespace

extern "C" {

int oracle_state_size(int fam, int A, int R, int S) { return state_size(fam, make_dims(fam, A, R, S)); }
int oracle_theta_size(int fam, int A, int R, int S) { return theta_size(fam, make_dims(fam, A, R, S)); }

// dy = f(t, y; theta, shared)
int oracle_rhs(int fam, int A, int R, int S, double t, const double* y, const double* theta,
               const double* shared, double* dy) {
  Dims dm = make_dims(fam, A, R, S);
  const int n = state_size(fam, dm), nt = theta_size(fam, dm);
  if (n < 0) return 1;
  std::vector<Dual<0>> yy(n), th(nt), out(n);
  for (int e = 0; e < n; ++e) yy[e] = Dual<0>(y[e]);
  for (int k = 0; k < nt; ++k) th[k] = Dual<0>(theta[k]);
  rhs<Dual<0>>(fam, dm, t, yy.data(), th.data(), shared, out.data());
  for (int e = 0; e < n; ++e) dy[e] = out[e].v;
  return 0;
}

void oracle_dense_weights(double theta, double* b7) { dense_weights(theta, b7); }
void oracle_tableau(double* c6, double* a21, double* berr7) {
  for (int i = 0; i < 6; ++i) c6[i] = C_[i];
  int q = 0;
  for (int i = 0; i < 6; ++i) for (int j = 0; j <= i; ++j) a21[q++] = A_[i][j];
  for (int i = 0; i < 7; ++i) berr7[i] = BERR_[i];
}

// Batched solve.  y0: [B][n] with batch stride y0_bs (0 = shared), theta likewise.
// wrt[n_wrt]: theta indices seeded with unit tangents (-1 = only dy0 seeds this direction);
// dy0: optional [B][n_wrt][n] initial-state tangents.  ys: [B][T][n_saved];
// dys: [B][T][n_saved][n_wrt]; stats: [B][4] = {result, accepted, rejected, steps}.
int oracle_solve(int fam, int A, int R, int S, int64_t B, const double* y0, int64_t y0_bs,
                 const double* theta, int64_t th_bs, const double* shared, double t0, double t1,
                 double rtol, double atol, int64_t max_steps, double const_dt,
                 const double* save_ts, int T, const int32_t* save_idx, int n_saved, int n_wrt,
                 const int32_t* wrt, const double* dy0, double* ys, double* dys, int32_t* stats,
                 int nthreads, const double* jump_ts, int n_jump) {
  Dims dm = make_dims(fam, A, R, S);
  if (state_size(fam, dm) < 0 || (fam != FAM_SEIP && fam != FAM_SEIPV && A * R > MAXG) || A > MAXG || dm.S > MAXS) return 1;
  SolveCfg c{t0, t1, rtol, atol, const_dt, max_steps, save_ts, T, save_idx, n_saved, jump_ts, n_jump};
#define ORC_CASE(PP) case PP: return solve_batch<PP>(fam, dm, c, B, y0, y0_bs, theta, th_bs, shared, \
                                                    wrt, dy0, ys, dys, stats, nthreads);
  switch (n_wrt) {
    ORC_CASE(0) ORC_CASE(1) ORC_CASE(2) ORC_CASE(3) ORC_CASE(4) ORC_CASE(5) ORC_CASE(6)
    ORC_CASE(7) ORC_CASE(8) ORC_CASE(9) ORC_CASE(10) ORC_CASE(11) ORC_CASE(12)
  }
#undef ORC_CASE
  return 2;  // unsupported tangent count
}

// Poisson log-likelihood on daily increments (sir_infer_parameters.py:30-38) from saved
// values and tangents:  rate = max(diff(ys, axis=time), 1e-6);
// lp = sum(obs*log(rate) - rate - lgamma(obs+1)).   ys: [B][T][m], dys: [B][T][m][P],
// obs: [T-1][m].  lp: [B], grad: [B][P].
int oracle_poisson_incidence(int64_t B, int T, int m, int P, const double* ys, const double* dys,
                             const double* obs, double* lp, double* grad) {
  for (int64_t b = 0; b < B; ++b) {
    double acc = 0.0; std::vector<double> g(P, 0.0);
    for (int k = 0; k + 1 < T; ++k)
      for (int q = 0; q < m; ++q) {
        const int64_t i1 = ((int64_t)b * T + k + 1) * m + q, i0 = ((int64_t)b * T + k) * m + q;
        double inc = ys[i1] - ys[i0];
        bool clamped = !(inc > 1e-6);
        double rate = clamped ? 1e-6 : inc;
        double o = obs[k * m + q];
        acc += o * std::log(rate) - rate - std::lgamma(o + 1.0);
        if (!clamped && dys)
          for (int p = 0; p < P; ++p) g[p] += (o / rate - 1.0) * (dys[i1 * P + p] - dys[i0 * P + p]);
      }
    lp[b] = acc;
    if (grad) for (int p = 0; p < P; ++p) grad[b * P + p] = g[p];
  }
  return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
