// tsit5.cuh -- Tsitouras 5(4) tableau, dense output and step-size controller pieces as device
// code.  Restates what diffrax 0.7 executes for the reference's
// diffeqsolve(ODETerm, Tsit5(), ..., PIDController(rtol, atol)) call
// (reference src/dynode/simulation/odes.py:125-144, config/params.py:28-34).
#pragma once
#include <cuda_runtime.h>

namespace dynode {
namespace tsit5 {

// The tableau lives in constant memory: after unrolling every use is a c[bank][offset] operand
// of DFMA/DMUL, so no instruction is spent materialising 64-bit literals (ncu r1: literal
// immediates cost 16% of all issued instructions as UMOV pairs).
namespace hv {
constexpr double c2 = 161.0 / 1000.0;
constexpr double c3 = 327.0 / 1000.0;
constexpr double c4 = 9.0 / 10.0;
constexpr double c5 = 0.9800255409045096857298102862870245954942137979563024768854764293221195950761080302604;
constexpr double a21 = 161.0 / 1000.0;
constexpr double a31 = -0.8480655492356988544426874250230774675121177393430391537369234245294192976164141156943e-2;
constexpr double a32 = 0.3354806554923569885444268742502307746751211773934303915373692342452941929761641411569;
constexpr double a41 = 2.897153057105493432130432594192938764924887287701866490314866693455023795137503079289;
constexpr double a42 = -6.359448489975074843148159912383825625952700647415626703305928850207288721235210244366;
constexpr double a43 = 4.362295432869581411017727318190886861027813359713760212991062156752264926097707165077;
constexpr double a51 = 5.325864828439256604428877920840511317836476253097040101202360397727981648835607691791;
constexpr double a52 = -11.74888356406282787774717033978577296188744178259862899288666928009020615663593781589;
constexpr double a53 = 7.495539342889836208304604784564358155658679161518186721010132816213648793440552049753;
constexpr double a54 = -0.9249506636175524925650207933207191611349983406029535244034750452930469056411389539635e-1;
constexpr double a61 = 5.861455442946420028659251486982647890394337666164814434818157239052507339770711679748;
constexpr double a62 = -12.92096931784710929170611868178335939541780751955743459166312250439928519268343184452;
constexpr double a63 = 8.159367898576158643180400794539253485181918321135053305748355423955009222648673734986;
constexpr double a64 = -0.7158497328140099722453054252582973869127213147363544882721139659546372402303777878835e-1;
constexpr double a65 = -0.2826905039406838290900305721271224146717633626879770007617876201276764571291579142206e-1;
constexpr double a71 = 0.9646076681806522951816731316512876333711995238157997181903319145764851595234062815396e-1;
constexpr double a72 = 1.0 / 100.0;
constexpr double a73 = 0.4798896504144995747752495322905965199130404621990332488332634944254542060153074523509;
constexpr double a74 = 1.379008574103741893192274821856872770756462643091360525934940067397245698027561293331;
constexpr double a75 = -3.290069515436080679901047585711363850115683290894936158531296799594813811049925401677;
constexpr double a76 = 2.324710524099773982415355918398765796109060233222962411944060046314465391054716027841;
constexpr double e1 = a71 - 0.9468075576583945807478876255758922856117527357724631226139574065785592789071067303271e-1;
constexpr double e2 = a72 - 0.9183565540343253096776363936645313759813746240984095238905939532922955247253608687270e-2;
constexpr double e3 = a73 - 0.4877705284247615707855642599631228241516691959761363774365216240304071651579571959813;
constexpr double e4 = a74 - 1.234297566930478985655109673884237654035539930748192848315425833500484878378061439761;
constexpr double e5 = a75 - -2.707712349983525454881109975059321670689605166938197378763992255714444407154902012702;
constexpr double e6 = a76 - 1.866628418170587035753719399566211498666255505244122593996591602841258328965767580089;
constexpr double e7 = 0.0 - 1.0 / 66.0;
}  // namespace hv
enum TabIdx { I_c2, I_c3, I_c4, I_c5, I_a21, I_a31, I_a32, I_a41, I_a42, I_a43, I_a51, I_a52, I_a53, I_a54, I_a61, I_a62, I_a63, I_a64, I_a65, I_a71, I_a72, I_a73, I_a74, I_a75, I_a76, I_e1, I_e2, I_e3, I_e4, I_e5, I_e6, I_e7 };
static __constant__ double kTab[] = {
    hv::c2,
    hv::c3,
    hv::c4,
    hv::c5,
    hv::a21,
    hv::a31,
    hv::a32,
    hv::a41,
    hv::a42,
    hv::a43,
    hv::a51,
    hv::a52,
    hv::a53,
    hv::a54,
    hv::a61,
    hv::a62,
    hv::a63,
    hv::a64,
    hv::a65,
    hv::a71,
    hv::a72,
    hv::a73,
    hv::a74,
    hv::a75,
    hv::a76,
    hv::e1,
    hv::e2,
    hv::e3,
    hv::e4,
    hv::e5,
    hv::e6,
    hv::e7};
#define T5_c2 (::dynode::tsit5::kTab[::dynode::tsit5::I_c2])
#define T5_c3 (::dynode::tsit5::kTab[::dynode::tsit5::I_c3])
#define T5_c4 (::dynode::tsit5::kTab[::dynode::tsit5::I_c4])
#define T5_c5 (::dynode::tsit5::kTab[::dynode::tsit5::I_c5])
#define T5_a21 (::dynode::tsit5::kTab[::dynode::tsit5::I_a21])
#define T5_a31 (::dynode::tsit5::kTab[::dynode::tsit5::I_a31])
#define T5_a32 (::dynode::tsit5::kTab[::dynode::tsit5::I_a32])
#define T5_a41 (::dynode::tsit5::kTab[::dynode::tsit5::I_a41])
#define T5_a42 (::dynode::tsit5::kTab[::dynode::tsit5::I_a42])
#define T5_a43 (::dynode::tsit5::kTab[::dynode::tsit5::I_a43])
#define T5_a51 (::dynode::tsit5::kTab[::dynode::tsit5::I_a51])
#define T5_a52 (::dynode::tsit5::kTab[::dynode::tsit5::I_a52])
#define T5_a53 (::dynode::tsit5::kTab[::dynode::tsit5::I_a53])
#define T5_a54 (::dynode::tsit5::kTab[::dynode::tsit5::I_a54])
#define T5_a61 (::dynode::tsit5::kTab[::dynode::tsit5::I_a61])
#define T5_a62 (::dynode::tsit5::kTab[::dynode::tsit5::I_a62])
#define T5_a63 (::dynode::tsit5::kTab[::dynode::tsit5::I_a63])
#define T5_a64 (::dynode::tsit5::kTab[::dynode::tsit5::I_a64])
#define T5_a65 (::dynode::tsit5::kTab[::dynode::tsit5::I_a65])
#define T5_a71 (::dynode::tsit5::kTab[::dynode::tsit5::I_a71])
#define T5_a72 (::dynode::tsit5::kTab[::dynode::tsit5::I_a72])
#define T5_a73 (::dynode::tsit5::kTab[::dynode::tsit5::I_a73])
#define T5_a74 (::dynode::tsit5::kTab[::dynode::tsit5::I_a74])
#define T5_a75 (::dynode::tsit5::kTab[::dynode::tsit5::I_a75])
#define T5_a76 (::dynode::tsit5::kTab[::dynode::tsit5::I_a76])
#define T5_e1 (::dynode::tsit5::kTab[::dynode::tsit5::I_e1])
#define T5_e2 (::dynode::tsit5::kTab[::dynode::tsit5::I_e2])
#define T5_e3 (::dynode::tsit5::kTab[::dynode::tsit5::I_e3])
#define T5_e4 (::dynode::tsit5::kTab[::dynode::tsit5::I_e4])
#define T5_e5 (::dynode::tsit5::kTab[::dynode::tsit5::I_e5])
#define T5_e6 (::dynode::tsit5::kTab[::dynode::tsit5::I_e6])
#define T5_e7 (::dynode::tsit5::kTab[::dynode::tsit5::I_e7])

// Dense output in monomial form.  diffrax evaluates y(theta) = y0 + sum_i b_i(theta) k_i with the
// factored quartics b_i below (dense_weights); expanding them once (200-bit arithmetic on the
// double-rounded literals) gives b_i(theta) = sum_{m=1..4} kDense[i][m-1] theta^m, so a step forms
//   Q_m = sum_i kDense[i][m-1] f_i     (once per accepted step)
//   y(theta) = y0 + h*theta*(Q_1 + theta*(Q_2 + theta*(Q_3 + theta*Q_4)))   (4 FMAs per saved value)
// instead of 7 weights + 7 FMAs per saved value.  theta = 0 returns y0 exactly, as the factored form.
static __constant__ double kDense[7][4] = {
    {0.9999999999999998421123581, -2.763706197274825756327939, 2.913255461821912638667277, -1.053088497729021577598019},
    {0.0, 0.1316999999999999922991787, -0.2233999999999999781239559, 0.1016999999999999987343458},
    {0.0, 3.930296236894751358945808, -5.941033872131504636636415, 2.490627285651252798004407},
    {0.0, -12.41107716693367686910002, 30.3381886302823205846932, -16.54810288924490180306748},
    {0.0, 37.5093134165110404482129, -88.17890489476640577512085, 47.37952196281928252119542},
    {0.0, -27.89652628919728582533808, 65.09189467479367010895699, -34.87065786149661050785653},
    {0.0, 1.499999999999999944488849, -3.999999999999999944488849, 2.5}};

// 4th-order dense output weights b_i(theta) (diffrax _Tsit5Interpolation.evaluate)
__device__ __forceinline__ void dense_weights(double t, double (&b)[7]) {
  const double t2 = t * t;
  b[0] = -1.0530884977290216 * t * (t - 1.3299890189751412) * (t2 - 1.4364028541716351 * t + 0.7139816917074209);
  b[1] = 0.1017 * t2 * (t2 - 2.1966568338249754 * t + 1.2949852507374631);
  b[2] = 2.490627285651252793 * t2 * (t2 - 2.38535645472061657 * t + 1.57803468208092486);
  b[3] = -16.54810288924490272 * (t - 1.21712927295533244) * (t - 0.61620406037800089) * t2;
  b[4] = 47.37952196281928122 * (t - 1.203071208372362603) * (t - 0.658047292653547382) * t2;
  b[5] = -34.87065786149660974 * (t - 1.2) * (t - 0.666666666666666667) * t2;
  b[6] = 2.5 * (t - 1.0) * (t - 0.6) * t2;
}

// I-controller factor: clip(0.9 * (1/err)^(1/5), keep ? 1 : 0.2, 10)   (PIDController defaults:
// pcoeff=0, icoeff=1, dcoeff=0, safety=0.9, factormin=0.2, factormax=10, error_order=5; accepted
// steps never shrink).  x^(-1/5) is seeded from the SFU in fp32 and polished by two Newton steps
// in fp64 (z <- z*(1.2 - 0.2*x*z^5), quadratic): ~2 ulp, at a fraction of the cost of pow().
// Outside [5e-6, 2000] the clip decides the result, so no root is taken (err==0 -> 10, as
// 0.9*inf clipped; err=inf/NaN -> reject with 0.2).
__device__ __forceinline__ double controller_factor(double err, bool keep) {
  double factor;
  if (err <= 5e-6) {
    factor = 10.0;
  } else if (!(err < 2000.0)) {
    factor = 0.2;
  } else {
    double z = (double)exp2f(-0.2f * __log2f((float)err));
    double z2 = z * z;
    double z5 = z2 * z2 * z;
    z = z * fma(-0.2 * err, z5, 1.2);
    z2 = z * z;
    z5 = z2 * z2 * z;
    z = z * fma(-0.2 * err, z5, 1.2);
    factor = 0.9 * z;
  }
  const double lo = keep ? 1.0 : 0.2;
  return fmin(fmax(factor, lo), 10.0);
}

// Same factor from the SQUARED error norm: err^(-1/5) = (err^2)^(-1/10), so the step needs no
// square root.  Seed from the SFU in fp32, two Newton steps z <- z*(1.1 - 0.1*x*z^10) in fp64.
// Clip thresholds in err^2: err <= 5e-6 -> 2.5e-11, err >= 2000 -> 4e6.
__device__ __forceinline__ double controller_factor_sq(double err2, bool keep) {
  // Branch-free: the root is always taken on err^2 clamped into [2.5e-11, 4e6]; at the clamp ends
  // 0.9 * x^(-1/10) is 10.3 resp. 0.197, i.e. already outside [lo, 10], so the final clip returns exactly
  // what the three-way branch did (NaN / inf -> upper clamp -> 0.2 on the rejected step).
  const double x = !(err2 < 4.0e6) ? 4.0e6 : fmax(err2, 2.5e-11);
  double z = (double)exp2f(-0.1f * __log2f((float)x));
  const double mx = -0.1 * x;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const double z2 = z * z;
    const double z4 = z2 * z2;
    const double z5 = z4 * z;
    z = z * fma(mx, z5 * z5, 1.1);
  }
  const double lo = keep ? 1.0 : 0.2;
  return fmin(fmax(0.9 * z, lo), 10.0);
}

}  // namespace tsit5
}  // namespace dynode
