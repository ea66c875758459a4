// dual.cuh -- forward-mode dual numbers held in registers: value + P tangents.
// Carrying tangents through the accepted Tsit5 steps differentiates the discrete scheme with the
// step sequence frozen, which is what the reference's reverse-mode pass through diffrax computes
// (stop_gradient on the controller factor and on the automatic dt0).
#pragma once
#include <cuda_runtime.h>

namespace dynode {

template <int P>
struct Dual {
  double v;
  double d[P];
};
template <>
struct Dual<0> {
  double v;
};

#define DYN_DI __device__ __forceinline__

template <int P> DYN_DI Dual<P> make_dual(double x) {
  Dual<P> r; r.v = x;
  if constexpr (P > 0) {
#pragma unroll
    for (int p = 0; p < P; ++p) r.d[p] = 0.0;
  }
  return r;
}
template <int P> DYN_DI Dual<P> operator+(const Dual<P>& a, const Dual<P>& b) {
  Dual<P> r; r.v = a.v + b.v;
  if constexpr (P > 0) {
#pragma unroll
    for (int p = 0; p < P; ++p) r.d[p] = a.d[p] + b.d[p];
  }
  return r;
}
template <int P> DYN_DI Dual<P> operator-(const Dual<P>& a, const Dual<P>& b) {
  Dual<P> r; r.v = a.v - b.v;
  if constexpr (P > 0) {
#pragma unroll
    for (int p = 0; p < P; ++p) r.d[p] = a.d[p] - b.d[p];
  }
  return r;
}
template <int P> DYN_DI Dual<P> operator*(const Dual<P>& a, const Dual<P>& b) {
  Dual<P> r; r.v = a.v * b.v;
  if constexpr (P > 0) {
#pragma unroll
    for (int p = 0; p < P; ++p) r.d[p] = fma(a.d[p], b.v, a.v * b.d[p]);
  }
  return r;
}
template <int P> DYN_DI Dual<P> operator*(double a, const Dual<P>& b) {
  Dual<P> r; r.v = a * b.v;
  if constexpr (P > 0) {
#pragma unroll
    for (int p = 0; p < P; ++p) r.d[p] = a * b.d[p];
  }
  return r;
}
template <int P> DYN_DI Dual<P> operator/(const Dual<P>& a, const Dual<P>& b) {
  Dual<P> r; r.v = a.v / b.v;
  if constexpr (P > 0) {
    const double inv = 1.0 / b.v;
#pragma unroll
    for (int p = 0; p < P; ++p) r.d[p] = fma(-r.v, b.d[p], a.d[p]) * inv;
  }
  return r;
}
// Branch-free FP64 reciprocal / division: MUFU.RCP64H seed (~2^-23) + two Newton steps + one
// residual correction on the quotient (<= 1 ulp for normal, non-zero divisors -- populations and
// error scales here).  The compiler's IEEE division costs ~2x the instructions plus a
// BSSY/BRA/CALL slow path per use (ncu r1).
DYN_DI double rcp_fast(double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  return r;
}
// one Newton step: ~2^-46 relative error, for quantities that only enter a norm
DYN_DI double rcp_fast1(double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  return fma(r, fma(-b, r, 1.0), r);
}
DYN_DI double div_fast(double a, double b) {
  const double r = rcp_fast(b);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
}
template <int P> DYN_DI Dual<P> ddiv_fast(const Dual<P>& a, const Dual<P>& b) {
  Dual<P> r;
  const double inv = rcp_fast(b.v);
  const double q = a.v * inv;
  r.v = fma(fma(-b.v, q, a.v), inv, q);
  if constexpr (P > 0) {
#pragma unroll
    for (int p = 0; p < P; ++p) r.d[p] = fma(-r.v, b.d[p], a.d[p]) * inv;
  }
  return r;
}
template <int P> DYN_DI Dual<P> drcp_fast(const Dual<P>& b) {
  Dual<P> r;
  r.v = rcp_fast(b.v);
  if constexpr (P > 0) {
    const double m = -r.v * r.v;
#pragma unroll
    for (int p = 0; p < P; ++p) r.d[p] = m * b.d[p];
  }
  return r;
}
template <int P> DYN_DI Dual<P> dneg(const Dual<P>& a) {
  Dual<P> r; r.v = -a.v;
  if constexpr (P > 0) {
#pragma unroll
    for (int p = 0; p < P; ++p) r.d[p] = -a.d[p];
  }
  return r;
}
// a*b - c   and   c - a*b   for dual a, b, c (one FMA on the value, two per tangent)
template <int P> DYN_DI Dual<P> dmsub(const Dual<P>& a, const Dual<P>& b, const Dual<P>& c) {
  Dual<P> r; r.v = fma(a.v, b.v, -c.v);
  if constexpr (P > 0) {
#pragma unroll
    for (int p = 0; p < P; ++p) r.d[p] = fma(a.d[p], b.v, fma(a.v, b.d[p], -c.d[p]));
  }
  return r;
}
template <int P> DYN_DI Dual<P> dnmadd(const Dual<P>& a, const Dual<P>& b, const Dual<P>& c) {
  Dual<P> r; r.v = fma(-a.v, b.v, c.v);
  if constexpr (P > 0) {
#pragma unroll
    for (int p = 0; p < P; ++p) r.d[p] = fma(-a.d[p], b.v, fma(-a.v, b.d[p], c.d[p]));
  }
  return r;
}
// r = a*b + c
template <int P> DYN_DI Dual<P> dfma(double a, const Dual<P>& b, const Dual<P>& c) {
  Dual<P> r; r.v = fma(a, b.v, c.v);
  if constexpr (P > 0) {
#pragma unroll
    for (int p = 0; p < P; ++p) r.d[p] = fma(a, b.d[p], c.d[p]);
  }
  return r;
}
template <int P> DYN_DI Dual<P> dual_shfl(const Dual<P>& a, int src) {
  Dual<P> r; r.v = __shfl_sync(0xffffffffu, a.v, src);
  if constexpr (P > 0) {
#pragma unroll
    for (int p = 0; p < P; ++p) r.d[p] = __shfl_sync(0xffffffffu, a.d[p], src);
  }
  return r;
}
// log of a double >= 1e-6 (the clamped incidence of the fused log-likelihood): the fdlibm algorithm (exponent split, s = f/(2+f), degree-7 polynomial in
// s^2; ~1-2 ulp as evaluated here) with every constant in c[][].  The libm call it
// replaces in the fused log-likelihood's save pass is ~45 instructions plus 45 UMOVs of literals, per pass and slot.
static __constant__ double kLogC[9] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
                                       2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
                                       1.479819860511658591e-01, 6.93147180369123816490e-01, 1.90821492927058770002e-10};
DYN_DI double log_fast(double x) {
  if (!(x < 1e300)) return x;  // +inf / NaN pass through; callers guarantee x >= 1e-6 (the clamped incidence)
  int hi = __double2hiint(x);
  int e = (hi >> 20) - 1023;
  double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
  const double f = m - 1.0;
  // The caller's accumulation waits for this value at two warps per scheduler, so the dependent chain is kept
  // short: the quotient is reciprocal x numerator without the correction step (2^-52 relative on s: 1e-16 on the
  // result), and the polynomial is evaluated in Estrin form (4 dependent FMAs instead of 7).
  const double s = f * rcp_fast(2.0 + f);
  const double z = s * s;
  const double z2 = z * z;
  const double q01 = fma(kLogC[1], z, kLogC[0]);
  const double q23 = fma(kLogC[3], z, kLogC[2]);
  const double q45 = fma(kLogC[5], z, kLogC[4]);
  const double q456 = fma(kLogC[6], z2, q45);
  double R = fma(z2, fma(z2, q456, q23), q01);
  R *= z;
  const double hfsq = 0.5 * f * f;
  const double dk = (double)e;
  return dk * kLogC[7] - ((hfsq - fma(s, hfsq + R, dk * kLogC[8])) - f);
}
template <int P> DYN_DI Dual<P> dual_log(const Dual<P>& a) {
  Dual<P> r; r.v = log_fast(a.v);
  if constexpr (P > 0) {
    const double inv = rcp_fast(a.v);
#pragma unroll
    for (int p = 0; p < P; ++p) r.d[p] = a.d[p] * inv;
  }
  return r;
}

}  // namespace dynode
