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
template <int P> DYN_DI Dual<P> dual_log(const Dual<P>& a) {
  Dual<P> r; r.v = log(a.v);
  if constexpr (P > 0) {
    const double inv = 1.0 / a.v;
#pragma unroll
    for (int p = 0; p < P; ++p) r.d[p] = a.d[p] * inv;
  }
  return r;
}

}  // namespace dynode
