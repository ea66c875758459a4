// ppl_kernels.cu -- the bijector of a latent site and its log-Jacobian in one elementwise kernel each way
// (include/dynode_b200_ppl.h).  Restates numpyro's biject_to(interval / greater_than / less_than) =
// SigmoidTransform / ExpTransform followed by AffineTransform, with SigmoidTransform.log_abs_det_jacobian
// = -softplus(z) - softplus(-z) (what the reference's NUTS evaluates around every model call,
// src/dynode/infer/inference.py:149-163).
#include <cuda_runtime.h>
#include <math_constants.h>

#include "../../include/dynode_b200_ppl.h"

namespace dynode {
int fail_msg(const char* fmt, ...);  // capi.cu

namespace {

__device__ __forceinline__ double softplus(double v) {  // log(1 + e^v) without overflow
  return fmax(v, 0.0) + log1p(exp(-fabs(v)));
}
__device__ __forceinline__ double sigmoid(double v) {
  const double e = exp(-fabs(v));
  const double s = 1.0 / (1.0 + e);  // sigmoid(|v|)
  return v >= 0.0 ? s : e * s;
}

__global__ void __launch_bounds__(256) bijector_kernel(int kind, int64_t n, const double* __restrict__ z, double a,
                                                        double b, double* __restrict__ x, double* __restrict__ ladj) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const double logb = (kind == DYNODE_BIJ_INTERVAL) ? log(fabs(b)) : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double v = z[i];
    if (kind == DYNODE_BIJ_INTERVAL) {
      x[i] = fma(b, sigmoid(v), a);
      ladj[i] = logb - softplus(v) - softplus(-v);
    } else {
      const double e = exp(v);
      x[i] = (kind == DYNODE_BIJ_GREATER_THAN) ? a + e : a - e;
      ladj[i] = v;
    }
  }
}

__global__ void __launch_bounds__(256) bijector_vjp_kernel(int kind, int64_t n, const double* __restrict__ z, double b,
                                                            const double* __restrict__ gx,
                                                            const double* __restrict__ gl, double* __restrict__ gz) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double v = z[i];
    if (kind == DYNODE_BIJ_INTERVAL) {
      const double s = sigmoid(v);
      // dx/dz = b s (1 - s);  d/dz (-softplus(z) - softplus(-z)) = 1 - 2 s
      gz[i] = fma(gx[i], b * s * (1.0 - s), gl[i] * (1.0 - 2.0 * s));
    } else {
      const double e = exp(v);
      gz[i] = fma(gx[i], (kind == DYNODE_BIJ_GREATER_THAN) ? e : -e, gl[i]);
    }
  }
}

int grid_for(int64_t n) {
  const int64_t g = (n + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

}  // namespace
}  // namespace dynode

using namespace dynode;

extern "C" {

int dynode_bijector_f64(int32_t kind, int64_t n, const double* z, double a, double b, double* x, double* ladj,
                        void* stream) {
  if (kind < 0 || kind > DYNODE_BIJ_LESS_THAN) return fail_msg("unknown bijector kind %d", kind);
  if (n < 0 || (n > 0 && (!z || !x || !ladj))) return fail_msg("bijector: null buffer");
  if (n == 0) return 0;
  bijector_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(kind, n, z, a, b, x, ladj);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail_msg("bijector launch failed: %s", cudaGetErrorString(e));
}

int dynode_bijector_vjp_f64(int32_t kind, int64_t n, const double* z, double b, const double* gx, const double* gl,
                            double* gz, void* stream) {
  if (kind < 0 || kind > DYNODE_BIJ_LESS_THAN) return fail_msg("unknown bijector kind %d", kind);
  if (n < 0 || (n > 0 && (!z || !gx || !gl || !gz))) return fail_msg("bijector vjp: null buffer");
  if (n == 0) return 0;
  bijector_vjp_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(kind, n, z, b, gx, gl, gz);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail_msg("bijector vjp launch failed: %s", cudaGetErrorString(e));
}

}  // extern "C"
