// Dev probe: dependent-issue latency of the FP64 pipe on this GPU (one warp, clock64 around N dependent
// operations), with 1 / 2 / 4 / 8 independent chains per thread, and of a double shuffle.  Build and run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_latency scripts/probes/fp64_latency.cu && /tmp/fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_chain(double* out, long long* cycles, int n, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) x[k] = 1.0 + 1e-9 * (threadIdx.x + k);
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += x[k];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

__global__ void shfl_chain(double* out, long long* cycles, int n) {
  double x = 1.0 + threadIdx.x;
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
  const long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

__global__ void shfl_fma_chain(double* out, long long* cycles, int n, double a) {
  double x = 1.0 + threadIdx.x;
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = fma(__shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31), a, x);
  const long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

int main() {
  double* out;
  long long *cyc, h;
  cudaMalloc(&out, 1024 * sizeof(double));
  cudaMalloc(&cyc, sizeof(long long));
  const int n = 4096;
#define RUN(ILP, WARPS)                                                                       \
  dfma_chain<ILP><<<1, 32 * WARPS>>>(out, cyc, n, 1.0000001, 1e-9);                            \
  dfma_chain<ILP><<<1, 32 * WARPS>>>(out, cyc, n, 1.0000001, 1e-9);                            \
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);                                     \
  printf("DFMA  chains/thread %d  warps %2d : %.2f cycles per dependent DFMA step (%.2f per DFMA issued by the warp)\n", ILP, WARPS, (double)h / n, (double)h / n / ILP);
  RUN(1, 1) RUN(2, 1) RUN(4, 1) RUN(8, 1) RUN(1, 4) RUN(1, 8) RUN(1, 12) RUN(1, 16) RUN(4, 12) RUN(5, 12)
  shfl_chain<<<1, 32>>>(out, cyc, n);
  shfl_chain<<<1, 32>>>(out, cyc, n);
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("SHFL.64 dependent chain: %.2f cycles per double shuffle\n", (double)h / n);
  shfl_fma_chain<<<1, 32>>>(out, cyc, n, 1e-9);
  shfl_fma_chain<<<1, 32>>>(out, cyc, n, 1e-9);
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("SHFL.64 -> DFMA dependent chain: %.2f cycles per pair\n", (double)h / n);
  return 0;
}
