// Dependent-issue latency of DFMA / DADD / DMUL / SHFL (64-bit) on the device, one warp, clock64 around a long dependent chain.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void chain(double *out, long long *cyc, double a, double b, int n)
{
  double x = a + threadIdx.x * 1e-9, y = b;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (OP == 0) x = fma(x, y, a);
      if (OP == 1) x = x + y;
      if (OP == 2) x = x * y;
      if (OP == 3) x = __shfl_down_sync(0xffffffffu, x, 1);
      if (OP == 4) { x = fma(x, y, a); x = __shfl_xor_sync(0xffffffffu, x, 1); }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x;
}
template <int OP>
__global__ void indep(double *out, long long *cyc, double a, double b, int n)   // 8 independent chains: issue rate
{
  double x[8];
  for (int k = 0; k < 8; ++k) x[k] = a + k + threadIdx.x * 1e-9;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (OP == 0) x[k] = fma(x[k], b, a);
        if (OP == 3) x[k] = __shfl_down_sync(0xffffffffu, x[k], 1);
      }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  double s = 0; for (int k = 0; k < 8; ++k) s += x[k];
  out[threadIdx.x] = s;
}
int main()
{
  double *d; long long *c, h;
  cudaMalloc(&d, 4096); cudaMalloc(&c, 8);
  const char *names[5] = {"DFMA", "DADD", "DMUL", "SHFL.64 (2 x SHFL.32)", "DFMA + SHFL.64"};
  const int n = 4096;
#define RUN(OP) chain<OP><<<1, 32>>>(d, c, 1.0000001, 0.9999999, n); chain<OP><<<1, 32>>>(d, c, 1.0000001, 0.9999999, n); \
  cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-24s dependent latency %.1f cycles\n", names[OP], (double)h / (16.0 * n));
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4)
  indep<0><<<1, 32>>>(d, c, 1.0000001, 0.9999999, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("DFMA, 8 independent chains, one warp: %.2f cycles per instruction\n", (double)h / (32.0 * n));
  indep<3><<<1, 32>>>(d, c, 1.0000001, 0.9999999, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("SHFL.64, 8 independent, one warp: %.2f cycles per 64-bit shuffle\n", (double)h / (32.0 * n));
  indep<0><<<1, 128>>>(d, c, 1.0000001, 0.9999999, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("DFMA, 8 independent chains, 4 warps (one per scheduler): %.2f cycles per instruction per warp\n", (double)h / (32.0 * n));
  indep<0><<<1, 512>>>(d, c, 1.0000001, 0.9999999, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("DFMA, 8 independent chains, 16 warps: %.2f cycles per instruction per warp\n", (double)h / (32.0 * n));
  indep<3><<<1, 128>>>(d, c, 1.0000001, 0.9999999, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("SHFL.64, 8 independent, 4 warps (one per scheduler): %.2f cycles per 64-bit shuffle per warp\n", (double)h / (32.0 * n));
  indep<3><<<1, 256>>>(d, c, 1.0000001, 0.9999999, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("SHFL.64, 8 independent, 8 warps: %.2f cycles per 64-bit shuffle per warp\n", (double)h / (32.0 * n));
  indep<3><<<1, 512>>>(d, c, 1.0000001, 0.9999999, n); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("SHFL.64, 8 independent, 16 warps: %.2f cycles per 64-bit shuffle per warp\n", (double)h / (32.0 * n));
  return 0;
}
