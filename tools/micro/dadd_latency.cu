// dadd_latency.cu — microbenchmark: latency of a DEPENDENT chain of fp64 adds on one warp (what bounds K4's bit-exact
// left-to-right sums: D adds per sum, no reassociation allowed), with and without the shared-memory operand loads K4 uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dadd_latency dadd_latency.cu && ./dadd_latency
#include <cstdio>
#include <cuda_runtime.h>

__global__ void chain(const double* __restrict__ in, double* out, long long* cycles, int n, int mode) {
  __shared__ double sh[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sh[i] = in[i];
  __syncthreads();
  double s = 0.0, t = 0.0;
  const long long t0 = clock64();
  if (mode == 0) {        // registers only: pure DADD latency
    double a = sh[threadIdx.x];
#pragma unroll 16
    for (int i = 0; i < n; i++) s = __dadd_rn(s, a);
  } else if (mode == 1) { // one chain fed by broadcast shared loads (K4's pattern)
#pragma unroll 16
    for (int i = 0; i < n; i++) s = __dadd_rn(s, sh[i & 2047]);
  } else {                // two interleaved chains (dot and ||x||^2)
#pragma unroll 16
    for (int i = 0; i < n; i++) { s = __dadd_rn(s, sh[i & 2047]); t = __dadd_rn(t, sh[(i + 7) & 2047]); }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) { out[mode] = s + t; cycles[mode] = t1 - t0; }
}

int main() {
  double* in; double* out; long long* cyc;
  cudaMalloc(&in, 2048 * 8); cudaMalloc(&out, 64); cudaMalloc(&cyc, 64);
  double h[2048];
  for (int i = 0; i < 2048; i++) h[i] = 1.0 / (i + 1);
  cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
  const int n = 1 << 16;
  for (int rep = 0; rep < 2; rep++)
    for (int mode = 0; mode < 3; mode++) chain<<<1, 32>>>(in, out, cyc, n, mode);
  cudaDeviceSynchronize();
  long long c[3];
  cudaMemcpy(c, cyc, sizeof c, cudaMemcpyDeviceToHost);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("dependent fp64 add chain, one warp: %.2f cycles/add (registers), %.2f (shared operand), %.2f per step (two chains)\n",
         (double)c[0] / n, (double)c[1] / n, (double)c[2] / n);
  printf("=> a 1536-element left-to-right sum costs >= %.1f us at %.2f GHz\n", 1536.0 * c[1] / n / (khz * 1e-3), khz * 1e-6);
  return 0;
}
