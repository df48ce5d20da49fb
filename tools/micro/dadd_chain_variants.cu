// dadd_chain_variants.cu — microbenchmark: code shapes for K4's bit-exact left-to-right sum of 1536 products parked in shared
// memory (one warp, every lane replays the same chain). What is on the critical path besides the DADD latency?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dadd_chain_variants dadd_chain_variants.cu && ./dadd_chain_variants
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 1536;

template <int MODE>
__global__ void chain(const double* __restrict__ in, double* out, long long* cycles) {
  __shared__ __align__(16) double sh[N];
  for (int i = threadIdx.x; i < N; i += blockDim.x) sh[i] = in[i];
  __syncthreads();
  double s = 0.0;
  const long long t0 = clock64();
  if (MODE == 0) {          // K4 today: 128-bit broadcast loads, unroll 16
    const double2* a = reinterpret_cast<const double2*>(sh);
#pragma unroll 16
    for (int i = 0; i < N / 2; i++) { const double2 u = a[i]; s = __dadd_rn(__dadd_rn(s, u.x), u.y); }
  } else if (MODE == 1) {   // 32 products into registers, then 32 adds; the next 32 are loaded before the adds start
    const double2* a = reinterpret_cast<const double2*>(sh);
    double2 r[16], nx[16];
#pragma unroll
    for (int j = 0; j < 16; j++) r[j] = a[j];
    for (int b = 0; b < N / 32; b++) {
      if (b + 1 < N / 32) {
#pragma unroll
        for (int j = 0; j < 16; j++) nx[j] = a[(b + 1) * 16 + j];
      }
#pragma unroll
      for (int j = 0; j < 16; j++) s = __dadd_rn(__dadd_rn(s, r[j].x), r[j].y);
#pragma unroll
      for (int j = 0; j < 16; j++) r[j] = nx[j];
    }
  } else if (MODE == 2) {   // lane-private products in registers (48 per lane), passed along by shuffles: 2 SHFL + DADD per element
    double p[N / 32];
#pragma unroll
    for (int t = 0; t < N / 32; t++) p[t] = sh[t * 32 + (threadIdx.x & 31)];
#pragma unroll
    for (int t = 0; t < N / 32; t++) {
      const int lo = __double2loint(p[t]), hi = __double2hiint(p[t]);
#pragma unroll
      for (int l = 0; l < 32; l++) s = __dadd_rn(s, __hiloint2double(__shfl_sync(0xFFFFFFFFu, hi, l), __shfl_sync(0xFFFFFFFFu, lo, l)));
    }
  } else if (MODE == 3) {   // only lane 0 adds (the other lanes idle): does the fp64 pipe care how many lanes are active?
    if ((threadIdx.x & 31) == 0) {
      const double2* a = reinterpret_cast<const double2*>(sh);
#pragma unroll 16
      for (int i = 0; i < N / 2; i++) { const double2 u = a[i]; s = __dadd_rn(__dadd_rn(s, u.x), u.y); }
    }
  } else if (MODE == 4) {   // 64-bit loads, unroll 32
#pragma unroll 32
    for (int i = 0; i < N; i++) s = __dadd_rn(s, sh[i]);
  } else if (MODE == 5) {   // registers only (the floor): same operand every step
    const double a = sh[threadIdx.x & 31];
#pragma unroll 16
    for (int i = 0; i < N; i++) s = __dadd_rn(s, a);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) { out[MODE] = s; cycles[MODE] = t1 - t0; }
  if (s == 123.456) out[8] = s;
}

int main() {
  double *in, *out; long long* cyc;
  cudaMalloc(&in, N * 8); cudaMalloc(&out, 128); cudaMalloc(&cyc, 128);
  double h[N];
  for (int i = 0; i < N; i++) h[i] = 1.0 / (i + 1);
  cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
  for (int warps : {1, 5}) {
    for (int rep = 0; rep < 2; rep++) {
      chain<0><<<1, 32 * warps>>>(in, out, cyc); chain<1><<<1, 32 * warps>>>(in, out, cyc); chain<2><<<1, 32 * warps>>>(in, out, cyc);
      chain<3><<<1, 32 * warps>>>(in, out, cyc); chain<4><<<1, 32 * warps>>>(in, out, cyc); chain<5><<<1, 32 * warps>>>(in, out, cyc);
    }
    cudaDeviceSynchronize();
    long long c[6]; double o[6];
    cudaMemcpy(c, cyc, sizeof c, cudaMemcpyDeviceToHost);
    cudaMemcpy(o, out, sizeof o, cudaMemcpyDeviceToHost);
    const char* names[6] = {"LDS.128 + 2 DADD, unroll 16 (K4 today)", "32 products preloaded into registers", "lane-private products, 2 SHFL + DADD",
                            "lane 0 only", "LDS.64 + DADD, unroll 32", "registers only (floor)"};
    printf("--- %d warp(s) per CTA, %d-element chain ---\n", warps, N);
    for (int m = 0; m < 6; m++) printf("%-44s %7lld cycles  %.2f per add   sum %.17g\n", names[m], c[m], (double)c[m] / N, o[m]);
  }
  return 0;
}
