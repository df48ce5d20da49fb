// k1_stream.cu — K1: batch-1 (small-batch) cosine scoring fused with top-K' selection.
//
// Replaces the hot loops of getTopKEmbeddings / similarity(cosine) (llamaindex,
// reached from src/lib/hybrid-search.ts:223-224): score every row, keep the best.
//
// Roofline: HBM. Algorithmic bytes per launch = rows * ld * sizeof(elem) (the corpus is
// streamed exactly once per query); everything else (query, candidate lists) is o(1%).
//   - every warp owns whole rows (row = global warp id + i * total warps), so a warp
//     reads 6 KB (fp32, D=1536) of contiguous memory per row with 128-bit loads that
//     bypass L1 (ld.global.nc.L1::no_allocate); all of a row's loads are issued before
//     the first FMA so each warp keeps a full row in flight
//   - dot(q,x) and ||x||^2 are accumulated from the same registers: the cosine
//     normalisation costs no extra HBM traffic (the reference recomputes both norms
//     per row; ||q|| is a positive per-query constant and cannot change the order)
//   - the query lives in shared memory, read as conflict-free 128-bit LDS
//   - selection: per-warp sorted list of K' packed keys in shared memory guarded by a
//     register threshold, so the common case is one 64-bit compare per row
//   - output: one sorted K' list per CTA; K3 merges them
//
// fp32 accumulation decides only WHICH K' rows survive; K4 rescored them exactly.
#include "common.cuh"

#include <cuda_fp16.h>
#include <stdio.h>

namespace {

constexpr int K1_THREADS = 512;
constexpr int K1_WARPS = K1_THREADS / 32;

// diagnostics (RAGERA_SMALL_PROF=1): cycles per phase of the single-query kernel, summed over CTAs
__device__ int g_k1_prof_on;
__device__ unsigned long long g_k1_prof[4];  // query staging, row loop (until the CTA's last warp is done), CTA merge, CTAs

__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

__device__ __forceinline__ float finish_score(float dot, float nrm) {
  float s = dot * rsqrtf(nrm);
  // zero-norm / non-finite rows: the reference yields NaN; defined here as never selected
  return (nrm > 0.0f && s == s) ? s : -INFINITY;
}

// ---- fp32 corpus: U float4 per lane per sub-chunk, nsub sub-chunks per row ----------
template <int U>
__global__ void __launch_bounds__(K1_THREADS, 1)
k1_stream_f32(const float* __restrict__ X, uint32_t n_rows, uint32_t ld, int nsub,
              const float* __restrict__ Q, uint32_t kp, uint32_t parts, uint64_t* __restrict__ partial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* qs = reinterpret_cast<float4*>(smem_raw);
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem_raw + (size_t)ld * sizeof(float));

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t b = blockIdx.y;
  const bool prof = g_k1_prof_on != 0;
  const long long pt0 = prof ? clock64() : 0;
  const float4* q4 = reinterpret_cast<const float4*>(Q + (size_t)b * ld);
  for (uint32_t i = threadIdx.x; i < ld / 4; i += K1_THREADS) qs[i] = q4[i];
  uint64_t* mylist = lists + (size_t)warp * kp;
  for (uint32_t i = lane; i < kp; i += 32) mylist[i] = 0ull;
  __syncthreads();
  const long long pt1 = prof ? clock64() : 0;

  uint64_t thresh = 0ull;
  const uint32_t stride = gridDim.x * K1_WARPS;
  for (uint32_t row = blockIdx.x * K1_WARPS + warp; row < n_rows; row += stride) {
    const float4* xr = reinterpret_cast<const float4*>(X + (size_t)row * ld);
    float d0 = 0.f, d1 = 0.f, n0 = 0.f, n1 = 0.f;
    for (int s = 0; s < nsub; s++) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = ldg_stream_f4(xr + (s * U + u) * 32 + lane);
#pragma unroll
      for (int u = 0; u < U; u++) {
        const float4 qq = qs[(s * U + u) * 32 + lane];
        d0 = fmaf(v[u].x, qq.x, d0); d1 = fmaf(v[u].y, qq.y, d1);
        d0 = fmaf(v[u].z, qq.z, d0); d1 = fmaf(v[u].w, qq.w, d1);
        n0 = fmaf(v[u].x, v[u].x, n0); n1 = fmaf(v[u].y, v[u].y, n1);
        n0 = fmaf(v[u].z, v[u].z, n0); n1 = fmaf(v[u].w, v[u].w, n1);
      }
    }
    const float dot = warp_sum(d0 + d1), nrm = warp_sum(n0 + n1);
    const uint64_t key = rag_pack_key(finish_score(dot, nrm), row);
    if (key > thresh) warp_list_insert(mylist, kp, key, lane, thresh);
  }
  __syncthreads();
  const long long pt2 = prof ? clock64() : 0;
  if (warp == 0)
    warp_merge_lists(lists, K1_WARPS, kp, kp, partial + ((size_t)b * parts + blockIdx.x) * kp, lane);
  if (prof && threadIdx.x == 0) {
    atomicAdd(&g_k1_prof[0], (unsigned long long)(pt1 - pt0));
    atomicAdd(&g_k1_prof[1], (unsigned long long)(pt2 - pt1));
    atomicAdd(&g_k1_prof[2], (unsigned long long)(clock64() - pt2));
    atomicAdd(&g_k1_prof[3], 1ull);
  }
}

// ---- bf16 corpus: U 16-byte loads (8 elements) per lane per sub-chunk, 2 rows in flight ----
__device__ __forceinline__ void bf16x8_fma(const uint4 v, const float4 qa, const float4 qb, float& d0,
                                           float& d1, float& n0, float& n1) {
  float x0 = __uint_as_float(v.x << 16), x1 = __uint_as_float(v.x & 0xFFFF0000u);
  float x2 = __uint_as_float(v.y << 16), x3 = __uint_as_float(v.y & 0xFFFF0000u);
  float x4 = __uint_as_float(v.z << 16), x5 = __uint_as_float(v.z & 0xFFFF0000u);
  float x6 = __uint_as_float(v.w << 16), x7 = __uint_as_float(v.w & 0xFFFF0000u);
  d0 = fmaf(x0, qa.x, d0); d1 = fmaf(x1, qa.y, d1); d0 = fmaf(x2, qa.z, d0); d1 = fmaf(x3, qa.w, d1);
  d0 = fmaf(x4, qb.x, d0); d1 = fmaf(x5, qb.y, d1); d0 = fmaf(x6, qb.z, d0); d1 = fmaf(x7, qb.w, d1);
  n0 = fmaf(x0, x0, n0); n1 = fmaf(x1, x1, n1); n0 = fmaf(x2, x2, n0); n1 = fmaf(x3, x3, n1);
  n0 = fmaf(x4, x4, n0); n1 = fmaf(x5, x5, n1); n0 = fmaf(x6, x6, n0); n1 = fmaf(x7, x7, n1);
}

template <int U>
__global__ void __launch_bounds__(K1_THREADS, 1)
k1_stream_bf16(const __nv_bfloat16* __restrict__ X, uint32_t n_rows, uint32_t ld, int nsub,
               const float* __restrict__ Q, uint32_t kp, uint32_t parts, uint64_t* __restrict__ partial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // query split into two planes so each 128-bit LDS has a 16-byte lane stride:
  // qa[i] = q[8i..8i+3], qb[i] = q[8i+4..8i+7]
  float4* qa = reinterpret_cast<float4*>(smem_raw);
  float4* qb = qa + ld / 8;
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem_raw + (size_t)ld * sizeof(float));

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t b = blockIdx.y;
  const float4* q4 = reinterpret_cast<const float4*>(Q + (size_t)b * ld);
  for (uint32_t i = threadIdx.x; i < ld / 4; i += K1_THREADS) {
    if (i & 1) qb[i >> 1] = q4[i]; else qa[i >> 1] = q4[i];
  }
  uint64_t* mylist = lists + (size_t)warp * kp;
  for (uint32_t i = lane; i < kp; i += 32) mylist[i] = 0ull;
  __syncthreads();

  uint64_t thresh = 0ull;
  const uint32_t stride = gridDim.x * K1_WARPS;
  const uint32_t first = blockIdx.x * K1_WARPS + warp;
  for (uint32_t row = first; row < n_rows; row += 2 * stride) {
    const uint32_t row2 = row + stride;
    const bool has2 = row2 < n_rows;
    const uint4* xr0 = reinterpret_cast<const uint4*>(X + (size_t)row * ld);
    const uint4* xr1 = reinterpret_cast<const uint4*>(X + (size_t)(has2 ? row2 : row) * ld);
    float d0 = 0.f, d1 = 0.f, n0 = 0.f, n1 = 0.f, e0 = 0.f, e1 = 0.f, m0 = 0.f, m1 = 0.f;
    for (int s = 0; s < nsub; s++) {
      uint4 v[U], w[U];
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = ldg_stream_u4(xr0 + (s * U + u) * 32 + lane);
#pragma unroll
      for (int u = 0; u < U; u++) w[u] = ldg_stream_u4(xr1 + (s * U + u) * 32 + lane);
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int i = (s * U + u) * 32 + lane;
        const float4 a = qa[i], c = qb[i];
        bf16x8_fma(v[u], a, c, d0, d1, n0, n1);
        bf16x8_fma(w[u], a, c, e0, e1, m0, m1);
      }
    }
    const float dotA = warp_sum(d0 + d1), nrmA = warp_sum(n0 + n1);
    const float dotB = warp_sum(e0 + e1), nrmB = warp_sum(m0 + m1);
    const uint64_t keyA = rag_pack_key(finish_score(dotA, nrmA), row);
    if (keyA > thresh) warp_list_insert(mylist, kp, keyA, lane, thresh);
    if (has2) {
      const uint64_t keyB = rag_pack_key(finish_score(dotB, nrmB), row2);
      if (keyB > thresh) warp_list_insert(mylist, kp, keyB, lane, thresh);
    }
  }
  __syncthreads();
  if (warp == 0)
    warp_merge_lists(lists, K1_WARPS, kp, kp, partial + ((size_t)b * parts + blockIdx.x) * kp, lane);
}

// ---- fp16 shadow of the NORMALISED rows (RAG_INDEX_F16_SHADOW): batch-1 scoring at half the bytes -------------------
// The tensor path's row operand — fp16( x / ||x|| ), built by aux_build when rows are loaded — is also the cheapest thing
// a single query can stream: 2 bytes per element instead of 4, and no ||x||^2 to accumulate (the rows are unit vectors up
// to their rounding residual, which the certification bound carries: |q.x~/||q|| - cos| <= rho_x + the fp32 accumulation).
// Same shape as the bf16 kernel: 16-byte loads, two rows in flight per warp, query in two planes. Rows whose norm is zero
// were stored as NaN and score -inf. K4 rescored the K' survivors from the fp32 rows, so reported values stay bit-exact.
__device__ __forceinline__ void f16x8_fma(const uint4 v, const float4 qa, const float4 qb, float& d0, float& d1) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
  const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&v.z)), d = __half22float2(*reinterpret_cast<const __half2*>(&v.w));
  d0 = fmaf(a.x, qa.x, d0); d1 = fmaf(a.y, qa.y, d1); d0 = fmaf(b.x, qa.z, d0); d1 = fmaf(b.y, qa.w, d1);
  d0 = fmaf(c.x, qb.x, d0); d1 = fmaf(c.y, qb.y, d1); d0 = fmaf(d.x, qb.z, d0); d1 = fmaf(d.y, qb.w, d1);
}
__device__ __forceinline__ float finish_dot(float dot) { return dot == dot ? dot : -INFINITY; }

template <int U>
__global__ void __launch_bounds__(K1_THREADS, 1)
k1_stream_f16n(const __half* __restrict__ X, uint32_t n_rows, uint32_t ld, int nsub,
               const float* __restrict__ Q, uint32_t kp, uint32_t parts, uint64_t* __restrict__ partial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* qa = reinterpret_cast<float4*>(smem_raw);   // qa[i] = q[8i..8i+3], qb[i] = q[8i+4..8i+7]
  float4* qb = qa + ld / 8;
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem_raw + (size_t)ld * sizeof(float));

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t b = blockIdx.y;
  const float4* q4 = reinterpret_cast<const float4*>(Q + (size_t)b * ld);
  for (uint32_t i = threadIdx.x; i < ld / 4; i += K1_THREADS) {
    if (i & 1) qb[i >> 1] = q4[i]; else qa[i >> 1] = q4[i];
  }
  uint64_t* mylist = lists + (size_t)warp * kp;
  for (uint32_t i = lane; i < kp; i += 32) mylist[i] = 0ull;
  __syncthreads();

  uint64_t thresh = 0ull;
  const uint32_t stride = gridDim.x * K1_WARPS;
  for (uint32_t row = blockIdx.x * K1_WARPS + warp; row < n_rows; row += 2 * stride) {
    const uint32_t row2 = row + stride;
    const bool has2 = row2 < n_rows;
    const uint4* xr0 = reinterpret_cast<const uint4*>(X + (size_t)row * ld);
    const uint4* xr1 = reinterpret_cast<const uint4*>(X + (size_t)(has2 ? row2 : row) * ld);
    float d0 = 0.f, d1 = 0.f, e0 = 0.f, e1 = 0.f;
    for (int s = 0; s < nsub; s++) {
      uint4 v[U], w[U];
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = ldg_stream_u4(xr0 + (s * U + u) * 32 + lane);
#pragma unroll
      for (int u = 0; u < U; u++) w[u] = ldg_stream_u4(xr1 + (s * U + u) * 32 + lane);
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int i = (s * U + u) * 32 + lane;
        const float4 a = qa[i], c = qb[i];
        f16x8_fma(v[u], a, c, d0, d1);
        f16x8_fma(w[u], a, c, e0, e1);
      }
    }
    const float dotA = warp_sum(d0 + d1), dotB = warp_sum(e0 + e1);
    const uint64_t keyA = rag_pack_key(finish_dot(dotA), row);
    if (keyA > thresh) warp_list_insert(mylist, kp, keyA, lane, thresh);
    if (has2) {
      const uint64_t keyB = rag_pack_key(finish_dot(dotB), row2);
      if (keyB > thresh) warp_list_insert(mylist, kp, keyB, lane, thresh);
    }
  }
  __syncthreads();
  if (warp == 0)
    warp_merge_lists(lists, K1_WARPS, kp, kp, partial + ((size_t)b * parts + blockIdx.x) * kp, lane);
}

// ---- fp32 corpus, several queries per corpus pass (small batches, escalated queries) -----------
// Same streaming pattern as k1_stream_f32, but every row a warp loads is scored against QB
// queries held in shared memory, and a warp keeps TWO rows in flight so each 128-bit query
// read from shared memory is used twice. HBM traffic stays rows*ld*4 bytes per PASS of QB
// queries (LDS traffic QB/2 x the row bytes: within the shared-memory budget up to QB = 8).
constexpr int K1M_THREADS = 512;
constexpr int K1M_WARPS = K1M_THREADS / 32;

template <int QB, int U>
__global__ void __launch_bounds__(K1M_THREADS, 1)
k1_multi_f32(const float* __restrict__ X, uint32_t n_rows, uint32_t ld, int nsub, const float* __restrict__ Q,
             uint32_t B, uint32_t kp, uint32_t parts, uint64_t* __restrict__ partial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* qs = reinterpret_cast<float4*>(smem_raw);  // [QB][ld/4]
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem_raw + (size_t)QB * ld * sizeof(float));  // [warps][QB][kp]

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t b0 = blockIdx.y * QB;
  const uint32_t ld4 = ld / 4;
  for (uint32_t i = threadIdx.x; i < QB * ld4; i += K1M_THREADS) {
    const uint32_t b = i / ld4, c = i % ld4;
    qs[i] = b0 + b < B ? reinterpret_cast<const float4*>(Q + (size_t)(b0 + b) * ld)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  uint64_t* mylists = lists + (size_t)warp * QB * kp;
  for (uint32_t i = lane; i < QB * kp; i += 32) mylists[i] = 0ull;
  __syncthreads();

  uint64_t thresh[QB];
#pragma unroll
  for (int b = 0; b < QB; b++) thresh[b] = 0ull;

  const uint32_t stride = gridDim.x * K1M_WARPS;
  for (uint32_t row = blockIdx.x * K1M_WARPS + warp; row < n_rows; row += 2 * stride) {
    const uint32_t row2 = row + stride;
    const bool has2 = row2 < n_rows;
    const float4* x0 = reinterpret_cast<const float4*>(X + (size_t)row * ld);
    const float4* x1 = reinterpret_cast<const float4*>(X + (size_t)(has2 ? row2 : row) * ld);
    float d0[QB], d1[QB], n0 = 0.f, n1 = 0.f;
#pragma unroll
    for (int b = 0; b < QB; b++) { d0[b] = 0.f; d1[b] = 0.f; }
    for (int s = 0; s < nsub; s++) {
      float4 v[U], w[U];
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = ldg_stream_f4(x0 + (s * U + u) * 32 + lane);
#pragma unroll
      for (int u = 0; u < U; u++) w[u] = ldg_stream_f4(x1 + (s * U + u) * 32 + lane);
#pragma unroll
      for (int u = 0; u < U; u++) {
        n0 = fmaf(v[u].x, v[u].x, n0); n0 = fmaf(v[u].y, v[u].y, n0); n0 = fmaf(v[u].z, v[u].z, n0); n0 = fmaf(v[u].w, v[u].w, n0);
        n1 = fmaf(w[u].x, w[u].x, n1); n1 = fmaf(w[u].y, w[u].y, n1); n1 = fmaf(w[u].z, w[u].z, n1); n1 = fmaf(w[u].w, w[u].w, n1);
#pragma unroll
        for (int b = 0; b < QB; b++) {
          const float4 qq = qs[b * ld4 + (s * U + u) * 32 + lane];
          d0[b] = fmaf(v[u].x, qq.x, d0[b]); d0[b] = fmaf(v[u].y, qq.y, d0[b]);
          d0[b] = fmaf(v[u].z, qq.z, d0[b]); d0[b] = fmaf(v[u].w, qq.w, d0[b]);
          d1[b] = fmaf(w[u].x, qq.x, d1[b]); d1[b] = fmaf(w[u].y, qq.y, d1[b]);
          d1[b] = fmaf(w[u].z, qq.z, d1[b]); d1[b] = fmaf(w[u].w, qq.w, d1[b]);
        }
      }
    }
    const float nrm0 = warp_sum(n0), nrm1 = warp_sum(n1);
#pragma unroll
    for (int b = 0; b < QB; b++) {
      const float dot0 = warp_sum(d0[b]), dot1 = warp_sum(d1[b]);
      if (b0 + b < B) {  // warp-uniform
        const uint64_t key0 = rag_pack_key(finish_score(dot0, nrm0), row);
        if (key0 > thresh[b]) warp_list_insert(mylists + (size_t)b * kp, kp, key0, lane, thresh[b]);
        if (has2) {
          const uint64_t key1 = rag_pack_key(finish_score(dot1, nrm1), row2);
          if (key1 > thresh[b]) warp_list_insert(mylists + (size_t)b * kp, kp, key1, lane, thresh[b]);
        }
      }
    }
  }
  __syncthreads();
  // one warp per query merges the 16 per-warp lists of that query
  for (uint32_t b = warp; b < QB; b += K1M_WARPS)
    if (b0 + b < B)
      warp_merge_lists(lists + (size_t)b * kp, K1M_WARPS, QB * kp, kp, partial + ((size_t)(b0 + b) * parts + blockIdx.x) * kp, lane);
}

// ---- bf16 corpus, several queries per corpus pass ------------------------------------------------
// The bf16 twin of k1_multi_f32 (escalated queries of the tensor path on a bf16 corpus used to cost one whole corpus
// pass EACH). The fp32 queries are twice as wide as the rows, so shared-memory traffic is QB x the row bytes with two
// rows in flight: QB = 4 keeps it under the shared-memory bandwidth while HBM streams rows*ld*2 bytes per pass.
template <int QB, int U>
__global__ void __launch_bounds__(K1M_THREADS, 1)
k1_multi_bf16(const __nv_bfloat16* __restrict__ X, uint32_t n_rows, uint32_t ld, int nsub, const float* __restrict__ Q,
              uint32_t B, uint32_t kp, uint32_t parts, uint64_t* __restrict__ partial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // per query two planes, so that every 128-bit LDS has a 16-byte lane stride: qa[i] = q[8i..8i+3], qb[i] = q[8i+4..8i+7]
  float4* qs = reinterpret_cast<float4*>(smem_raw);  // [QB][2][ld/8]
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem_raw + (size_t)QB * ld * sizeof(float));  // [warps][QB][kp]

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t b0 = blockIdx.y * QB;
  const uint32_t ld4 = ld / 4, ld8 = ld / 8;
  for (uint32_t i = threadIdx.x; i < QB * ld4; i += K1M_THREADS) {
    const uint32_t b = i / ld4, c = i % ld4;
    const float4 v = b0 + b < B ? reinterpret_cast<const float4*>(Q + (size_t)(b0 + b) * ld)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    qs[(size_t)b * ld4 + (c & 1) * ld8 + (c >> 1)] = v;
  }
  uint64_t* mylists = lists + (size_t)warp * QB * kp;
  for (uint32_t i = lane; i < QB * kp; i += 32) mylists[i] = 0ull;
  __syncthreads();

  uint64_t thresh[QB];
#pragma unroll
  for (int b = 0; b < QB; b++) thresh[b] = 0ull;

  const uint32_t stride = gridDim.x * K1M_WARPS;
  for (uint32_t row = blockIdx.x * K1M_WARPS + warp; row < n_rows; row += 2 * stride) {
    const uint32_t row2 = row + stride;
    const bool has2 = row2 < n_rows;
    const uint4* x0 = reinterpret_cast<const uint4*>(X + (size_t)row * ld);
    const uint4* x1 = reinterpret_cast<const uint4*>(X + (size_t)(has2 ? row2 : row) * ld);
    float d0[QB], d1[QB], n0 = 0.f, n1 = 0.f;
#pragma unroll
    for (int b = 0; b < QB; b++) { d0[b] = 0.f; d1[b] = 0.f; }
    for (int s = 0; s < nsub; s++) {
      uint4 v[U], w[U];
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = ldg_stream_u4(x0 + (s * U + u) * 32 + lane);
#pragma unroll
      for (int u = 0; u < U; u++) w[u] = ldg_stream_u4(x1 + (s * U + u) * 32 + lane);
#pragma unroll
      for (int u = 0; u < U; u++) {
        const uint32_t vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w}, ww[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
        float xv[8], xw[8];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          xv[2 * j] = __uint_as_float(vv[j] << 16); xv[2 * j + 1] = __uint_as_float(vv[j] & 0xFFFF0000u);
          xw[2 * j] = __uint_as_float(ww[j] << 16); xw[2 * j + 1] = __uint_as_float(ww[j] & 0xFFFF0000u);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) { n0 = fmaf(xv[j], xv[j], n0); n1 = fmaf(xw[j], xw[j], n1); }
        const int i = (s * U + u) * 32 + lane;
#pragma unroll
        for (int b = 0; b < QB; b++) {
          const float4 a = qs[(size_t)b * ld4 + i], c = qs[(size_t)b * ld4 + ld8 + i];
          d0[b] = fmaf(xv[0], a.x, d0[b]); d0[b] = fmaf(xv[1], a.y, d0[b]); d0[b] = fmaf(xv[2], a.z, d0[b]); d0[b] = fmaf(xv[3], a.w, d0[b]);
          d0[b] = fmaf(xv[4], c.x, d0[b]); d0[b] = fmaf(xv[5], c.y, d0[b]); d0[b] = fmaf(xv[6], c.z, d0[b]); d0[b] = fmaf(xv[7], c.w, d0[b]);
          d1[b] = fmaf(xw[0], a.x, d1[b]); d1[b] = fmaf(xw[1], a.y, d1[b]); d1[b] = fmaf(xw[2], a.z, d1[b]); d1[b] = fmaf(xw[3], a.w, d1[b]);
          d1[b] = fmaf(xw[4], c.x, d1[b]); d1[b] = fmaf(xw[5], c.y, d1[b]); d1[b] = fmaf(xw[6], c.z, d1[b]); d1[b] = fmaf(xw[7], c.w, d1[b]);
        }
      }
    }
    const float nrm0 = warp_sum(n0), nrm1 = warp_sum(n1);
#pragma unroll
    for (int b = 0; b < QB; b++) {
      const float dot0 = warp_sum(d0[b]), dot1 = warp_sum(d1[b]);
      if (b0 + b < B) {  // warp-uniform
        const uint64_t key0 = rag_pack_key(finish_score(dot0, nrm0), row);
        if (key0 > thresh[b]) warp_list_insert(mylists + (size_t)b * kp, kp, key0, lane, thresh[b]);
        if (has2) {
          const uint64_t key1 = rag_pack_key(finish_score(dot1, nrm1), row2);
          if (key1 > thresh[b]) warp_list_insert(mylists + (size_t)b * kp, kp, key1, lane, thresh[b]);
        }
      }
    }
  }
  __syncthreads();
  for (uint32_t b = warp; b < QB; b += K1M_WARPS)
    if (b0 + b < B)
      warp_merge_lists(lists + (size_t)b * kp, K1M_WARPS, QB * kp, kp, partial + ((size_t)(b0 + b) * parts + blockIdx.x) * kp, lane);
}

template <typename F>
int launch_cfg(F kernel, size_t smem) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return rag_set_error(RAG_ERR_CUDA, "K1 smem opt-in failed: %s", cudaGetErrorString(e));
  }
  return RAG_OK;
}

// largest supported unroll that divides the per-lane vector count of a row
int pick_unroll(uint32_t per_lane, const int* opts, int nopts) {
  for (int i = 0; i < nopts; i++)
    if (per_lane % opts[i] == 0) return opts[i];
  return 1;
}

}  // namespace

void k1_small_prof(int enable, FILE* dump) {
  if (enable >= 0) {
    unsigned long long z[4] = {0, 0, 0, 0};
    cudaMemcpyToSymbol(g_k1_prof_on, &enable, sizeof(int));
    cudaMemcpyToSymbol(g_k1_prof, z, sizeof(z));
  }
  if (dump) {
    unsigned long long h[4];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(h, g_k1_prof, sizeof(h));
    if (h[3])
      fprintf(dump, "[k1_stream_f32 prof] avg cycles per CTA over %llu CTAs: query staging %.0f, row loop %.0f, CTA merge + store %.0f\n",
              h[3], (double)h[0] / h[3], (double)h[1] / h[3], (double)h[2] / h[3]);
  }
}

int k1_plan(const rag_index* idx, uint32_t B, uint32_t kp, uint32_t* parts) {
  (void)kp;
  (void)B;
  *parts = (uint32_t)idx->sm_count;  // persistent: one CTA per SM (per group of queries), one candidate list each
  return RAG_OK;
}

int k1_launch(rag_index* idx, uint32_t B, uint32_t kp, uint32_t parts, bool on_shadow) {
  rag_prof_scope ps(idx, RAG_PROF_STREAM);
  if (idx->rows == 0) return rag_set_error(RAG_ERR_STATE, "search on an empty index");
  if (idx->rows >= 0xFFFFFFFFull) return rag_set_error(RAG_ERR_UNSUPPORTED, "more than 2^32-2 rows per shard");
  const uint32_t n = (uint32_t)idx->rows, ld = idx->ld;
  const size_t smem = (size_t)ld * sizeof(float) + (size_t)K1_WARPS * kp * sizeof(uint64_t);
  dim3 grid(parts, B), block(K1_THREADS);
#define K1_GO(KERNEL, TYPE)                                                                     \
  do {                                                                                          \
    RAG_CHECK(launch_cfg(KERNEL, smem));                                                        \
    KERNEL<<<grid, block, smem, idx->stream>>>((const TYPE*)idx->corpus, n, ld, nsub, idx->cur->d_q, \
                                               kp, parts, idx->cur->d_partial);                      \
  } while (0)
  // several queries per corpus pass when they fit in shared memory: 8 (or 4) queries share every row a
  // warp streams; otherwise (very wide rows, very long candidate lists) one query per pass, grid.y = B
  if (on_shadow) {
    // the fp16 shadow of the normalised rows as the streamed operand (RAG_PATH_SHADOW_STREAM): one pass per query
    if (!idx->shadow || !idx->shadow_f16 || idx->desc.dtype != RAG_F32)
      return rag_set_error(RAG_ERR_UNSUPPORTED, "shadow stream path needs an fp32 index with RAG_INDEX_F16_SHADOW");
    static const int opts[] = {6, 4, 3, 2, 1};
    const uint32_t per_lane = ld / 256;  // 16-byte loads per lane per row
    const int U = pick_unroll(per_lane, opts, 5);
    const int nsub = (int)(per_lane / U);
#define K1H_GO(Uv)                                                                                                   \
  do {                                                                                                               \
    RAG_CHECK(launch_cfg(k1_stream_f16n<Uv>, smem));                                                                 \
    k1_stream_f16n<Uv><<<grid, block, smem, idx->stream>>>(reinterpret_cast<const __half*>(idx->shadow), n, ld, nsub, \
                                                           idx->cur->d_q, kp, parts, idx->cur->d_partial);           \
  } while (0)
    switch (U) {
      case 6: K1H_GO(6); break;
      case 4: K1H_GO(4); break;
      case 3: K1H_GO(3); break;
      case 2: K1H_GO(2); break;
      default: K1H_GO(1); break;
    }
#undef K1H_GO
    RAG_CUDA(cudaGetLastError());
    idx->launches++;
    return RAG_OK;
  }
  uint32_t QB = 0;
  if (idx->desc.dtype == RAG_F32 && B > 1) {
    const size_t budget = 200 * 1024;
    auto need = [&](uint32_t qb) { return (size_t)qb * ld * sizeof(float) + (size_t)K1M_WARPS * qb * kp * sizeof(uint64_t); };
    if (B > 4 && need(8) <= budget) QB = 8;
    else if (need(4) <= budget) QB = 4;
  }
  if (QB != 0) {
    const uint32_t per_lane = ld / 128;
    const int U = per_lane % 6 == 0 ? 6 : (per_lane % 4 == 0 ? 4 : 2);
    const int nsub = (int)(per_lane / U);
    const size_t msmem = (size_t)QB * ld * sizeof(float) + (size_t)K1M_WARPS * QB * kp * sizeof(uint64_t);
    dim3 mgrid(parts, (B + QB - 1) / QB);
#define K1M_GO(QBv, Uv)                                                                                        \
  do {                                                                                                         \
    RAG_CHECK(launch_cfg(k1_multi_f32<QBv, Uv>, msmem));                                                       \
    k1_multi_f32<QBv, Uv><<<mgrid, K1M_THREADS, msmem, idx->stream>>>((const float*)idx->corpus, n, ld, nsub, \
                                                                      idx->cur->d_q, B, kp, parts, idx->cur->d_partial); \
  } while (0)
    if (QB == 8) { if (U == 6) K1M_GO(8, 6); else if (U == 4) K1M_GO(8, 4); else K1M_GO(8, 2); }
    else         { if (U == 6) K1M_GO(4, 6); else if (U == 4) K1M_GO(4, 4); else K1M_GO(4, 2); }
#undef K1M_GO
  } else if (idx->desc.dtype == RAG_BF16 && B > 1 && (size_t)4 * ld * sizeof(float) + (size_t)K1M_WARPS * 4 * kp * sizeof(uint64_t) <= 200 * 1024) {
    const uint32_t per_lane = ld / 256;  // 16-byte loads per lane per row
    const int U = per_lane % 3 == 0 ? 3 : (per_lane % 2 == 0 ? 2 : 1);
    const int nsub = (int)(per_lane / U);
    const size_t msmem = (size_t)4 * ld * sizeof(float) + (size_t)K1M_WARPS * 4 * kp * sizeof(uint64_t);
    dim3 mgrid(parts, (B + 3) / 4);
#define K1MB_GO(Uv)                                                                                          \
  do {                                                                                                        \
    RAG_CHECK(launch_cfg(k1_multi_bf16<4, Uv>, msmem));                                                       \
    k1_multi_bf16<4, Uv><<<mgrid, K1M_THREADS, msmem, idx->stream>>>((const __nv_bfloat16*)idx->corpus, n, ld, nsub, \
                                                                     idx->cur->d_q, B, kp, parts, idx->cur->d_partial); \
  } while (0)
    if (U == 3) K1MB_GO(3); else if (U == 2) K1MB_GO(2); else K1MB_GO(1);
#undef K1MB_GO
  } else if (idx->desc.dtype == RAG_F32) {
    static const int opts[] = {12, 8, 6, 4, 2};
    const uint32_t per_lane = ld / 128;  // float4 per lane per row
    const int U = pick_unroll(per_lane, opts, 5);
    const int nsub = (int)(per_lane / U);
    switch (U) {
      case 12: K1_GO(k1_stream_f32<12>, float); break;
      case 8: K1_GO(k1_stream_f32<8>, float); break;
      case 6: K1_GO(k1_stream_f32<6>, float); break;
      case 4: K1_GO(k1_stream_f32<4>, float); break;
      default: K1_GO(k1_stream_f32<2>, float); break;
    }
  } else {
    static const int opts[] = {6, 4, 3, 2, 1};
    const uint32_t per_lane = ld / 256;  // 16-byte loads per lane per row
    const int U = pick_unroll(per_lane, opts, 5);
    const int nsub = (int)(per_lane / U);
    switch (U) {
      case 6: K1_GO(k1_stream_bf16<6>, __nv_bfloat16); break;
      case 4: K1_GO(k1_stream_bf16<4>, __nv_bfloat16); break;
      case 3: K1_GO(k1_stream_bf16<3>, __nv_bfloat16); break;
      case 2: K1_GO(k1_stream_bf16<2>, __nv_bfloat16); break;
      default: K1_GO(k1_stream_bf16<1>, __nv_bfloat16); break;
    }
  }
#undef K1_GO
  RAG_CUDA(cudaGetLastError());
  idx->launches++;
  return RAG_OK;
}
