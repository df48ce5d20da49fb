// exact_chain.cuh — the reference's cosine arithmetic, bit for bit, for 32 rows per warp.
//
// similarity(q, x, cosine) of llamaindex (upstream-recalled; reached from
// src/lib/hybrid-search.ts:223-224) is three left-to-right binary64 sums
//   dot += q[i]*x[i];  nx += x[i]*x[i];  nq += q[i]*q[i]
// followed by dot / (sqrt(nq) * sqrt(nx)). JavaScript never contracts a*b+c, so every
// operation here is an explicit round-to-nearest intrinsic (no FMA).
//
// Mapping: one LANE per row. The 32 rows of a warp are staged through shared memory with
// cp.async in 512-byte column chunks (double buffered); 16-byte slot c of row j is stored
// at slot (c ^ j) so the per-lane 128-bit reads (lane j reads row j) are bank-conflict
// free. The sums are latency-bound dependent chains — 32 rows advance in lock step.
#pragma once
#include "common.cuh"

namespace rag_exact {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct chains {
  double dot, nx, nq;
};
__device__ __forceinline__ void step(chains& c, float q, float x) {
  const double qd = (double)q, xd = (double)x;
  c.dot = __dadd_rn(c.dot, __dmul_rn(qd, xd));
  c.nx = __dadd_rn(c.nx, __dmul_rn(xd, xd));
  c.nq = __dadd_rn(c.nq, __dmul_rn(qd, qd));
}

constexpr int X_BYTES = 32 * 512;  // 32 rows x 512 B per stage
constexpr int Q_BYTES = 1024;      // up to 256 fp32 query elements per stage
constexpr int STAGE_BYTES = X_BYTES + Q_BYTES;
constexpr int WARP_BYTES = 2 * STAGE_BYTES;

// Exact sums for row `row` (one per lane; every lane must pass a readable row) against q[0..ld).
// wsm: this warp's WARP_BYTES of shared memory. All 32 lanes must call.
template <bool BF16>
__device__ __forceinline__ chains warp_exact_sums(const void* __restrict__ X, uint32_t ld,
                                                  const float* __restrict__ q, uint32_t row,
                                                  unsigned char* wsm, int lane) {
  constexpr int ELEMS = BF16 ? 256 : 128;  // columns per 512-byte chunk
  const int nch = (int)(ld / ELEMS);
  const size_t row_bytes = (size_t)ld * (BF16 ? 2 : 4);

  auto issue = [&](int ch, int stage) {
    unsigned char* xs = wsm + stage * STAGE_BYTES;
    unsigned char* qs = xs + X_BYTES;
#pragma unroll 8
    for (int r = 0; r < 32; r++) {
      const uint32_t rr = __shfl_sync(0xFFFFFFFFu, row, r);
      const unsigned char* src = (const unsigned char*)X + (size_t)rr * row_bytes + (size_t)ch * 512 + lane * 16;
      cp_async16(xs + r * 512 + ((lane ^ r) * 16), src);
    }
    const unsigned char* qsrc = (const unsigned char*)(q + (size_t)ch * ELEMS);
    cp_async16(qs + lane * 16, qsrc + lane * 16);
    if (BF16) cp_async16(qs + 512 + lane * 16, qsrc + 512 + lane * 16);
  };

  chains c = {0.0, 0.0, 0.0};
  issue(0, 0);
  cp_async_commit();
  for (int ch = 0; ch < nch; ch++) {
    if (ch + 1 < nch) issue(ch + 1, (ch + 1) & 1);
    cp_async_commit();
    cp_async_wait1();
    __syncwarp();
    const unsigned char* xs = wsm + (ch & 1) * STAGE_BYTES;
    const float4* qv = reinterpret_cast<const float4*>(xs + X_BYTES);
    if (!BF16) {
      const float4* xrow = reinterpret_cast<const float4*>(xs + lane * 512);
#pragma unroll 4
      for (int s = 0; s < 32; s++) {
        const float4 x = xrow[s ^ lane];
        const float4 qq = qv[s];
        step(c, qq.x, x.x); step(c, qq.y, x.y); step(c, qq.z, x.z); step(c, qq.w, x.w);
      }
    } else {
      const uint4* xrow = reinterpret_cast<const uint4*>(xs + lane * 512);
#pragma unroll 2
      for (int s = 0; s < 32; s++) {
        const uint4 x = xrow[s ^ lane];
        const float4 qa = qv[2 * s], qb = qv[2 * s + 1];
        step(c, qa.x, __uint_as_float(x.x << 16)); step(c, qa.y, __uint_as_float(x.x & 0xFFFF0000u));
        step(c, qa.z, __uint_as_float(x.y << 16)); step(c, qa.w, __uint_as_float(x.y & 0xFFFF0000u));
        step(c, qb.x, __uint_as_float(x.z << 16)); step(c, qb.y, __uint_as_float(x.z & 0xFFFF0000u));
        step(c, qb.z, __uint_as_float(x.w << 16)); step(c, qb.w, __uint_as_float(x.w & 0xFFFF0000u));
      }
    }
    __syncwarp();
  }
  cp_async_wait0();
  return c;
}

// similarity = dot / (norm(q) * norm(x))
__device__ __forceinline__ double finish(const chains& c) {
  return __ddiv_rn(c.dot, __dmul_rn(__dsqrt_rn(c.nq), __dsqrt_rn(c.nx)));
}

}  // namespace rag_exact
