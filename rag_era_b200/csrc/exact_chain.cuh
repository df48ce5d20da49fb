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
// The query chunk is widened to binary64 ONCE per warp (each lane converts its share into a
// shared buffer that all lanes then read by broadcast) instead of once per lane, and ||q||^2 —
// the same chain for every row — is left to the caller (query_norm_sq, once per query):
// ncu showed the fp32->fp64 conversions (XU pipe) and the fp64 pipe as the busiest units.
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
  double dot, nx, nq;  // nq is filled in by the caller (query_norm_sq)
};
__device__ __forceinline__ void step(chains& c, double qd, float x) {
  const double xd = (double)x;
  c.dot = __dadd_rn(c.dot, __dmul_rn(qd, xd));
  c.nx = __dadd_rn(c.nx, __dmul_rn(xd, xd));
}

// sum of q[i]*q[i], left to right (the reference's third chain; identical for every row). Whole warp: the
// lanes form 32 exact products at a time, then every lane replays the adds in order, fetching product i
// from lane i with a shuffle — only the dependent fp64 adds are on the critical path. ld % 32 == 0.
__device__ __forceinline__ double query_norm_sq(const float* __restrict__ q, uint32_t ld, int lane) {
  double nq = 0.0;
  for (uint32_t c = 0; c < ld; c += 32) {
    const double qd = (double)__ldg(q + c + lane);
    const long long p = __double_as_longlong(__dmul_rn(qd, qd));
#pragma unroll
    for (int l = 0; l < 32; l++) nq = __dadd_rn(nq, __longlong_as_double(__shfl_sync(0xFFFFFFFFu, p, l)));
  }
  return nq;
}

constexpr int X_BYTES = 32 * 512;  // 32 rows x 512 B per stage
constexpr int Q_BYTES = 1024;      // up to 256 fp32 query elements per stage
constexpr int STAGE_BYTES = X_BYTES + Q_BYTES;
constexpr int QD_BYTES = 256 * 8;  // the current query chunk as binary64
constexpr int WARP_BYTES = 2 * STAGE_BYTES + QD_BYTES;

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
    // widen this chunk of the query once per warp: lane l converts elements l, l+32, ...
    const float* qf = reinterpret_cast<const float*>(xs + X_BYTES);
    double* qd = reinterpret_cast<double*>(wsm + 2 * STAGE_BYTES);
#pragma unroll
    for (int i = 0; i < ELEMS / 32; i++) qd[lane + 32 * i] = (double)qf[lane + 32 * i];
    __syncwarp();
    const double2* qv = reinterpret_cast<const double2*>(qd);  // every lane reads the same address: broadcast
    if (!BF16) {
      const float4* xrow = reinterpret_cast<const float4*>(xs + lane * 512);
#pragma unroll 4
      for (int s = 0; s < 32; s++) {
        const float4 x = xrow[s ^ lane];
        const double2 qa = qv[2 * s], qb = qv[2 * s + 1];
        step(c, qa.x, x.x); step(c, qa.y, x.y); step(c, qb.x, x.z); step(c, qb.y, x.w);
      }
    } else {
      const uint4* xrow = reinterpret_cast<const uint4*>(xs + lane * 512);
#pragma unroll 2
      for (int s = 0; s < 32; s++) {
        const uint4 x = xrow[s ^ lane];
        const double2 q0 = qv[4 * s], q1 = qv[4 * s + 1], q2 = qv[4 * s + 2], q3 = qv[4 * s + 3];
        step(c, q0.x, __uint_as_float(x.x << 16)); step(c, q0.y, __uint_as_float(x.x & 0xFFFF0000u));
        step(c, q1.x, __uint_as_float(x.y << 16)); step(c, q1.y, __uint_as_float(x.y & 0xFFFF0000u));
        step(c, q2.x, __uint_as_float(x.z << 16)); step(c, q2.y, __uint_as_float(x.z & 0xFFFF0000u));
        step(c, q3.x, __uint_as_float(x.w << 16)); step(c, q3.y, __uint_as_float(x.w & 0xFFFF0000u));
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
