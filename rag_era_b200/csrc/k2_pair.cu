// k2_pair.cu — K2 on CTA pairs: tcgen05.mma.cta_group::2 (UMMA M=256 across two SMs).
//
// Same job as k2_tensor.cu (batched cosine scoring + fused top-K' selection; the reference loop
// it replaces is getTopKEmbeddings behind src/lib/hybrid-search.ts:223-224), re-tiled so that the
// epilogue runs UNDER the next tile's MMAs:
//
//   cluster = 2 CTAs (one TPC). The pair owns 256 queries (128 TMEM lanes in each CTA) and walks
//   corpus tiles of 256 rows. Per 64-element k-slice each CTA TMA-loads only its own 128 queries
//   (16 KB) and its own HALF of the corpus tile (128 rows, 16 KB); UMMAs 256x256x16 issued by the
//   leader reads both CTAs' shared memory and writes 128 lanes x 256 columns of fp32 into EACH
//   CTA's TMEM. A 256-column accumulator is half of TMEM, so there are two: tile t+1 accumulates
//   into one while the epilogue drains tile t from the other. L2->smem traffic per MMA cycle is
//   the same as the single-CTA kernel's (16 KB per 2 MMAs); the ring holds 3 stages of 32 KB.
//
//   Epilogue warps come in two sets of four (set = tile parity): one LANE per query, 256 scores per
//   tile each; scale by 1/||x||, threshold test, rare survivors appended to the query's 64-slot
//   buffer, warp-pruned to the K' best when it fills (see k2_tensor.cu). Each set keeps its own
//   buffers, so a (pair, query) publishes two lists: partial[B][2*pairs][K'].
//
// Barriers: full[s] lives in the leader (it counts both CTAs' TMA bytes), empty[s] and
// tmem_full[b] are signalled in both CTAs by multicast tcgen05.commit, tmem_empty[b] lives in the
// leader and collects the 8 epilogue warps of both CTAs (remote mbarrier.arrive via mapa).
//
// Roofline: tensor pipe for B >= ~64: 2*rows*ld*B flop per launch; HBM below (rows*ld*2 bytes).
#include "common.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

namespace {
using namespace tc;

constexpr int ROW_BYTES = 128;         // bytes of one operand row per stage = one whole L2 line, SWIZZLE_128B:
                                       // 64 bf16 elements, or 32 fp32 elements read as tf32
constexpr int TILE_N = 256;            // corpus rows per tile (UMMA N)
constexpr int HALF_N = TILE_N / 2;     // rows each CTA of the pair loads
constexpr int CTA_M = 128;             // queries per CTA (TMEM lanes)
constexpr int PAIR_M = 2 * CTA_M;      // UMMA M
constexpr int A_BYTES = CTA_M * ROW_BYTES;    // 16 KB
constexpr int B_BYTES = HALF_N * ROW_BYTES;   // 16 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int KP_THREADS = 384;        // w0 TMA · w1 MMA · w2 TMEM alloc · w3 inv-norm loader · w4..11 epilogue
constexpr int MAX_STAGES = 8;
constexpr int TMEM_COLS = 512;
constexpr int CAP = 64;
constexpr int SETS = 2;
constexpr int KP_PREFETCH = 8;         // k-slices of L2 prefetch ahead of the TMA loads
constexpr int KP_MAX_KP = CAP - 16;   // keep a useful append window above K'
constexpr uint32_t kIdescBf16 = idesc_bf16(PAIR_M, TILE_N);
constexpr uint32_t kIdescTf32 = idesc_tf32(PAIR_M, TILE_N);

struct kp_params {
  uint32_t n_rows, ld, B, kp, parts, stages, n_tiles, pairs;
  const float* inv_norm;
  uint64_t* partial;
  float* dbg_scores;
  uint32_t mode;
  unsigned long long* cyc;  // diagnostics (RAGERA_K2_PROF): [ctas][12 warps][8] cycle counters
};
__device__ __forceinline__ long long clk() { return clock64(); }

// see k2_tensor.cu::warp_prune
__device__ __noinline__ uint64_t warp_prune(uint64_t* buf, int cnt, int kp, int rot, int lane) {
  const uint64_t k0 = lane < cnt ? buf[(lane + rot) & (CAP - 1)] : 0ull;
  const uint64_t k1 = lane + 32 < cnt ? buf[(lane + 32 + rot) & (CAP - 1)] : 0ull;
  int r0 = 0, r1 = 0;
#pragma unroll
  for (int j = 0; j < 32; j++) {
    const uint64_t a = shfl_u64(k0, j), b = shfl_u64(k1, j);
    r0 += (a > k0 ? 1 : 0) + (b > k0 ? 1 : 0);
    r1 += (a > k1 ? 1 : 0) + (b > k1 ? 1 : 0);
  }
  __syncwarp();
  buf[lane] = 0ull;
  buf[lane + 32] = 0ull;
  __syncwarp();
  if (k0 != 0ull && r0 < kp) buf[(r0 + rot) & (CAP - 1)] = k0;
  if (k1 != 0ull && r1 < kp) buf[(r1 + rot) & (CAP - 1)] = k1;
  __syncwarp();
  return cnt >= kp ? buf[(kp - 1 + rot) & (CAP - 1)] : 0ull;
}

// Sorting-network prune for K' <= 32 (the default window): lane l holds logical entries l (lower half)
// and 32+l (upper half) of one query's buffer. Each half is bitonic-sorted across the lanes
// (descending; the lower half is skipped when it is still the sorted result of the previous prune),
// the upper half is reversed so that max(lower[l], upper[31-l]) is the bitonic sequence of the 32
// largest keys, and one bitonic merge sorts it: ~3x fewer instructions than rank-by-counting.
// On return the lower half holds the K' best in rank order (zeros after), the upper half is empty;
// the caller sets the entry count to 32 so new keys are appended to the upper half.
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
  const uint32_t hi = __shfl_xor_sync(0xFFFFFFFFu, (uint32_t)(v >> 32), m);
  const uint32_t lo = __shfl_xor_sync(0xFFFFFFFFu, (uint32_t)v, m);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t bitonic_sort32_desc(uint64_t key, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const uint64_t other = shfl_xor_u64(key, j);
      const bool take_max = ((lane & j) == 0) == ((lane & k) == 0);
      key = take_max ? max(key, other) : min(key, other);
    }
  }
  return key;
}
__device__ __noinline__ uint64_t warp_prune_sort(uint64_t* buf, int cnt, int kp, int rot, int lane, int lower_sorted) {
  uint64_t lo = lane < cnt ? buf[(lane + rot) & (CAP - 1)] : 0ull;
  uint64_t hi = lane + 32 < cnt ? buf[(lane + 32 + rot) & (CAP - 1)] : 0ull;
  if (!lower_sorted) lo = bitonic_sort32_desc(lo, lane);
  hi = bitonic_sort32_desc(hi, lane);
  const uint64_t hr = shfl_u64(hi, 31 - lane);
  uint64_t c = max(lo, hr);
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const uint64_t other = shfl_xor_u64(c, j);
    c = ((lane & j) == 0) ? max(c, other) : min(c, other);
  }
  __syncwarp();
  buf[(lane + rot) & (CAP - 1)] = lane < kp ? c : 0ull;
  buf[(lane + 32 + rot) & (CAP - 1)] = 0ull;
  __syncwarp();
  return shfl_u64(c, kp - 1);
}

// TF32 = false: bf16 operands (bf16 corpus or bf16 shadow, queries rounded to bf16), UMMA K = 16.
// TF32 = true : the fp32 corpus and the fp32 queries themselves, rounded to tf32 by TMA on the way into
//               shared memory (CU_TENSOR_MAP_DATA_TYPE_TFLOAT32), kind::tf32 UMMA K = 8 — batched scoring
//               of an fp32 index without the extra memory of a bf16 shadow (HBM streams 4 B/element).
template <bool TF32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(KP_THREADS, 1)
k2_pair_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, const kp_params P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // layout (identical in both CTAs): [stages][A | B] · lists [SETS][128][CAP] u64 · inv [2][256] f32 · barriers · tmem ptr
  unsigned char* stage_base = smem;
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem + (size_t)P.stages * STAGE_BYTES);
  float* s_inv = reinterpret_cast<float*>(lists + (size_t)SETS * CTA_M * CAP);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_inv + 2 * TILE_N);
  uint64_t* full = bars;                        // [stages]  (the leader's are used)
  uint64_t* empty = bars + MAX_STAGES;          // [stages]  (local)
  uint64_t* tmem_full = bars + 2 * MAX_STAGES;  // [2]       (local)
  uint64_t* tmem_empty = tmem_full + 2;         // [2]       (the leader's are used)
  uint64_t* inv_full = tmem_empty + 2;          // [2]       (local)
  uint64_t* inv_empty = inv_full + 2;           // [2]       (local)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(inv_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t pair = blockIdx.x >> 1;
  const uint32_t q0 = blockIdx.y * PAIR_M + rank * CTA_M;  // first query of this CTA
  constexpr int BK = TF32 ? ROW_BYTES / 4 : ROW_BYTES / 2;  // k elements per stage
  const uint32_t nkb = P.ld / BK;

  for (uint32_t i = threadIdx.x; i < SETS * CTA_M * CAP; i += KP_THREADS) lists[i] = 0ull;
  if (threadIdx.x == 0) {
    // full[s]: one arrival (the leader's expect_tx of BOTH CTAs' bytes). The peer's TMA completes on the
    // leader's barrier without an arrival of its own; its bytes can only land in the phase they belong
    // to because the peer refills a stage only after the leader's commit has released it.
    for (uint32_t s = 0; s < P.stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; b++) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 8);  // 4 epilogue warps of the set x 2 CTAs
      mbar_init(&inv_full[b], 1);
      mbar_init(&inv_empty[b], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tcgen05_fence_before();
  cluster_sync();  // barriers of BOTH CTAs are initialised before anyone signals across the pair
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own queries + own half of the corpus tile =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      long long w_empty = 0;
      const long long t_begin = clk();
      // L2 prefetch runs KP_PREFETCH k-slices ahead of the loads: a corpus slice is new to L2 for the first
      // query group that reaches it, and three stages of shared memory cannot hide an HBM miss
      uint32_t pf_tile = pair, pf_kb = 0;
      auto prefetch_next = [&]() {
        if (pf_tile < P.n_tiles) {
          tma_prefetch_2d(&map_x, (int)(pf_kb * BK), (int)(pf_tile * TILE_N + rank * HALF_N));
          if (++pf_kb == nkb) { pf_kb = 0; pf_tile += P.pairs; }
        }
      };
      for (int i = 0; i < KP_PREFETCH; i++) prefetch_next();
      for (uint32_t tile = pair; tile < P.n_tiles; tile += P.pairs) {
        for (uint32_t kb = 0; kb < nkb; kb++) {
          prefetch_next();
          const long long t0 = clk();
          mbar_wait(&empty[stage], phase ^ 1);
          w_empty += clk() - t0;
          unsigned char* sa = stage_base + (size_t)stage * STAGE_BYTES;
          const uint32_t bar = mapa(smem_u32(&full[stage]), 0);  // the leader's full barrier
          if (leader) mbar_expect_tx(&full[stage], 2 * STAGE_BYTES);
          tma_load_2d_pair(sa, &map_q, bar, (int)(kb * BK), (int)q0);
          tma_load_2d_pair(sa + A_BYTES, &map_x, bar, (int)(kb * BK), (int)(tile * TILE_N + rank * HALF_N));
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
      }
      if (P.cyc) {
        unsigned long long* c = P.cyc + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 12 + warp) * 8;
        c[0] = (unsigned long long)(clk() - t_begin); c[1] = (unsigned long long)w_empty;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (leader && lane == 0) {
      uint32_t stage = 0, phase = 0, it = 0;
      long long w_tmem = 0, w_full = 0;
      const long long t_begin = clk();
      for (uint32_t tile = pair; tile < P.n_tiles; tile += P.pairs, it++) {
        const uint32_t buf = it & 1;
        long long t0 = clk();
        mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);  // both CTAs' epilogues have drained this accumulator
        w_tmem += clk() - t0;
        tcgen05_fence_after();
        for (uint32_t kb = 0; kb < nkb; kb++) {
          t0 = clk();
          mbar_wait(&full[stage], phase);
          w_full += clk() - t0;
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(stage_base + (size_t)stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (uint32_t ks = 0; ks < ROW_BYTES / 32; ks++) {  // one UMMA consumes 32 bytes of K per row
            if (TF32)
              tcgen05_mma_tf32_pair(tmem_base + buf * TILE_N, umma_desc_sw128(sa + ks * 32), umma_desc_sw128(sb + ks * 32),
                                    kIdescTf32, (kb | ks) != 0 ? 1u : 0u);
            else
              tcgen05_mma_f16_pair(tmem_base + buf * TILE_N, umma_desc_sw128(sa + ks * 32), umma_desc_sw128(sb + ks * 32),
                                   kIdescBf16, (kb | ks) != 0 ? 1u : 0u);
          }
          tcgen05_commit_pair(&empty[stage], 3);  // both CTAs may refill this stage once the MMAs have read it
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit_pair(&tmem_full[buf], 3);  // accumulator complete in both CTAs
      }
      if (P.cyc) {
        unsigned long long* c = P.cyc + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 12 + warp) * 8;
        c[0] = (unsigned long long)(clk() - t_begin); c[1] = (unsigned long long)w_tmem; c[2] = (unsigned long long)w_full;
      }
    }
  } else if (warp == 3) {
    // ===== inverse-norm tile loader (NaN marks rows that must never be selected) =====
    uint32_t it = 0;
    for (uint32_t tile = pair; tile < P.n_tiles; tile += P.pairs, it++) {
      const uint32_t buf = it & 1;
      mbar_wait(&inv_empty[buf], ((it >> 1) & 1) ^ 1);
      const uint32_t r0 = tile * TILE_N + lane * 8;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const uint32_t r = r0 + i;
        const float x = r < P.n_rows ? __ldg(P.inv_norm + r) : 0.f;
        v[i] = x > 0.f ? x : __int_as_float(0x7FC00000);
      }
      float4* dst = reinterpret_cast<float4*>(s_inv + buf * TILE_N + lane * 8);
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
      __syncwarp();
      if (lane == 0) mbar_arrive(&inv_full[buf]);
    }
  } else if (warp >= 4) {
    // ===== epilogue: set = tile parity, one lane per query =====
    const uint32_t set = (uint32_t)(warp - 4) >> 2, quarter = warp & 3;
    const uint32_t ql = quarter * 32 + lane;  // query within the CTA
    const uint32_t qg = q0 + ql;              // query in the batch
    const bool live = qg < P.B;
    uint64_t* warp_bufs = lists + ((size_t)set * CTA_M + quarter * 32) * CAP;
    uint64_t* mybuf = warp_bufs + (size_t)lane * CAP;
    float thr = live ? -INFINITY : INFINITY;
    int cnt = 0;
    int sorted = 0;  // the lower half of this lane's buffer is the sorted result of a previous prune
    const int kp = (int)P.kp;
    const bool fast_prune = kp <= 32;
    const uint32_t taddr0 = tmem_base + ((quarter * 32u) << 16) + set * TILE_N;
    const uint32_t bar_tmem_empty = mapa(smem_u32(&tmem_empty[set]), 0);
    long long c_wait = 0, c_ld = 0, c_sel = 0, c_prune = 0, n_prune = 0, n_app = 0;
    const long long t_begin = clk();

    // trim the buffer of lane `src` to its K' best and raise that lane's threshold (whole warp helps)
    auto prune_lane = [&](int src) {
      n_prune++;
      const int c = __shfl_sync(0xFFFFFFFFu, cnt, src);
      if (fast_prune) {
        const int was_sorted = __shfl_sync(0xFFFFFFFFu, sorted, src);
        const uint64_t t = warp_prune_sort(warp_bufs + (size_t)src * CAP, c, kp, src, lane, was_sorted);
        if (lane == src) {
          cnt = 32;  // K' best in the lower half (zero padded), appends continue in the upper half
          sorted = 1;
          if (t != 0ull) thr = rag_key_score(t);
        }
      } else {
        const uint64_t t = warp_prune(warp_bufs + (size_t)src * CAP, c, kp, src, lane);
        if (lane == src) {
          cnt = min(c, kp);
          if (t != 0ull) thr = rag_key_score(t);
        }
      }
    };

    auto process16 = [&](const uint32_t (&v)[16], const float* inv, uint32_t row) {
      float s[16];
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 w = *reinterpret_cast<const float4*>(inv + i);
        s[i] = __uint_as_float(v[i]) * w.x;
        s[i + 1] = __uint_as_float(v[i + 1]) * w.y;
        s[i + 2] = __uint_as_float(v[i + 2]) * w.z;
        s[i + 3] = __uint_as_float(v[i + 3]) * w.w;
      }
      if (P.dbg_scores && live) {
#pragma unroll
        for (int i = 0; i < 16; i++)
          if (row + i < P.n_rows) P.dbg_scores[(size_t)qg * P.n_rows + row + i] = s[i];
      }
      if (P.mode != 0) return;
      const long long ts = clk();
      // rows arrive in increasing order within a set, so a later equal score can never displace an
      // earlier one: the strict float compare against the K'-th best is exact. NaN never passes.
      const long long tp = clk();
      unsigned pm = 0;
#pragma unroll
      for (int i = 0; i < 16; i++) pm |= s[i] > thr ? (1u << i) : 0u;
      if (__any_sync(0xFFFFFFFFu, pm != 0u)) {
        // straight-line predicated appends; the slot of hit i is cnt + (hits below i), so the sixteen
        // stores are independent of each other (no serial dependence through cnt)
        // (inline PTX keeps the sixteen stores predicated instead of sixteen divergent branches)
        const uint32_t buf_addr = smem_u32(mybuf);
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const uint64_t key = rag_pack_key(s[i], row + i);
          const uint32_t slot = (uint32_t)(cnt + __popc(pm & ((1u << i) - 1u)) + lane) & (CAP - 1);
          asm volatile(
              "{\n\t.reg .pred q;\n\t"
              "setp.ne.b32 q, %0, 0;\n\t"
              "@q st.shared.b64 [%1], %2;\n\t}"
              ::"r"(pm & (1u << i)), "r"(buf_addr + slot * 8u), "l"(key)
              : "memory");
        }
        cnt += __popc(pm);
        __syncwarp();  // the appends above are visible to the lanes that help prune
        unsigned need = __ballot_sync(0xFFFFFFFFu, cnt > CAP - 16);  // must make room now
        while (need) {
          const int src = __ffs(need) - 1;
          need &= need - 1;
          prune_lane(src);
        }
      }
      c_prune += clk() - tp;
      c_sel += clk() - ts;
    };

    uint32_t it = 0;
    for (uint32_t tile = pair; tile < P.n_tiles; tile += P.pairs, it++) {
      if ((it & 1) != set) continue;
      const uint32_t use = it >> 1;  // how many times this set's buffers have been used before
      const long long tw = clk();
      mbar_wait(&inv_full[set], use & 1);
      mbar_wait(&tmem_full[set], use & 1);
      c_wait += clk() - tw;
      tcgen05_fence_after();
      const float* inv = s_inv + set * TILE_N;
      const uint32_t row0 = tile * TILE_N;
      if (P.mode != 2) {
#pragma unroll 1
        for (uint32_t c0 = 0; c0 < TILE_N; c0 += 32) {
          uint32_t va[16], vb[16];
          const long long tl = clk();
          tmem_ld16(taddr0 + c0, va);
          tmem_ld16(taddr0 + c0 + 16, vb);
          tmem_ld_wait();
          c_ld += clk() - tl;
          process16(va, inv + c0, row0 + c0);
          process16(vb, inv + c0 + 16, row0 + c0 + 16);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(bar_tmem_empty);  // the leader's barrier (also from the leader itself)
        mbar_arrive(&inv_empty[set]);
      }
      // Off the critical path (the accumulator is already released): trim buffers that are getting
      // full, so that the next tile rarely has to stop and prune while it holds TMEM.
      if (P.mode == 0) {
        const long long tq = clk();
        unsigned need = __ballot_sync(0xFFFFFFFFu, cnt > CAP - 24);
        while (need) {
          const int src = __ffs(need) - 1;
          need &= need - 1;
          prune_lane(src);
        }
        c_prune += clk() - tq;
      }
    }
    if (P.cyc && lane == 0) {
      unsigned long long* c = P.cyc + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 12 + warp) * 8;
      c[0] = (unsigned long long)(clk() - t_begin); c[1] = (unsigned long long)c_wait; c[2] = (unsigned long long)c_ld;
      c[3] = (unsigned long long)c_sel; c[4] = (unsigned long long)c_prune; c[5] = (unsigned long long)n_app;
      c[6] = (unsigned long long)n_prune;
    }
    // final prune of every query of this warp, then publish the K' best (sorted) of this (pair, set)
    for (int src = 0; src < 32; src++) {
      const int c = __shfl_sync(0xFFFFFFFFu, cnt, src);
      warp_prune(warp_bufs + (size_t)src * CAP, c, kp, src, lane);
    }
    if (live) {
      uint64_t* out = P.partial + ((size_t)qg * P.parts + pair * SETS + set) * P.kp;
      for (int j = 0; j < kp; j++) out[j] = mybuf[(j + lane) & (CAP - 1)];
    }
  }

  // neither CTA may leave (or free TMEM) while its peer can still touch its shared memory / barriers
  tcgen05_fence_before();
  cluster_sync();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// ---- host side ------------------------------------------------------------------------------
struct kp_state {
  PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  int max_smem = 0;
  bool attr_set = false;
  float* dbg = nullptr;
  uint32_t mode = 0;
  bool prof = false;
  unsigned long long* d_cyc = nullptr;
};

size_t kp_smem_bytes(uint32_t stages) {
  return (size_t)stages * STAGE_BYTES + (size_t)SETS * CTA_M * CAP * 8 + 2 * TILE_N * 4 + (2 * MAX_STAGES + 8) * 8 + 16;
}

uint32_t kp_pick_stages(const kp_state* st) {
  uint32_t s = MAX_STAGES;
  while (s > 0 && kp_smem_bytes(s) > (size_t)st->max_smem) s--;
  return s;
}

int kp_init(rag_index* idx) {
  if (idx->k2p_state) return RAG_OK;
  kp_state* st = new kp_state();
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
    delete st;
    cudaGetLastError();
    return rag_set_error(RAG_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  }
  st->encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  if (const char* m = getenv("RAGERA_K2_MODE")) st->mode = (uint32_t)atoi(m);
  if (const char* m = getenv("RAGERA_K2_PROF")) st->prof = atoi(m) != 0;
  cudaDeviceGetAttribute(&st->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, idx->device);
  idx->k2p_state = st;
  return RAG_OK;
}

int kp_make_map(kp_state* st, CUtensorMap* m, const void* base, uint64_t rows, uint32_t ld, uint32_t box_rows, bool tf32) {
  const uint32_t es = tf32 ? 4 : 2;
  cuuint64_t dims[2] = {ld, rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * es};
  cuuint32_t box[2] = {ROW_BYTES / es, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = st->encode(m, tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return rag_set_error(RAG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return RAG_OK;
}

}  // namespace

int k2p_plan(rag_index* idx, uint32_t B, uint32_t kp, uint32_t* parts) {
  RAG_CHECK(kp_init(idx));
  kp_state* st = (kp_state*)idx->k2p_state;
  if (kp > KP_MAX_KP) return rag_set_error(RAG_ERR_UNSUPPORTED, "tensor path keeps at most %d candidates per query (K'=%u)", KP_MAX_KP, kp);
  if (kp_pick_stages(st) < 3) return rag_set_error(RAG_ERR_UNSUPPORTED, "tensor path: not enough shared memory");
  const uint32_t groups = (B + PAIR_M - 1) / PAIR_M;
  const uint32_t all_pairs = (uint32_t)idx->sm_count / 2;
  if (groups > all_pairs)
    return rag_set_error(RAG_ERR_UNSUPPORTED, "tensor path: batch %u exceeds %u queries per launch", B, all_pairs * PAIR_M);
  const uint64_t n_tiles = (idx->rows + TILE_N - 1) / TILE_N;
  uint64_t pairs = all_pairs / groups;
  if (pairs > n_tiles) pairs = n_tiles;
  if (pairs < 1) pairs = 1;
  *parts = (uint32_t)pairs * SETS;
  return RAG_OK;
}

int k2p_launch(rag_index* idx, uint32_t B, uint32_t kp, uint32_t parts) {
  if (idx->rows == 0) return rag_set_error(RAG_ERR_STATE, "search on an empty index");
  if (idx->rows >= 0xFFFFFF00ull) return rag_set_error(RAG_ERR_UNSUPPORTED, "more than 2^32-257 rows per shard");
  RAG_CHECK(kp_init(idx));
  kp_state* st = (kp_state*)idx->k2p_state;
  rag_batch* bt = idx->cur;
  const bool tf32 = idx->shadow == nullptr;  // fp32 index without a bf16 shadow: score the fp32 rows as tf32
  if (tf32 && idx->desc.dtype != RAG_F32) return rag_set_error(RAG_ERR_STATE, "tensor path: no bf16 operand");
  CUtensorMap map_q, map_x;
  if (!tf32) {
    const uint32_t Bpad = (B + CTA_M - 1) / CTA_M * CTA_M;
    const size_t need = (size_t)Bpad * idx->ld * 2;
    if (need > bt->c_qb || !bt->d_qb) {
      if (bt->d_qb) RAG_CUDA(cudaFree(bt->d_qb));
      bt->d_qb = nullptr;
      bt->c_qb = 0;
      RAG_CUDA(cudaMalloc((void**)&bt->d_qb, need));
      bt->c_qb = need;
    }
    RAG_CHECK(q_to_bf16_launch(idx, B, Bpad));
    RAG_CHECK(kp_make_map(st, &map_q, bt->d_qb, Bpad, idx->ld, CTA_M, false));
    RAG_CHECK(kp_make_map(st, &map_x, idx->shadow, idx->rows, idx->ld, HALF_N, false));
  } else {
    // the fp32 queries as staged ([B][ld], zero padded columns); rows past B read as zeros (TMA OOB fill)
    RAG_CHECK(kp_make_map(st, &map_q, bt->d_q, B, idx->ld, CTA_M, true));
    RAG_CHECK(kp_make_map(st, &map_x, idx->corpus, idx->rows, idx->ld, HALF_N, true));
  }

  rag_prof_scope ps(idx, RAG_PROF_TENSOR);
  kp_params P;
  P.n_rows = (uint32_t)idx->rows;
  P.ld = idx->ld;
  P.B = B;
  P.kp = kp;
  P.parts = parts;
  P.pairs = parts / SETS;
  P.stages = kp_pick_stages(st);
  P.n_tiles = (uint32_t)((idx->rows + TILE_N - 1) / TILE_N);
  P.inv_norm = idx->inv_norm;
  P.partial = bt->d_partial;
  P.dbg_scores = st->dbg;
  P.mode = st->mode;
  P.cyc = nullptr;
  const size_t n_cyc = (size_t)P.pairs * 2 * ((B + PAIR_M - 1) / PAIR_M) * 12 * 8;
  if (st->prof) {
    if (st->d_cyc) cudaFree(st->d_cyc);
    RAG_CUDA(cudaMalloc((void**)&st->d_cyc, n_cyc * 8));
    RAG_CUDA(cudaMemsetAsync(st->d_cyc, 0, n_cyc * 8, idx->stream));
    P.cyc = st->d_cyc;
  }
  const size_t smem = kp_smem_bytes(P.stages);
  if (!st->attr_set) {
    RAG_CUDA(cudaFuncSetAttribute(k2_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, st->max_smem));
    RAG_CUDA(cudaFuncSetAttribute(k2_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, st->max_smem));
    st->attr_set = true;
  }
  const uint32_t groups = (B + PAIR_M - 1) / PAIR_M;
  if (tf32) k2_pair_kernel<true><<<dim3(P.pairs * 2, groups), KP_THREADS, smem, idx->stream>>>(map_q, map_x, P);
  else k2_pair_kernel<false><<<dim3(P.pairs * 2, groups), KP_THREADS, smem, idx->stream>>>(map_q, map_x, P);
  RAG_CUDA(cudaGetLastError());
  idx->launches++;
  if (st->prof) {
    static int printed = 0;
    std::vector<unsigned long long> h(n_cyc);
    RAG_CUDA(cudaMemcpyAsync(h.data(), st->d_cyc, n_cyc * 8, cudaMemcpyDeviceToHost, idx->stream));
    RAG_CUDA(cudaStreamSynchronize(idx->stream));
    if (printed++ % 16 == 8) {
      const size_t nct = n_cyc / 96;
      double acc[2][12][8] = {{{0}}};
      for (size_t c = 0; c < nct; c++)
        for (int w = 0; w < 12; w++)
          for (int j = 0; j < 8; j++) acc[c & 1][w][j] += (double)h[(c * 12 + w) * 8 + j] / (nct / 2);
      fprintf(stderr, "[k2 pair prof] B=%u rows=%u kp=%u stages=%u ctas=%zu (avg cycles per CTA; rank0 | rank1)\n", B, P.n_rows, kp, P.stages, nct);
      fprintf(stderr, "  producer: total %.0f wait_empty %.0f | total %.0f wait_empty %.0f\n", acc[0][0][0], acc[0][0][1], acc[1][0][0], acc[1][0][1]);
      fprintf(stderr, "  mma     : total %.0f wait_tmem_empty %.0f wait_full %.0f\n", acc[0][1][0], acc[0][1][1], acc[0][1][2]);
      for (int w = 4; w < 12; w++)
        fprintf(stderr, "  epi w%-2d : total %.0f wait %.0f ld %.0f select %.0f (prune %.0f) appends(lane0) %.0f prunes %.0f | wait %.0f ld %.0f select %.0f\n", w,
                acc[0][w][0], acc[0][w][1], acc[0][w][2], acc[0][w][3], acc[0][w][4], acc[0][w][5], acc[0][w][6], acc[1][w][1], acc[1][w][2], acc[1][w][3]);
    }
  }
  return RAG_OK;
}

void k2p_set_debug(rag_index* idx, float* d_scores) {
  if (kp_init(idx) == RAG_OK) ((kp_state*)idx->k2p_state)->dbg = d_scores;
}

void k2p_destroy(rag_index* idx) {
  if (idx->k2p_state && ((kp_state*)idx->k2p_state)->d_cyc) cudaFree(((kp_state*)idx->k2p_state)->d_cyc);
  delete (kp_state*)idx->k2p_state;
  idx->k2p_state = nullptr;
}
