// k2_pair.cu — K2 on CTA pairs: tcgen05.mma.cta_group::2 (UMMA M=256 across two SMs).
//
// Batched cosine scoring + fused top-K' selection (the reference loop it replaces is getTopKEmbeddings
// behind src/lib/hybrid-search.ts:223-224), tiled so that the epilogue runs UNDER the next tile's MMAs:
//
//   cluster = 2 CTAs (one TPC). The pair owns 256 queries (128 TMEM lanes in each CTA) and walks
//   corpus tiles of 256 rows. Per 64-element k-slice each CTA TMA-loads only its own 128 queries
//   (16 KB) and its own HALF of the corpus tile (128 rows, 16 KB); UMMAs 256x256x16 issued by the
//   leader read both CTAs' shared memory and write 128 lanes x 256 columns of fp32 into EACH
//   CTA's TMEM. A 256-column accumulator is half of TMEM, so there are two: tile t+1 accumulates
//   into one while the epilogue drains tile t from the other. The ring holds 5 stages of 32 KB
//   (4 for K' > 32); 5, 6 and 7 stages measure the same, 3 left the MMA waiting for TMA.
//
//   Epilogue = 4 warps, one LANE per query, 256 scores per tile each (the score matrix is never written):
//   * scan: scale 32 scores by 1/||x||, take their maximum, one vote against the per-query threshold
//     (~2 instructions per score when nothing survives); otherwise 32 branch-free predicated appends
//     to the query's 32-slot WINDOW in shared memory.
//   * fold: a full window is merged into the query's sorted K' best (shared memory, XOR-swizzled by the
//     lane like the window, so warp-cooperative and lane-local accesses are both conflict free).
//     One query at a time by RANK COUNTING (k2p_fold_one: broadcast loads + a binary search by indexed
//     shuffles — the warp is alone on its scheduler, so a dependent shuffle chain is pure latency),
//     or, when several windows are full at once (the first tiles), LANE-LOCALLY (k2p_fold_local: 32
//     register sorting networks side by side).
//   * first tile: no threshold yet. Instead of appending all 256 scores every lane finds its exact K'-th
//     largest score of the tile with a register sorting network and starts from that threshold
//     (k2p_first_tile_threshold); the start-up transient used to stall the tensor pipe ~10% of the kernel.
//   * cooperative threshold: the gridDim.x/2 pairs that share a query publish their ceil(K'/pairs)-th best
//     score; at least K' rows score >= the minimum of the published values, so nothing below it can be a
//     candidate anywhere. Cuts survivors and folds ~3x; the merged top-K' does not depend on timing.
//   * lockstep of the query groups: the gridDim.y pairs with the same pair index read the same corpus
//     tiles; a pair runs at most 1 tile ahead of the slowest group, so the tile is still in L2 for the
//     others (without it K2 is 18% slower at 50M rows: every group streams the corpus from HBM itself,
//     and under the power cap the extra DRAM traffic also costs SM clock; leads of 0/1/2/4/8 tiles measure
//     137.6/135.7/138.9/147.2/150.8 ms there).
//
// Barriers: full[s] lives in the leader (it counts both CTAs' TMA bytes), empty[s] and
// tmem_full[b] are signalled in both CTAs by multicast tcgen05.commit, tmem_empty[b] lives in the
// leader and collects the 4 epilogue warps of both CTAs (remote mbarrier.arrive via mapa).
//
// Roofline: tensor pipe for B >= ~64: 2*rows*ld*B flop per launch; HBM below (rows*ld*2 bytes).
#include "common.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <math_constants.h>
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
constexpr int KP_THREADS = 256;        // w0 TMA · w1 MMA · w2 TMEM alloc · w3 inv-norm loader · w4..7 epilogue
constexpr int KP_WARPS = KP_THREADS / 32;
constexpr int MAX_STAGES = 7;         // barrier slots; the product runs what fits beside the windows and best lists (5)
constexpr int TMEM_COLS = 512;
constexpr int WIN = 32;                // append window per query (shared memory), entries
constexpr int KP_PREFETCH = 8;         // k-slices of L2 prefetch ahead of the TMA loads
constexpr int KP_MAX_KP = 64;          // two ranks per lane
constexpr uint32_t kIdescTf32 = idesc_tf32(PAIR_M, TILE_N);

struct kp_params {
  uint32_t n_rows, ld, B, kp, parts, stages, n_tiles, pairs;
  uint32_t idesc;      // kind::f16 instruction descriptor: fp16 x fp16 (fp16 shadow) or bf16 x bf16
  const float* inv_norm;
  float floor;             // a row must score ABOVE it to be a candidate (-inf: none; see rag_eps::floor)
  uint64_t* partial;
  uint32_t* pub;       // [pairs][Bpub] ordered score of each (pair, query)'s pub_rank-th best so far (0 = none yet)
  uint32_t Bpub, pub_rank, pub_every;
  uint32_t prefetch;   // k-slices of L2 prefetch ahead of the TMA loads
  uint32_t local_min;  // fold lane-locally (all 32 queries at once) when at least this many windows are full
  uint32_t* prog;      // [groups][pairs] tiles started by each CTA pair (lockstep of the query groups)
  uint32_t lockstep;   // a pair starts tile i only when every group's same-index pair has started tile i - lockstep (0 = off)
  float* dbg_scores;
  uint32_t mode;
  uint32_t nolists;
  unsigned long long* cyc;  // diagnostics (RAGERA_K2_PROF): [ctas][8 warps][8] cycle counters
};
__device__ __forceinline__ long long clk() { return clock64(); }

// ---- lane-local fold: every lane folds ITS OWN query's window into its own best list ---------------
// Used when several queries of the warp need folding at once (the first tiles of a CTA, before the
// thresholds have risen): 32 register sorting networks run side by side in the 32 lanes, instead of one
// warp-wide network per query. Window and best list are XOR-swizzled by the lane (entry j of lane l sits
// at slot j ^ l), so lane-local and warp-cooperative accesses are both bank-conflict free.
__device__ __forceinline__ void ce64_desc(uint64_t& a, uint64_t& b) {  // a >= b afterwards
  const uint64_t hi = max(a, b), lo = min(a, b);
  a = hi;
  b = lo;
}
__device__ __forceinline__ void reg_sort32_desc_u64(uint64_t (&v)[32]) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < 32; i++) {
        const int l = i ^ j;
        if (l > i) {
          if ((i & k) == 0) ce64_desc(v[i], v[l]);
          else ce64_desc(v[l], v[i]);
        }
      }
    }
  }
}
__device__ __forceinline__ void reg_merge32_desc_u64(uint64_t (&v)[32]) {  // bitonic -> sorted descending
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
#pragma unroll
    for (int i = 0; i < 32; i++) {
      const int l = i ^ j;
      if (l > i) ce64_desc(v[i], v[l]);
    }
  }
}
// mywin/mybest: this lane's window (raw entries) and best list; c, nb: their entry counts. Returns the K'-th
// best key afterwards (0 while fewer than K' exist). Half-cleaner identities: with B and W sorted descending,
// max(B[i], W[31-i]) is a bitonic sequence of the 32 largest of B u W, min(...) of the other 32.
__device__ __noinline__ uint64_t k2p_fold_local(const uint64_t* mywin, uint64_t* mybest, int c, int nb, int kp, int lane) {
  uint64_t W[32], B[32];
#pragma unroll
  for (int i = 0; i < 32; i++) {
    const uint64_t raw = mywin[i ^ lane];
    W[i] = i < c ? rag_pack_key(__uint_as_float((uint32_t)(raw >> 32)), (uint32_t)raw) : 0ull;
  }
  reg_sort32_desc_u64(W);
#pragma unroll
  for (int i = 0; i < 32; i++) B[i] = i < nb ? mybest[i ^ lane] : 0ull;
#pragma unroll
  for (int i = 0; i < 16; i++) {  // half-cleaner: B <- the 32 largest, W <- the rest (both bitonic)
    const uint64_t wa = W[i], wb = W[31 - i];
    W[i] = min(B[i], wb);
    B[i] = max(B[i], wb);
    W[31 - i] = min(B[31 - i], wa);
    B[31 - i] = max(B[31 - i], wa);
  }
  reg_merge32_desc_u64(B);
#pragma unroll
  for (int i = 0; i < 32; i++) mybest[i ^ lane] = B[i];
  if (kp <= 32) {
    uint64_t kth = B[31];
#pragma unroll
    for (int i = 0; i < 31; i++)
      if (i == kp - 1) kth = B[i];
    return kth;
  }
  // ranks 32..63: the 32 largest of (the displaced half W) u (the old ranks 32..63)
  reg_merge32_desc_u64(W);
#pragma unroll
  for (int i = 0; i < 32; i++) {
    const uint64_t b1 = 32 + (31 - i) < nb ? mybest[32 + ((31 - i) ^ lane)] : 0ull;
    W[i] = max(W[i], b1);
  }
  reg_merge32_desc_u64(W);
#pragma unroll
  for (int i = 0; i < 32; i++) mybest[32 + (i ^ lane)] = W[i];
  uint64_t kth = W[31];
#pragma unroll
  for (int i = 0; i < 31; i++)
    if (i == kp - 33) kth = W[i];
  return kth;
}

// Fold ONE query's window into its best list by rank counting — the steady-state case (windows fill one at
// a time once the thresholds have risen). Every key's final rank = how many keys of the union beat it: the
// window keys are broadcast from shared memory (independent loads and compares, no dependent shuffle chain —
// this warp is alone on its scheduler, so latency is what costs), the sorted best list is binary-searched
// with indexed shuffles. Keys are distinct (they contain the row), so the ranks are a permutation; a key
// whose rank is < K' is stored at best[rank]. Same result as k2p_fold_local for that lane.
__device__ __noinline__ uint64_t k2p_fold_one(uint64_t* wsrc, uint64_t* bsrc, int c, int nb, int kp, int src, int lane) {
  const uint64_t b0 = lane < nb ? bsrc[lane ^ src] : 0ull;  // rank r of query src sits at slot (r & 32) | ((r ^ src) & 31)
  const uint64_t b1 = lane + 32 < nb ? bsrc[32 + (lane ^ src)] : 0ull;
  uint64_t w = 0ull;
  if (lane < c) {
    const uint64_t raw = wsrc[lane ^ src];
    w = rag_pack_key(__uint_as_float((uint32_t)(raw >> 32)), (uint32_t)raw);
    wsrc[lane ^ src] = w;
  }
  __syncwarp();
  int rw = 0, r0 = 0, r1 = 0;
#pragma unroll 4
  for (int j = 0; j < c; j++) {
    const uint64_t wj = wsrc[j ^ src];  // uniform address: one broadcast load
    rw += wj > w ? 1 : 0;
    r0 += wj > b0 ? 1 : 0;
    r1 += wj > b1 ? 1 : 0;
  }
  // how many of the nb sorted best keys beat w: invariant best[0..lo) > w > best[hi..)
  int lo = 0, hi = nb;
  const int steps = nb > 32 ? 7 : 6;
  for (int t = 0; t < steps; t++) {
    const int mid = min((lo + hi) >> 1, 63);
    uint64_t bm = shfl_u64(b0, mid & 31);
    if (nb > 32) {
      const uint64_t bm1 = shfl_u64(b1, mid & 31);
      if (mid >= 32) bm = bm1;
    }
    if (lo < hi) {
      if (bm > w) lo = mid + 1;
      else hi = mid;
    }
  }
  rw += lo;
  const int p0 = lane + r0, p1 = 32 + lane + r1;
  const bool has_w = lane < c, has0 = lane < nb, has1 = lane + 32 < nb;
  if (has_w && rw < kp) bsrc[(rw & 32) | ((rw ^ src) & 31)] = w;
  if (has0 && r0 != 0 && p0 < kp) bsrc[(p0 & 32) | ((p0 ^ src) & 31)] = b0;
  if (has1 && r1 != 0 && p1 < kp) bsrc[(p1 & 32) | ((p1 ^ src) & 31)] = b1;
  uint64_t kth = 0ull;
  if (nb + c >= kp) {  // uniform
    const unsigned mw = __ballot_sync(0xFFFFFFFFu, has_w && rw == kp - 1);
    const unsigned m0 = __ballot_sync(0xFFFFFFFFu, has0 && p0 == kp - 1);
    const unsigned m1 = __ballot_sync(0xFFFFFFFFu, has1 && p1 == kp - 1);
    const uint64_t kw = shfl_u64(w, mw ? __ffs(mw) - 1 : 0);
    const uint64_t k0 = shfl_u64(b0, m0 ? __ffs(m0) - 1 : 0);
    const uint64_t k1 = shfl_u64(b1, m1 ? __ffs(m1) - 1 : 0);
    kth = mw ? kw : (m0 ? k0 : k1);
  }
  __syncwarp();  // the stores above are ordered before the next fold's loads of the same list
  return kth;
}

// ---- register sorting networks (one query per lane, static indices only) -------------------
__device__ __forceinline__ void ce_desc(float& a, float& b) {  // a >= b afterwards
  const float hi = fmaxf(a, b), lo = fminf(a, b);
  a = hi;
  b = lo;
}
__device__ __forceinline__ void reg_sort32_desc(float (&v)[32]) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < 32; i++) {
        const int l = i ^ j;
        if (l > i) {
          if ((i & k) == 0) ce_desc(v[i], v[l]);
          else ce_desc(v[l], v[i]);
        }
      }
    }
  }
}
__device__ __forceinline__ void reg_merge32_desc(float (&v)[32]) {  // bitonic -> sorted descending
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
#pragma unroll
    for (int i = 0; i < 32; i++) {
      const int l = i ^ j;
      if (l > i) ce_desc(v[i], v[l]);
    }
  }
}
// 32 accumulator columns of this lane's query, scaled by 1/||x||; rows that must never be selected
// (inverse norm = NaN) read as -inf
template <bool SCALED>
__device__ __forceinline__ void load32_scaled(uint32_t taddr, const float* inv, float (&s)[32]) {
  uint32_t va[16], vb[16];
  tmem_ld16(taddr, va);
  tmem_ld16(taddr + 16, vb);
  tmem_ld_wait();
  if constexpr (!SCALED) {  // pre-normalised rows: the accumulator is the score
#pragma unroll
    for (int i = 0; i < 16; i++) {
      s[i] = __uint_as_float(va[i]);
      s[16 + i] = __uint_as_float(vb[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 wa = *reinterpret_cast<const float4*>(inv + i);
      const float4 wb = *reinterpret_cast<const float4*>(inv + 16 + i);
      s[i] = __uint_as_float(va[i]) * wa.x;
      s[i + 1] = __uint_as_float(va[i + 1]) * wa.y;
      s[i + 2] = __uint_as_float(va[i + 2]) * wa.z;
      s[i + 3] = __uint_as_float(va[i + 3]) * wa.w;
      s[16 + i] = __uint_as_float(vb[i]) * wb.x;
      s[16 + i + 1] = __uint_as_float(vb[i + 1]) * wb.y;
      s[16 + i + 2] = __uint_as_float(vb[i + 2]) * wb.z;
      s[16 + i + 3] = __uint_as_float(vb[i + 3]) * wb.w;
    }
  }
}
// First tile of a CTA: a score T such that at least K' of this lane's 256 scores are >= T, found without
// touching the candidate lists. K' <= 32: the exact K'-th largest — a running sorted top-32 in registers,
// each new group of 32 scores is sorted by a network and merged (half-cleaner + bitonic merge).
// 32 < K' <= 64: the same over the 128 minima of adjacent score pairs — the ceil(K'/2)-th largest minimum
// has that many PAIRS, so at least K' scores, at or above it. Branch-free: ~5k instructions per warp, once.
template <bool SCALED>
__device__ __noinline__ float k2p_first_tile_threshold(uint32_t taddr, const float* inv, int kp) {
  float R[32];
#pragma unroll
  for (int i = 0; i < 32; i++) R[i] = -CUDART_INF_F;
  const bool by_pairs = kp > 32;
#pragma unroll 1
  for (uint32_t c0 = 0; c0 < TILE_N; c0 += by_pairs ? 64u : 32u) {
    float C[32];
    if (!by_pairs) {
      load32_scaled<SCALED>(taddr + c0, inv + c0, C);
#pragma unroll
      for (int i = 0; i < 32; i++) C[i] = fmaxf(C[i], -CUDART_INF_F);  // NaN -> -inf
    } else {
      float s[32];
      load32_scaled<SCALED>(taddr + c0, inv + c0, s);
#pragma unroll
      for (int i = 0; i < 16; i++) C[i] = fminf(fmaxf(s[2 * i], -CUDART_INF_F), fmaxf(s[2 * i + 1], -CUDART_INF_F));
      load32_scaled<SCALED>(taddr + c0 + 32, inv + c0 + 32, s);
#pragma unroll
      for (int i = 0; i < 16; i++) C[16 + i] = fminf(fmaxf(s[2 * i], -CUDART_INF_F), fmaxf(s[2 * i + 1], -CUDART_INF_F));
    }
    reg_sort32_desc(C);
#pragma unroll
    for (int i = 0; i < 32; i++) R[i] = fmaxf(R[i], C[31 - i]);
    reg_merge32_desc(R);
  }
  const int want = (by_pairs ? (kp + 1) / 2 : kp) - 1;
  float T = R[31];
#pragma unroll
  for (int i = 0; i < 31; i++)
    if (i == want) T = R[i];
  return T;
}

// TF32 = false: 16-bit operands, kind::f16, UMMA K = 16: fp16( q/||q|| ) x the fp16 shadow of the normalised fp32 rows
//               (SCALED = false: the accumulator IS the cosine estimate), or bf16( q/||q|| ) x the bf16 shadow / the
//               bf16 corpus itself (SCALED = true: times 1/||x|| in the epilogue). P.idesc names the format.
// TF32 = true : the fp32 corpus and the fp32 queries themselves, converted to tf32 by TMA on the way into
//               shared memory (CU_TENSOR_MAP_DATA_TYPE_TFLOAT32), kind::tf32 UMMA K = 8 — batched scoring
//               of an fp32 index without the extra memory of a shadow (HBM streams 4 B/element). SCALED = true.
// SHARE: how many CTA pairs form one cluster and what they share (launched with the matching cluster shape):
//   0  cluster = one pair (2,1,1). Every CTA fetches its own 16 KB of queries and 16 KB of rows per k-slice: 64 B per
//      SM-cycle at the full tensor rate, 9.5 KB/cycle over 148 SMs — ABOVE the ~6.3 KB/cycle the L2 slices deliver
//      (B300_MICROARCH.md "LTS throughput cap"), so the pair kernel is L2-feed bound at ~2/3 of the tensor peak.
//   1  cluster = two pairs with DIFFERENT query groups walking the SAME corpus tiles (2,2,1): each CTA fetches half of
//      its 128 corpus rows and TMA-multicasts it to the CTA of the same parity in the other pair — 24 KB per k-slice.
//   2  cluster = two pairs of the SAME query group walking DIFFERENT tiles (4,1,1): the queries are the shared operand.
//   3, 4  the same two with FOUR pairs per cluster, (2,4,1) and (8,1,1): each CTA fetches a quarter of the shared operand
//      (20 KB per k-slice). B300_MICROARCH.md: L2 merges identical requests of up to ~4 CTAs anyway, so multicast only
//      starts to save L2 bandwidth beyond a cluster of 4.
// Stage hand-back: a producer multicasts into the other pairs' shared memory too, so empty[s] counts the commits of
// ALL the cluster's pairs' MMAs (tcgen05.commit multicast to every CTA of the cluster).
template <bool TF32, bool SCALED, int SHARE>
__global__ void __launch_bounds__(KP_THREADS, 1)
k2_pair_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x,
               const __grid_constant__ CUtensorMap map_h, const kp_params P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // layout (identical in both CTAs): [stages][A | B] · windows [128][WIN] u64 · best [128][32 or 64] u64 ·
  // inv [2][256] f32 · barriers · tmem ptr
  unsigned char* stage_base = smem;
  // P.nolists (diagnostics, RAGERA_K2_MODE=2 with RAGERA_K2_STAGES): no selection runs, so the windows and best lists
  // get no memory of their own and the pipeline can be measured with more stages than the product has room for
  uint64_t* win = reinterpret_cast<uint64_t*>(P.nolists ? smem : smem + (size_t)P.stages * STAGE_BYTES);
  uint64_t* best = win + (size_t)CTA_M * WIN;
  const uint32_t best_stride = P.kp > 32 ? 64u : 32u;  // sorted K' best per query, lane l owns ranks l and 32+l
  float* s_inv = P.nolists ? reinterpret_cast<float*>(smem + (size_t)P.stages * STAGE_BYTES)
                           : reinterpret_cast<float*>(best + (size_t)CTA_M * best_stride);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_inv + 2 * TILE_N);
  uint64_t* full = bars;                        // [stages]  (the leader's are used)
  uint64_t* empty = bars + MAX_STAGES;          // [stages]  (local)
  uint64_t* tmem_full = bars + 2 * MAX_STAGES;  // [2]       (local)
  uint64_t* tmem_empty = tmem_full + 2;         // [2]       (the leader's are used)
  uint64_t* inv_full = tmem_empty + 2;          // [2]       (local)
  uint64_t* inv_empty = inv_full + 2;           // [2]       (local)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(inv_empty + 2);
  volatile uint32_t* gate = tmem_ptr + 1;  // [0] tiles the slowest query group has started  [1] this CTA's producer is done

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool SB = SHARE == 1 || SHARE == 3;   // the corpus tile is the shared operand
  constexpr bool SA = SHARE == 2 || SHARE == 4;   // the queries are
  constexpr uint32_t CL = SHARE == 0 ? 1u : (SHARE <= 2 ? 2u : 4u);   // pairs per cluster
  const uint32_t crank = cluster_ctarank();  // 0..2*CL-1; x is the fastest cluster dimension, so the pair is (crank & ~1, crank | 1)
  const uint32_t rank = crank & 1u;          // CTA within its pair
  const uint32_t pic = crank >> 1;           // pair within the cluster
  const uint32_t lead = crank & ~1u;         // cluster rank of this pair's leader
  const bool leader = rank == 0;
  const uint32_t pair = blockIdx.x >> 1;
  const uint32_t q0 = blockIdx.y * PAIR_M + rank * CTA_M;  // first query of this CTA
  // pairs of one cluster must run the same number of tiles (they feed each other): a pair that runs out of corpus
  // walks tiles past the end — TMA fills zeros, the epilogue masks rows >= n_rows
  const uint32_t n_iter = SA ? (P.n_tiles + P.pairs - 1) / P.pairs : (P.n_tiles > pair ? (P.n_tiles - pair + P.pairs - 1) / P.pairs : 0u);
  constexpr int BK = TF32 ? ROW_BYTES / 4 : ROW_BYTES / 2;  // k elements per stage
  const bool gated = leader && P.lockstep != 0 && gridDim.y > 1;  // lockstep of the query groups (see the producer)
  const uint32_t nkb = P.ld / BK;
  unsigned long long* cyc = P.cyc ? P.cyc + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * KP_WARPS + warp) * 8 : nullptr;

  if (threadIdx.x == 0) {
    gate[0] = 0u;
    gate[1] = 0u;
    // full[s]: one arrival (the leader's expect_tx of BOTH CTAs' bytes). The peer's TMA completes on the
    // leader's barrier without an arrival of its own; its bytes can only land in the phase they belong
    // to because the peer refills a stage only after the leader's commit has released it.
    for (uint32_t s = 0; s < P.stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], CL); }
    for (int b = 0; b < 2; b++) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 8);  // 4 epilogue warps x 2 CTAs
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
      // query group that reaches it
      uint32_t pf_it = 0, pf_kb = 0;
      // what this CTA fetches of the shared operand: rows [own 128-row half] + pic*64 .. +64, multicast to the CTA of
      // the same parity in both pairs of the cluster
      uint16_t mc_mask = 0;
      for (uint32_t j = 0; j < CL; j++) mc_mask |= (uint16_t)(1u << (rank + 2 * j));
      auto prefetch_next = [&]() {
        if (pf_it < n_iter) {
          const uint32_t t = pair + pf_it * P.pairs;
          if (SB) tma_prefetch_2d(&map_h, (int)(pf_kb * BK), (int)(t * TILE_N + rank * HALF_N + pic * (HALF_N / CL)));
          else tma_prefetch_2d(&map_x, (int)(pf_kb * BK), (int)(t * TILE_N + rank * HALF_N));
          if (++pf_kb == nkb) { pf_kb = 0; pf_it++; }
        }
      };
      for (uint32_t i = 0; i < P.prefetch; i++) prefetch_next();
      for (uint32_t it = 0; it < n_iter; it++) {
        const uint32_t tile = pair + it * P.pairs;
        // Lockstep of the query groups: the gridDim.y pairs with this pair index read the SAME corpus tiles.
        // Left alone they drift apart until L2 no longer holds a tile for the laggard and every group
        // streams the corpus from HBM by itself (measured: DRAM traffic 1.85x algorithmic at 1M rows, K2
        // 25% slower per flop at 25M rows). A pair may run at most P.lockstep tiles (default 1) ahead of the slowest
        // group; the wait is bounded, so it is a pacing hint and can never deadlock.
        // The progress board is polled by warp 2 (idle once TMEM is allocated), which mirrors the slowest group's count
        // into shared memory: the L2 round trips of the poll stay off this thread, whose stalls delay the loads
        // (measured at C2b: polling here cost ~3% of K2 and the MMA thread waited 19% of its cycles for operands; 11% now,
        // and lockstep on and off time the same).
        if (gated) {
          __stcg(P.prog + (size_t)blockIdx.y * P.pairs + pair, it + 1);
          if (it >= P.lockstep) {
            const uint32_t need = it + 1 - P.lockstep;
            const long long t0 = clk();
            while (gate[0] < need && clk() - t0 < 200000) {}
          }
        }
        for (uint32_t kb = 0; kb < nkb; kb++) {
          prefetch_next();
          const long long t0 = cyc ? clk() : 0;
          mbar_wait(&empty[stage], phase ^ 1);
          if (cyc) w_empty += clk() - t0;
          unsigned char* sa = stage_base + (size_t)stage * STAGE_BYTES;
          const uint32_t bar = mapa(smem_u32(&full[stage]), lead);  // the pair leader's full barrier
          if (leader) mbar_expect_tx(&full[stage], 2 * STAGE_BYTES);
          if (SA)   // the queries are shared: fetch 1/CL of this CTA's 128 and multicast them
            tma_load_2d_pair_mc(sa + pic * (A_BYTES / CL), &map_h, smem_u32(&full[stage]), mc_mask, (int)(kb * BK), (int)(q0 + pic * (CTA_M / CL)));
          else
            tma_load_2d_pair(sa, &map_q, bar, (int)(kb * BK), (int)q0);
          if (SB)   // the corpus tile is shared: fetch 1/CL of this CTA's 128 rows and multicast them
            tma_load_2d_pair_mc(sa + A_BYTES + pic * (B_BYTES / CL), &map_h, smem_u32(&full[stage]), mc_mask, (int)(kb * BK),
                                (int)(tile * TILE_N + rank * HALF_N + pic * (HALF_N / CL)));
          else
            tma_load_2d_pair(sa + A_BYTES, &map_x, bar, (int)(kb * BK), (int)(tile * TILE_N + rank * HALF_N));
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
      }
      if (gated) gate[1] = 1u;  // the poller may leave
      if (cyc) { cyc[0] = (unsigned long long)(clk() - t_begin); cyc[1] = (unsigned long long)w_empty; }
    }
  } else if (warp == 2) {
    // ===== lockstep poller (leader CTA only): slowest same-index pair over the query groups -> shared memory =====
    if (gated && lane == 0) {
      uint32_t shown = 0;
      while (gate[1] == 0u) {
        uint32_t slowest = 0xFFFFFFFFu;
        for (uint32_t g = 0; g < gridDim.y; g++) slowest = min(slowest, __ldcg(P.prog + (size_t)g * P.pairs + pair));
        if (slowest != shown) gate[0] = shown = slowest;
        if (slowest >= n_iter) break;
        __nanosleep(128);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (leader && lane == 0) {
      uint32_t stage = 0, phase = 0, it = 0;
      long long w_tmem = 0, w_full = 0;
      const long long t_begin = clk();
      for (; it < n_iter; it++) {
        const uint32_t buf = it & 1;
        long long t0 = cyc ? clk() : 0;
        mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);  // both CTAs' epilogues have drained this accumulator
        if (cyc) w_tmem += clk() - t0;
        tcgen05_fence_after();
        for (uint32_t kb = 0; kb < nkb; kb++) {
          t0 = cyc ? clk() : 0;
          mbar_wait(&full[stage], phase);
          if (cyc) w_full += clk() - t0;
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
                                   P.idesc, (kb | ks) != 0 ? 1u : 0u);
          }
          // every CTA that writes into this pair's stage may refill it once the MMAs have read it: the pair itself,
          // and with a shared operand the other pair of the cluster too
          tcgen05_commit_pair(&empty[stage], (uint16_t)((1u << (2 * CL)) - 1u));
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit_pair(&tmem_full[buf], (uint16_t)(0x3u << lead));  // accumulator complete in both CTAs of the pair
      }
      if (cyc) {
        cyc[0] = (unsigned long long)(clk() - t_begin); cyc[1] = (unsigned long long)w_tmem; cyc[2] = (unsigned long long)w_full;
      }
    }
  } else if (warp == 3 && SCALED) {
    // ===== inverse-norm tile loader (NaN marks rows that must never be selected) =====
    for (uint32_t it = 0; it < n_iter; it++) {
      const uint32_t tile = pair + it * P.pairs;
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
    // ===== epilogue: one lane per query =====
    const uint32_t quarter = warp & 3;
    const uint32_t ql = quarter * 32 + lane;  // query within the CTA
    const uint32_t qg = q0 + ql;              // query in the batch
    const bool live = qg < P.B;
    const int kp = (int)P.kp;
    uint64_t* warp_win = win + (size_t)quarter * 32 * WIN;
    uint64_t* mywin = warp_win + (size_t)lane * WIN;
    uint64_t* warp_best = best + (size_t)quarter * 32 * best_stride;
    float thr = live ? P.floor : CUDART_INF_F;  // a score must beat it to be a candidate (the floor: what the caller filters out anyway)
    uint32_t off = 0;  // bytes used in this lane's window (8 per entry)
    int nbest = 0;     // keys in this lane's sorted best list
    const uint32_t win_addr = smem_u32(mywin);  // 256-byte aligned; entry j lives at (j*8) ^ (lane*8)
    const uint32_t lane8 = (uint32_t)lane << 3;
    const uint32_t taddr_q = tmem_base + ((quarter * 32u) << 16);
    const uint32_t bar_tmem_empty0 = mapa(smem_u32(&tmem_empty[0]), lead);
    const uint32_t bar_tmem_empty1 = mapa(smem_u32(&tmem_empty[1]), lead);
    // Cooperative threshold: the P.pairs CTA pairs that share this query each publish their pub_rank-th best
    // score so far (pairs * pub_rank >= K'). At least K' scored rows are >= the MINIMUM of the published values,
    // so that minimum is a lower bound of the final K'-th best score and nothing below it can be a candidate.
    // The local threshold only knows this CTA's 1/pairs of the corpus; the shared one cuts the survivors (and
    // the window folds) several times. Published values only grow, so any snapshot is valid; ties with the
    // bound are admitted (>=), which keeps the merged top-K' independent of timing.
    uint32_t* my_pub = P.pub + (size_t)pair * P.Bpub + qg;
    const uint32_t* q_pub = P.pub + qg;
    uint32_t last_pub = 0;
    long long c_wait = 0, c_first = 0, n_fold = 0, n_fold_local = 0, c_fold = 0, n_slow = 0, c_ld = 0;
    const long long t_begin = clk();

    // fold the windows of the lanes in `need` into their best lists and raise those lanes' thresholds
    // called by a lane whose best list just changed
    auto publish = [&]() {
      if (nbest >= (int)P.pub_rank) {
        const uint32_t r = P.pub_rank - 1;
        const uint32_t v = (uint32_t)(warp_best[(size_t)lane * best_stride + ((r & 32) | ((r ^ lane) & 31))] >> 32);
        if (v > last_pub) {
          last_pub = v;
          __stcg(my_pub, v);
        }
      }
    };
    auto fold_lanes = [&](unsigned need) {
      const long long tf0 = cyc ? clk() : 0;
      if (__popc(need) < (int)P.local_min) {
        while (need) {
          const int src = __ffs(need) - 1;
          need &= need - 1;
          n_fold++;
          const int c = (int)(__shfl_sync(0xFFFFFFFFu, off, src) >> 3), nb = __shfl_sync(0xFFFFFFFFu, nbest, src);
          const uint64_t t = k2p_fold_one(warp_win + (size_t)src * WIN, warp_best + (size_t)src * best_stride, c, nb, kp, src, lane);
          if (lane == src) {
            nbest = min(kp, nb + c);
            off = 0;
            // rows arrive in increasing order, so a later equal score can never displace an earlier one: the
            // strict float compare against the K'-th best is exact
            if (t != 0ull) thr = fmaxf(thr, rag_key_score(t));
            publish();
          }
        }
      } else {
        // several at once (the first tiles of a CTA): every lane folds its own window, side by side
        n_fold_local++;
        const int c = (int)(off >> 3);
        if (__any_sync(0xFFFFFFFFu, c != 0)) {
          const uint64_t t = k2p_fold_local(mywin, warp_best + (size_t)lane * best_stride, c, nbest, kp, lane);
          if (c != 0) {
            nbest = min(kp, nbest + c);
            off = 0;
            if (t != 0ull && nbest >= kp) thr = fmaxf(thr, rag_key_score(t));
            publish();
          }
        }
        __syncwarp();
      }
      if (cyc) c_fold += clk() - tf0;
    };
    // 16 columns: branch-free predicated appends (raw entry = score bits << 32 | row), then make room.
    // Precondition: every lane has at least 16 free entries.
    auto scan16 = [&](const float* sc, uint32_t row) {
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const uint32_t addr = win_addr | (off ^ lane8);
        asm volatile(
            "{\n\t.reg .pred q;\n\t"
            "setp.gt.f32 q, %1, %2;\n\t"
            "@q st.shared.v2.b32 [%3], {%4, %5};\n\t"
            "@q add.u32 %0, %0, 8;\n\t}"
            : "+r"(off)
            : "f"(sc[i]), "f"(thr), "r"(addr), "r"(row + i), "r"(__float_as_uint(sc[i]))
            : "memory");
      }
      __syncwarp();  // the appends are visible to the lanes that help fold
      const unsigned need = __ballot_sync(0xFFFFFFFFu, off > (WIN - 16) * 8);
      if (need) fold_lanes(need);
    };

    for (uint32_t it = 0; it < n_iter; it++) {
      const uint32_t tile = pair + it * P.pairs;
      const uint32_t buf = it & 1, par = (it >> 1) & 1;
      const long long tw = cyc ? clk() : 0;
      // refresh the shared bound while the tile is still being accumulated (the loads fly during the wait)
      if (live && it != 0 && it % P.pub_every == 0 && P.mode == 0) {
        uint32_t g = 0xFFFFFFFFu;
#pragma unroll 6
        for (uint32_t p = 0; p < P.pairs; p++) g = min(g, __ldcg(q_pub + (size_t)p * P.Bpub));
        if (g > 1u) thr = fmaxf(thr, rag_unorder_f32(g - 1u));  // admit scores >= the bound
      }
      if (SCALED) mbar_wait(&inv_full[buf], par);
      mbar_wait(&tmem_full[buf], par);
      if (cyc) c_wait += clk() - tw;
      tcgen05_fence_after();
      const float* inv = s_inv + buf * TILE_N;
      const uint32_t taddr = taddr_q + buf * TILE_N;
      const uint32_t row0 = tile * TILE_N;
      if (P.mode != 2) {
        if (it == 0 && P.mode == 0) {
          const long long tf = cyc ? clk() : 0;
          const float T = k2p_first_tile_threshold<SCALED>(taddr, inv, kp);
          // admit scores >= T: the threshold is the next float below T
          if (live && T > -CUDART_INF_F) thr = fmaxf(thr, rag_unorder_f32(rag_order_f32(T) - 1u));
          if (cyc) c_first = clk() - tf;
        }
#pragma unroll 1
        for (uint32_t c0 = 0; c0 < TILE_N; c0 += 32) {
          float s[32];
          const long long tl0 = cyc ? clk() : 0;
          load32_scaled<SCALED>(taddr + c0, inv + c0, s);
          if (cyc) c_ld += clk() - tl0;
          const uint32_t row = row0 + c0;
          if (row + 32 > P.n_rows) {  // the corpus ends inside (or before) these columns: rows that do not exist never score
#pragma unroll
            for (int i = 0; i < 32; i++)
              if (row + i >= P.n_rows) s[i] = __int_as_float(0x7FC00000);
          }
          if (P.dbg_scores && live) {
#pragma unroll
            for (int i = 0; i < 32; i++)
              if (row + i < P.n_rows) P.dbg_scores[(size_t)qg * P.n_rows + row + i] = s[i];
          }
          if (P.mode != 0) continue;
          // common case once the thresholds have risen: nothing in these 32 columns beats any query's
          // threshold (fmaxf drops NaN)
          float m0 = fmaxf(s[0], s[1]), m1 = fmaxf(s[2], s[3]), m2 = fmaxf(s[4], s[5]), m3 = fmaxf(s[6], s[7]);
#pragma unroll
          for (int i = 8; i < 32; i += 4) {
            m0 = fmaxf(m0, s[i]);
            m1 = fmaxf(m1, s[i + 1]);
            m2 = fmaxf(m2, s[i + 2]);
            m3 = fmaxf(m3, s[i + 3]);
          }
          if (!__any_sync(0xFFFFFFFFu, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) > thr)) continue;
          n_slow++;
          scan16(s, row);
          scan16(s + 16, row + 16);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(buf ? bar_tmem_empty1 : bar_tmem_empty0);  // the leader's barrier (also from the leader itself)
        if (SCALED) mbar_arrive(&inv_empty[buf]);
      }
    }
    if (cyc && lane == 0) {
      cyc[0] = (unsigned long long)(clk() - t_begin); cyc[1] = (unsigned long long)c_wait; cyc[2] = (unsigned long long)c_first;
      cyc[3] = (unsigned long long)n_fold; cyc[4] = (unsigned long long)c_fold; cyc[5] = (unsigned long long)n_slow; cyc[6] = (unsigned long long)c_ld; cyc[7] = (unsigned long long)n_fold_local;
    }
    // fold what is left in the windows
    fold_lanes(__ballot_sync(0xFFFFFFFFu, live && off != 0));
    __syncwarp();
    // publish partial[q][pair][K']: sorted, zero padded when fewer than K' candidates exist
    for (int src = 0; src < 32; src++) {
      if (q0 + quarter * 32 + src >= P.B) break;
      const int nb = __shfl_sync(0xFFFFFFFFu, nbest, src);
      uint64_t* out = P.partial + ((size_t)(q0 + quarter * 32 + src) * P.parts + pair) * P.kp;
      for (int r = lane; r < kp; r += 32) out[r] = r < nb ? warp_best[(size_t)src * best_stride + ((r & 32) | ((r ^ src) & 31))] : 0ull;
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
  uint32_t stage_cap = 5;   // RAGERA_K2_STAGES (diagnostics): 3..7; 6 and 7 only fit without the lists (mode 2)
  bool nolists = false;
  uint32_t prefetch = KP_PREFETCH;
  uint32_t local_min = 3;
  uint32_t lockstep = 1;
  int cluster = 0;        // RAGERA_K2_CLUSTER: 0 = pairs only, 1 / 2 = share an operand across two / four pairs when the shape allows
  int max_clusters[5] = {0, 0, 0, 0, 0};  // co-resident clusters per SHARE variant (cudaOccupancyMaxActiveClusters), lazily
  bool prof = false;
  unsigned long long* d_cyc = nullptr;
  uint32_t* d_pub = nullptr;  // cooperative-threshold board [pairs][Bpub]
  size_t c_pub = 0;
};

size_t kp_smem_bytes(const kp_state* st, uint32_t stages, uint32_t kp) {
  const size_t lists = st->nolists ? 0 : (size_t)CTA_M * WIN * 8 + (size_t)CTA_M * (kp > 32 ? 64 : 32) * 8;
  return (size_t)stages * STAGE_BYTES + lists + 2 * TILE_N * 4 + (2 * MAX_STAGES + 8) * 8 + 16;
}

uint32_t kp_pick_stages(const kp_state* st, uint32_t kp) {
  uint32_t s = st->stage_cap;
  while (s > 0 && kp_smem_bytes(st, s, kp) > (size_t)st->max_smem) s--;
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
  if (const char* m = getenv("RAGERA_K2_STAGES")) {
    const int v = atoi(m);
    if (v >= 3 && v <= MAX_STAGES) st->stage_cap = (uint32_t)v;
    st->nolists = st->mode == 2 && v > 5;
  }
  if (const char* m = getenv("RAGERA_K2_PROF")) st->prof = atoi(m) != 0;
  if (const char* m = getenv("RAGERA_K2_PREFETCH")) st->prefetch = (uint32_t)atoi(m);
  if (const char* m = getenv("RAGERA_K2_LOCAL_MIN")) st->local_min = (uint32_t)atoi(m);
  if (const char* m = getenv("RAGERA_K2_LOCKSTEP")) st->lockstep = (uint32_t)atoi(m);
  if (const char* m = getenv("RAGERA_K2_CLUSTER")) st->cluster = atoi(m);
  cudaDeviceGetAttribute(&st->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, idx->device);
  idx->k2p_state = st;
  return RAG_OK;
}

// dtype: 0 = tf32 (fp32 memory), 1 = bf16, 2 = fp16 (the two 16-bit types only differ in how TMA would fill out-of-bounds
// elements, which are zeros either way)
int kp_make_map(kp_state* st, CUtensorMap* m, const void* base, uint64_t rows, uint32_t ld, uint32_t box_rows, int dtype) {
  const bool tf32 = dtype == 0;
  const uint32_t es = tf32 ? 4 : 2;
  cuuint64_t dims[2] = {ld, rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * es};
  cuuint32_t box[2] = {ROW_BYTES / es, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : (dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  CUresult r = st->encode(m, dt, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return rag_set_error(RAG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return RAG_OK;
}

// the kernel variant of an index: <TF32, SCALED, SHARE>
typedef void (*kp_kernel_t)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const kp_params);
kp_kernel_t kp_kernel(const rag_index* idx, int share) {
  const bool tf32 = idx->shadow == nullptr;
#define KP_PICK(T, S) (share == 0 ? k2_pair_kernel<T, S, 0> : share == 1 ? k2_pair_kernel<T, S, 1> : share == 2 ? k2_pair_kernel<T, S, 2> \
                       : share == 3 ? k2_pair_kernel<T, S, 3> : k2_pair_kernel<T, S, 4>)
  if (tf32) return KP_PICK(true, true);
  return idx->shadow_f16 ? KP_PICK(false, false) : KP_PICK(false, true);
#undef KP_PICK
}
dim3 kp_cluster_dim(int share) {
  return share == 1 ? dim3(2, 2, 1) : share == 2 ? dim3(4, 1, 1) : share == 3 ? dim3(2, 4, 1) : share == 4 ? dim3(8, 1, 1) : dim3(2, 1, 1);
}
uint32_t kp_cluster_pairs(int share) { return share == 0 ? 1u : (share <= 2 ? 2u : 4u); }

int kp_launch_cfg(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, dim3 grid, size_t smem, cudaStream_t stream, int share) {
  *cfg = cudaLaunchConfig_t();
  cfg->gridDim = grid;
  cfg->blockDim = dim3(KP_THREADS);
  cfg->dynamicSmemBytes = smem;
  cfg->stream = stream;
  attr->id = cudaLaunchAttributeClusterDimension;
  const dim3 c = kp_cluster_dim(share);
  attr->val.clusterDim.x = c.x;
  attr->val.clusterDim.y = c.y;
  attr->val.clusterDim.z = c.z;
  cfg->attrs = attr;
  cfg->numAttrs = 1;
  return RAG_OK;
}

// Which variant runs a batch of `groups` query groups, and with how many tile walkers (CTA pairs) per group.
// Sharing needs two pairs per cluster: two query groups on the same tiles (groups even), or — one group only — two
// walkers. A 4-CTA cluster does not pack the chip as well as a pair does (GPCs of 16/18/20 SMs), so the walkers are
// sized from the number of clusters that are co-resident.
int kp_shape(rag_index* idx, uint32_t groups, uint32_t kp, int* share, uint32_t* pairs) {
  kp_state* st = (kp_state*)idx->k2p_state;
  const uint32_t all_pairs = (uint32_t)idx->sm_count / 2;
  const uint64_t n_tiles = (idx->rows + TILE_N - 1) / TILE_N;
  int sh = 0;
  if (st->cluster == 1) sh = (groups % 2 == 0) ? 1 : (groups == 1 && n_tiles >= 2 ? 2 : 0);
  if (st->cluster >= 2) sh = (groups % 4 == 0) ? 3 : (groups % 2 == 0) ? 1 : (groups == 1 && n_tiles >= 4 ? 4 : 0);
  const uint32_t cl = kp_cluster_pairs(sh);
  uint64_t p = all_pairs / groups;
  if (sh) {
    if (!st->attr_set) return rag_set_error(RAG_ERR_STATE, "k2: kernel attributes not set");
    if (st->max_clusters[sh] == 0) {
      cudaLaunchConfig_t cfg;
      cudaLaunchAttribute attr;
      const dim3 c = kp_cluster_dim(sh);
      kp_launch_cfg(&cfg, &attr, dim3(c.x * 64, c.y, 1), kp_smem_bytes(st, kp_pick_stages(st, kp), kp), idx->stream, sh);
      int n = 0;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, (const void*)kp_kernel(idx, sh), &cfg);
      if (e != cudaSuccess || n <= 0) { cudaGetLastError(); n = -1; }
      st->max_clusters[sh] = n;
    }
    const int mc = st->max_clusters[sh];
    if (mc <= 0) sh = 0;
    else if (sh == 1 || sh == 3) p = std::min<uint64_t>(p, (uint64_t)mc / (groups / cl));   // a cluster = cl groups on one walker
    else p = std::min<uint64_t>(p, (uint64_t)mc * cl) / cl * cl;                            // a cluster = cl walkers of one group
    if (p < ((sh == 2 || sh == 4) ? cl : 1u)) { sh = 0; p = all_pairs / groups; }
  }
  if (p > n_tiles) p = (sh == 2 || sh == 4) ? (n_tiles / cl * cl) : n_tiles;
  if (p < 1) p = 1;
  *share = sh;
  *pairs = (uint32_t)p;
  return RAG_OK;
}

int kp_set_attrs(kp_state* st) {
  if (st->attr_set) return RAG_OK;
#define KP_ATTR(T, S, H) RAG_CUDA(cudaFuncSetAttribute(k2_pair_kernel<T, S, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, st->max_smem))
  KP_ATTR(false, false, 0); KP_ATTR(false, false, 1); KP_ATTR(false, false, 2); KP_ATTR(false, false, 3); KP_ATTR(false, false, 4);
  KP_ATTR(false, true, 0);  KP_ATTR(false, true, 1);  KP_ATTR(false, true, 2);  KP_ATTR(false, true, 3);  KP_ATTR(false, true, 4);
  KP_ATTR(true, true, 0);   KP_ATTR(true, true, 1);   KP_ATTR(true, true, 2);   KP_ATTR(true, true, 3);   KP_ATTR(true, true, 4);
#undef KP_ATTR
  st->attr_set = true;
  return RAG_OK;
}

}  // namespace

// 16-bit operand (bf16 corpus / fp16 or bf16 shadow), or the fp32 corpus read as tf32
int k2_available(const rag_index* idx) {
  if (!idx->inv_norm) return 0;
  return idx->shadow != nullptr || idx->desc.dtype == RAG_F32;
}

int k2_plan(rag_index* idx, uint32_t B, uint32_t kp, uint32_t* parts) {
  RAG_CHECK(kp_init(idx));
  kp_state* st = (kp_state*)idx->k2p_state;
  if (kp > KP_MAX_KP) return rag_set_error(RAG_ERR_UNSUPPORTED, "tensor path keeps at most %d candidates per query (K'=%u)", KP_MAX_KP, kp);
  if (kp_pick_stages(st, kp) < 3) return rag_set_error(RAG_ERR_UNSUPPORTED, "tensor path: not enough shared memory");
  const uint32_t groups = (B + PAIR_M - 1) / PAIR_M;
  const uint32_t all_pairs = (uint32_t)idx->sm_count / 2;
  if (groups > all_pairs)
    return rag_set_error(RAG_ERR_UNSUPPORTED, "tensor path: batch %u exceeds %u queries per launch", B, all_pairs * PAIR_M);
  RAG_CHECK(kp_set_attrs(st));
  int share = 0;
  RAG_CHECK(kp_shape(idx, groups, kp, &share, parts));
  return RAG_OK;
}

int k2_launch(rag_index* idx, uint32_t B, uint32_t kp, uint32_t parts, float floor) {
  if (idx->rows == 0) return rag_set_error(RAG_ERR_STATE, "search on an empty index");
  if (idx->rows >= 0xFFFFFF00ull) return rag_set_error(RAG_ERR_UNSUPPORTED, "more than 2^32-257 rows per shard");
  RAG_CHECK(kp_init(idx));
  kp_state* st = (kp_state*)idx->k2p_state;
  rag_batch* bt = idx->cur;
  const bool tf32 = idx->shadow == nullptr;  // fp32 index without a shadow: score the fp32 rows as tf32
  if (tf32 && idx->desc.dtype != RAG_F32) return rag_set_error(RAG_ERR_STATE, "tensor path: no bf16 operand");
  const uint32_t groups = (B + PAIR_M - 1) / PAIR_M;
  RAG_CHECK(kp_set_attrs(st));
  int share = 0;
  uint32_t pairs_now = 0;
  RAG_CHECK(kp_shape(idx, groups, kp, &share, &pairs_now));
  if (pairs_now != parts) return rag_set_error(RAG_ERR_STATE, "k2_launch: parts=%u but the plan says %u", parts, pairs_now);
  CUtensorMap map_q, map_x, map_h;
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
    RAG_CHECK(q_operand_launch(idx, B, Bpad, false));
    RAG_CHECK(kp_make_map(st, &map_q, bt->d_qb, Bpad, idx->ld, CTA_M, idx->shadow_f16 ? 2 : 1));
    RAG_CHECK(kp_make_map(st, &map_x, idx->shadow, idx->rows, idx->ld, HALF_N, idx->shadow_f16 ? 2 : 1));
    // the shared operand in 64-row boxes (half of what a CTA needs: the other half arrives by multicast)
    const uint32_t cl = kp_cluster_pairs(share) > 1 ? kp_cluster_pairs(share) : 2;
    if (share == 2 || share == 4) RAG_CHECK(kp_make_map(st, &map_h, bt->d_qb, Bpad, idx->ld, CTA_M / cl, idx->shadow_f16 ? 2 : 1));
    else RAG_CHECK(kp_make_map(st, &map_h, idx->shadow, idx->rows, idx->ld, HALF_N / cl, idx->shadow_f16 ? 2 : 1));
  } else {
    // the fp32 queries as staged ([B][ld], zero padded columns); rows past B read as zeros (TMA OOB fill)
    RAG_CHECK(q_operand_launch(idx, B, B, true));
    RAG_CHECK(kp_make_map(st, &map_q, bt->d_q, B, idx->ld, CTA_M, 0));
    RAG_CHECK(kp_make_map(st, &map_x, idx->corpus, idx->rows, idx->ld, HALF_N, 0));
    const uint32_t cl = kp_cluster_pairs(share) > 1 ? kp_cluster_pairs(share) : 2;
    if (share == 2 || share == 4) RAG_CHECK(kp_make_map(st, &map_h, bt->d_q, B, idx->ld, CTA_M / cl, 0));
    else RAG_CHECK(kp_make_map(st, &map_h, idx->corpus, idx->rows, idx->ld, HALF_N / cl, 0));
  }

  rag_prof_scope ps(idx, RAG_PROF_TENSOR);
  kp_params P;
  P.n_rows = (uint32_t)idx->rows;
  P.ld = idx->ld;
  P.B = B;
  P.kp = kp;
  P.parts = parts;
  P.pairs = parts;
  P.stages = kp_pick_stages(st, kp);
  P.n_tiles = (uint32_t)((idx->rows + TILE_N - 1) / TILE_N);
  P.inv_norm = idx->inv_norm;
  P.floor = floor;
  // both operands in one 16-bit format: fp16 x fp16 (fp16 shadow) or bf16 x bf16 (bf16 corpus / bf16 shadow) — a mixed
  // descriptor raises an illegal-instruction error on sm_100a
  P.idesc = idesc_f16(PAIR_M, TILE_N, idx->shadow_f16 ? 0u : 1u, idx->shadow_f16 ? 0u : 1u);
  P.partial = bt->d_partial;
  {
    const uint32_t groups_ = (B + PAIR_M - 1) / PAIR_M;
    P.Bpub = groups_ * PAIR_M;
    P.pub_rank = (kp + P.pairs - 1) / P.pairs;  // pairs * pub_rank >= K'
    P.pub_every = P.pairs <= 24 ? 1 : 4;
    const size_t pub_bytes = (size_t)P.pairs * P.Bpub * sizeof(uint32_t);
    const size_t need = pub_bytes + (size_t)groups_ * P.pairs * sizeof(uint32_t);  // + the lockstep board
    if (need > st->c_pub) {
      if (st->d_pub) RAG_CUDA(cudaFree(st->d_pub));
      st->d_pub = nullptr;
      st->c_pub = 0;
      RAG_CUDA(cudaMalloc((void**)&st->d_pub, need));
      st->c_pub = need;
    }
    RAG_CUDA(cudaMemsetAsync(st->d_pub, 0, need, idx->stream));
    P.pub = st->d_pub;
    P.prog = st->d_pub + pub_bytes / sizeof(uint32_t);
    P.lockstep = st->lockstep;
  }
  P.dbg_scores = st->dbg;
  P.mode = st->mode;
  P.nolists = st->nolists ? 1u : 0u;
  P.prefetch = st->prefetch;
  P.local_min = st->local_min;
  P.cyc = nullptr;
  const size_t n_cyc = (size_t)P.pairs * 2 * ((B + PAIR_M - 1) / PAIR_M) * KP_WARPS * 8;
  if (st->prof) {
    if (st->d_cyc) cudaFree(st->d_cyc);
    RAG_CUDA(cudaMalloc((void**)&st->d_cyc, n_cyc * 8));
    RAG_CUDA(cudaMemsetAsync(st->d_cyc, 0, n_cyc * 8, idx->stream));
    P.cyc = st->d_cyc;
  }
  const size_t smem = kp_smem_bytes(st, P.stages, kp);
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr;
  kp_launch_cfg(&cfg, &attr, dim3(P.pairs * 2, groups), smem, idx->stream, share);
  RAG_CUDA(cudaLaunchKernelEx(&cfg, kp_kernel(idx, share), map_q, map_x, map_h, P));
  RAG_CUDA(cudaGetLastError());
  idx->launches++;
  if (st->prof) {
    static int printed = 0;
    std::vector<unsigned long long> h(n_cyc);
    RAG_CUDA(cudaMemcpyAsync(h.data(), st->d_cyc, n_cyc * 8, cudaMemcpyDeviceToHost, idx->stream));
    RAG_CUDA(cudaStreamSynchronize(idx->stream));
    if (printed++ % 16 == 8) {
      const size_t nct = n_cyc / (KP_WARPS * 8);
      double acc[2][KP_WARPS][8] = {{{0}}};
      for (size_t c = 0; c < nct; c++)
        for (int w = 0; w < KP_WARPS; w++)
          for (int j = 0; j < 8; j++) acc[c & 1][w][j] += (double)h[(c * KP_WARPS + w) * 8 + j] / (nct / 2);
      fprintf(stderr, "[k2 pair prof] B=%u rows=%u kp=%u stages=%u ctas=%zu (avg cycles per CTA; rank0 | rank1)\n", B, P.n_rows, kp, P.stages, nct);
      fprintf(stderr, "  producer: total %.0f wait_empty %.0f | total %.0f wait_empty %.0f\n", acc[0][0][0], acc[0][0][1], acc[1][0][0], acc[1][0][1]);
      fprintf(stderr, "  mma     : total %.0f wait_tmem_empty %.0f wait_full %.0f\n", acc[0][1][0], acc[0][1][1], acc[0][1][2]);
      for (int w = 4; w < KP_WARPS; w++)
        fprintf(stderr, "  epi w%-2d : total %.0f wait %.0f first-tile %.0f fold1 %.0f fold-local %.0f fold cycles %.0f slow chunks %.0f ld cycles %.0f | wait %.0f fold cycles %.0f\n", w,
                acc[0][w][0], acc[0][w][1], acc[0][w][2], acc[0][w][3], acc[0][w][7], acc[0][w][4], acc[0][w][5], acc[0][w][6], acc[1][w][1], acc[1][w][4]);
    }
  }
  return RAG_OK;
}

void k2_set_debug(rag_index* idx, float* d_scores) {
  if (kp_init(idx) == RAG_OK) ((kp_state*)idx->k2p_state)->dbg = d_scores;
}

void k2_destroy(rag_index* idx) {
  if (idx->k2p_state && ((kp_state*)idx->k2p_state)->d_cyc) cudaFree(((kp_state*)idx->k2p_state)->d_cyc);
  if (idx->k2p_state && ((kp_state*)idx->k2p_state)->d_pub) cudaFree(((kp_state*)idx->k2p_state)->d_pub);
  delete (kp_state*)idx->k2p_state;
  idx->k2p_state = nullptr;
}
