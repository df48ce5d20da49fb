// k4_rescore.cu — K4: exact rescoring of the K' surviving candidates + local top-k.
//
// The reference computes every score in IEEE binary64, left to right, without FMA:
//   similarity(q,x) = dot(q,x) / (norm(q) * norm(x))        (llamaindex, upstream-recalled;
//   reached from src/lib/hybrid-search.ts:223-224). K1/K2 pick K' >= k candidates with
// fp32 / bf16 arithmetic; this kernel recomputes those K' scores with exactly the
// reference's operation order (__dmul_rn / __dadd_rn / __dsqrt_rn / __ddiv_rn, one
// sequential chain per sum), so reported scores are bit-identical to the oracle and the
// final order (score desc, lower chunk id first — the reference's stable sort over
// insertion order) is decided on exact values.
//
// Mapping: one CTA per query, one LANE per candidate. The 32 candidate rows of a warp
// are staged through shared memory with cp.async in 512-byte column chunks (double
// buffered); 16-byte slot c of row j is stored at slot (c ^ j) so that the per-lane
// 128-bit reads (lane j reads row j) are bank-conflict free. The three sums are
// latency-bound dependent chains (D steps each) — 32 candidates advance in lock step.
//
// Also here: certification. Let t be the approximate score of the K'-th candidate and
// eps the selection-error bound of the producing kernel. Every row outside the
// candidate set has exact score <= t + eps, so if the exact k-th score is > t + eps the
// local top-k equals an exact scan. Otherwise the query is flagged and the host
// escalates it to a stronger path.
#include "exact_chain.cuh"
#include "k5_body.cuh"

#include <stdio.h>

namespace {
using namespace rag_exact;
constexpr int K4_WARP_BYTES = WARP_BYTES;

struct k4_meta {
  uint64_t id_base;
  const uint8_t* ctype;
  const double* conf;
  const int32_t* access;
  const int64_t* last_ms;
  const uint64_t* row_keys;
  int64_t now_ms;
  double decay, bonus;
};

// Shared tail of both K4 kernels, executed by one thread per candidate (j < blockDim.x, all threads
// of the CTA must call): exact rank, certification, record write. s_score/s_row/s_tmp are shared
// arrays of >= kp entries that this function fills.
// what k4_finalize reads of a candidate's row besides the row itself; the latency variant loads it right after the merge, so
// that the loads fly under the exact chains instead of sitting (two dependent L2/HBM round trips) on the tail's critical path
struct k4_row_meta {
  uint64_t key;
  uint32_t ctype;
  double conf;
  int32_t access;
  int64_t last_ms;
};
__device__ __forceinline__ k4_row_meta k4_load_row_meta(const k4_meta& M, uint32_t row) {
  k4_row_meta r;
  r.key = M.row_keys ? M.row_keys[row] : M.id_base + row;
  r.ctype = M.ctype ? M.ctype[row] : (uint32_t)RAG_CT_DOCUMENT;
  // the Memory columns are read whatever the row's type turns out to be: independent loads, nobody waits for ctype here
  const bool mem = M.ctype && M.conf && M.access && M.last_ms;
  r.conf = mem ? M.conf[row] : 0.0;
  r.access = mem ? M.access[row] : 0;
  r.last_ms = mem ? M.last_ms[row] : 0;
  return r;
}

__device__ __forceinline__ void k4_finalize(uint32_t j, uint32_t kp, uint32_t k, bool valid, uint32_t row, double score,
                                            double nq, uint64_t last_key, double eps, double floor_u, double min_eff,
                                            int key_has_qnorm, const k4_meta& M, double* s_score, uint32_t* s_row, double* s_kth,
                                            rag_rec* __restrict__ out, uint32_t* __restrict__ out_cnt, rag_rec* s_out = nullptr,
                                            const k4_row_meta* pre = nullptr) {
  // similarity = dot / (norm(q) * norm(x)); NaN (zero norm) is defined as never selected
  const bool good = valid && score == score;
  if (j < kp) {
    s_score[j] = good ? score : -INFINITY;
    s_row[j] = row;
  }
  __syncthreads();

  // exact order: score desc, chunk id asc (== the reference's stable sort over insertion order)
  uint32_t rank = 0, n_good = 0;
  for (uint32_t i = 0; i < kp; i++) {
    const double si = s_score[i];
    const uint32_t ri = s_row[i];
    const bool gi = si != -INFINITY;
    n_good += gi ? 1u : 0u;
    if (gi && (si > score || (si == score && ri < row))) rank++;
  }
  const uint32_t cnt = n_good < k ? n_good : k;

  // certification of the local top-k against rows that were never candidates
  if (good && rank + 1 == cnt) *s_kth = score;
  __syncthreads();
  // U = what the exact score of a row OUTSIDE the candidate set can be at most: the selection score of the K'-th candidate
  // when the list is full, else the floor below which the selecting kernel dropped rows (-inf: every scored row is a
  // candidate), plus the selection-error bound. The local top-k is the exact scan's if its k-th score clears U — or,
  // whatever the candidates are, if U lies below the score the caller filters at (hybridSearch's minVectorScore,
  // hybrid-search.ts:308-314; MemoryStore's minRelevance, store.ts:151): a row that is not a candidate could then never
  // survive the filter, and filtering commutes with taking the best k by the same score.
  bool certified = true;
  double U = -INFINITY;
  if (last_key != 0ull && cnt > 0) {  // list full: rows outside the candidate set exist
    // K1/K2 keys hold dot/||x|| (||q|| cannot change the order); K1x keys hold the cosine itself
    const double t = key_has_qnorm ? (double)rag_key_score(last_key) : (double)rag_key_score(last_key) / sqrt(nq);
    U = t + eps;
  } else if (floor_u > -INFINITY) {
    U = floor_u + eps;
  }
  if (U > -INFINITY) certified = U < min_eff || (cnt == k && *s_kth > U);
  const uint32_t flags = certified ? 0u : 1u;

  if (good && rank < k) {
    const k4_row_meta m = pre ? *pre : k4_load_row_meta(M, row);
    rag_rec r;
    r.score = score;
    r.id = M.id_base + row;
    r.key = m.key;
    r.ctype = m.ctype;
    r.flags = flags;
    r.fresh = 0.0;
    r.conf_pad = 0.0;
    if (r.ctype == RAG_CT_MEMORY && M.conf && M.access && M.last_ms) {
      // calculateFreshnessScore — src/lib/memory/freshness.ts:43-55 (exp/log: <=1 ulp libm variance)
      const double hours = (double)(M.now_ms - m.last_ms) / 3600000.0;
      const double dec = exp(__dmul_rn(-M.decay, hours));
      const double fb = __dmul_rn(log((double)m.access + 1.0), M.bonus);
      const double sc = __dmul_rn(__dmul_rn(m.conf, dec), __dadd_rn(1.0, fb));
      r.fresh = fmax(0.0, fmin(1.0, sc));
    }
    out[rank] = r;
    if (s_out) s_out[rank] = r;
  }
  // empty tail + flags on every slot so the merge sees them even for empty shards
  for (uint32_t i = cnt + j; i < k; i += blockDim.x) {
    rag_rec e;
    e.score = -INFINITY; e.id = ~0ull; e.key = ~0ull; e.fresh = 0.0; e.ctype = 0; e.flags = flags; e.conf_pad = 0.0;
    out[i] = e;
    if (s_out) s_out[i] = e;
  }
  if (j == 0) *out_cnt = cnt;
}

// ---- throughput variant (large batches): one CTA per query, one LANE per candidate ---------------
template <bool BF16>
__global__ void __launch_bounds__(RAG_MAX_CANDIDATES + 32)
k4_rescore_kernel(const void* __restrict__ X, uint32_t ld, const float* __restrict__ Q,
                  const uint64_t* __restrict__ cand, uint32_t kp, uint32_t k, rag_eps E,
                  int key_has_qnorm, k4_meta M, rag_rec* __restrict__ local, uint32_t* __restrict__ local_cnt) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double s_score[RAG_MAX_CANDIDATES];
  __shared__ uint32_t s_row[RAG_MAX_CANDIDATES];
  __shared__ double s_nq, s_kth;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t b = blockIdx.x;
  const uint32_t j = threadIdx.x;  // candidate index
  const uint64_t key = j < kp ? cand[(size_t)b * RAG_MAX_CANDIDATES + j] : 0ull;
  const bool valid = key != 0ull;
  // empty slots re-read the query's best candidate (a row this CTA fetches anyway) instead of all
  // hammering row 0 — thousands of CTAs on one L2 line set is a measurable hot spot
  const uint64_t key0 = cand[(size_t)b * RAG_MAX_CANDIDATES];
  const uint32_t row = valid ? rag_key_row(key) : (key0 != 0ull ? rag_key_row(key0) : 0u);

  // the last warp of the CTA holds no candidates: it runs the ||q||^2 chain meanwhile
  chains c = {0.0, 0.0, 0.0};
  if ((uint32_t)warp * 32 < kp) {
    unsigned char* wsm = smem + (size_t)warp * K4_WARP_BYTES;
    c = warp_exact_sums<BF16>(X, ld, Q + (size_t)b * ld, row, wsm, lane);
  } else {
    const double nq = query_norm_sq(Q + (size_t)b * ld, ld, lane);
    if (lane == 0) s_nq = nq;
  }
  __syncthreads();
  c.nq = s_nq;
  const double eps = E.eps_q ? E.eps + (double)E.eps_q[b] * E.eps_q_mul : E.eps;
  k4_finalize(j, kp, k, valid, row, finish(c), s_nq, cand[(size_t)b * RAG_MAX_CANDIDATES + kp - 1], eps, E.floor, E.min_score, key_has_qnorm, M,
              s_score, s_row, &s_kth, local + (size_t)b * k, local_cnt + b);
}

// ---- latency variant (small batches): K3 + K4 fused, one WARP per candidate, several CTAs per query --
// The three sums of a row are strictly sequential (the reference adds left to right), so one
// candidate is a dependent chain of `ld` fp64 adds no matter how it is mapped. Here the 32 lanes
// of a warp compute the products of 32 consecutive elements in parallel (exact: one rounding
// each, same as the reference's q[i]*x[i]) and then every lane replays the adds in order, fetching
// product i from lane i%32 with a shuffle — 2 SHFL + 1 DADD per element per sum, no shared
// memory, nothing but the add latency on the critical path. The work items of a query — one per (candidate, sum)
// pair, dot and ||x||^2 on DIFFERENT warps, plus one for ||q||^2 — are spread over (2K'+1)/5 CTAs: a warp that
// carries one chain retires an add every ~10.6 cycles, a warp that interleaves two needs 14 per pair
// (tools/micro/dadd_latency.cu), and a batch-1 search spreads over 5-7 SMs instead of one warp of one SM. Every CTA first merges the per-CTA candidate lists of K1/K2
// itself (sorted lists, k-way merge from shared memory — this is K3, fused); the last CTA of a
// query to finish (atomic ticket) ranks, certifies and writes the records.
// diagnostics (RAGERA_SMALL_PROF=1): cycles per phase of the small-batch kernel as seen by the LAST CTA of a query
__device__ int g_k34_prof_on;
__device__ unsigned long long g_k34_prof[8];  // stage + K3 merge, exact chains, ticket, (last CTA) finalize, K5, count

constexpr int K4S_WARPS = 5;                    // every warp takes work items (one exact sum each)
constexpr int K4S_THREADS = K4S_WARPS * 32;
constexpr size_t K4S_MAX_STAGE = 96 * 1024;     // staged candidate keys (parts * K' * 8 bytes)

// One exact left-to-right sum over a row, by one warp. WHICH: 0 = sum q[i]*x[i], 1 = sum x[i]*x[i], 2 = sum q[i]*q[i]
// (the reference's three loops). The 32 lanes compute the (exact) products of a 256-element block in parallel and park
// them in the warp's shared-memory scratch `sp` [2 buffers][256]; the adds are then replayed in order from broadcast
// 128-bit shared loads, so the only thing on the critical path is the fp64 add latency. The global loads of block b+1 fly
// under the first half of block b's adds, its products are formed in the issue slots between the second half's adds.
template <bool BF16, int WHICH>
__device__ __forceinline__ double warp_chain1(const void* __restrict__ X, uint32_t row, uint32_t ld,
                                              const float* __restrict__ q, int lane, double* sp) {
  double s0 = 0.0;
  const int nblk = (int)(ld / 256);
  const size_t row_off = (size_t)row * ld;
  float xr[8], qr[8];
  auto load = [&](int blk) {
#pragma unroll
    for (int t = 0; t < 8; t++) {
      const int e = blk * 256 + t * 32 + lane;
      if (WHICH != 1) qr[t] = q[e];
      if (WHICH != 2) {
        if (BF16) xr[t] = __uint_as_float((uint32_t)reinterpret_cast<const uint16_t*>(X)[row_off + e] << 16);
        else xr[t] = reinterpret_cast<const float*>(X)[row_off + e];
      }
    }
  };
  auto park = [&](int buf) {
    double* p0 = sp + buf * 256;
#pragma unroll
    for (int t = 0; t < 8; t++) {
      const double a = WHICH == 1 ? (double)xr[t] : (double)qr[t];
      const double c = WHICH == 2 ? (double)qr[t] : (double)xr[t];
      p0[t * 32 + lane] = __dmul_rn(a, c);
    }
  };
  auto park_one = [&](int buf, int t) {
    const double a = WHICH == 1 ? (double)xr[t] : (double)qr[t];
    const double c = WHICH == 2 ? (double)qr[t] : (double)xr[t];
    sp[buf * 256 + t * 32 + lane] = __dmul_rn(a, c);
  };
  load(0);
  park(0);
  __syncwarp();
  for (int blk = 0; blk < nblk; blk++) {
    const bool more = blk + 1 < nblk;
    if (more) load(blk + 1);
    const double2* a0 = reinterpret_cast<const double2*>(sp + (blk & 1) * 256);
    // first half of the block's adds: the loads of the next block are in flight under them
#pragma unroll 16
    for (int i = 0; i < 64; i++) {
      const double2 u = a0[i];
      s0 = __dadd_rn(__dadd_rn(s0, u.x), u.y);
    }
    // second half: the next block's products (conversions, multiply, store — independent of the chain) are slotted between
    // the dependent adds, one of this lane's eight per 8 adds, instead of standing between two blocks' chains
#pragma unroll
    for (int t = 0; t < 8; t++) {
      if (more) park_one((blk + 1) & 1, t);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const double2 u = a0[64 + t * 8 + i];
        s0 = __dadd_rn(__dadd_rn(s0, u.x), u.y);
      }
    }
    if (more) __syncwarp();
  }
  return s0;
}

template <bool BF16>
__global__ void __launch_bounds__(K4S_THREADS)
k34_small_kernel(const void* __restrict__ X, uint32_t ld, const float* __restrict__ Q,
                 const uint64_t* __restrict__ partial, uint32_t parts, uint32_t kp, uint32_t k, rag_eps E,
                 int key_has_qnorm, k4_meta M, double* __restrict__ scratch /*[B][2*128+2]*/,
                 unsigned int* __restrict__ ticket /*[B]*/, uint64_t* __restrict__ cand_out /*[B][128]*/,
                 rag_rec* __restrict__ local, uint32_t* __restrict__ local_cnt, int fuse_k5, rag_k5::k5_io io,
                 rag_p2p_view pv) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint64_t* stg = reinterpret_cast<uint64_t*>(smem);               // [parts][kp] staged lists
  uint64_t* wl = stg + (size_t)parts * kp;                         // [2][K4S_WARPS][kp] per-warp merged lists
  __shared__ __align__(16) double s_prod[K4S_WARPS][2 * 2 * 256];  // per-warp product scratch of warp_chain1 (and K5's working set)
  __shared__ __align__(16) rag_rec s_recs[RAG_MAX_TOPK];            // this query's records for the in-place K5
  __shared__ uint64_t s_cand[RAG_MAX_CANDIDATES];
  __shared__ double s_score[RAG_MAX_CANDIDATES];
  __shared__ uint32_t s_row[RAG_MAX_CANDIDATES];
  __shared__ double s_kth;
  __shared__ int s_last;
  __shared__ uint64_t s_kw[RAG_MAX_KEYWORDS];
  __shared__ uint32_t s_kwc;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t b = blockIdx.x, slice = blockIdx.y, nslices = gridDim.y;
  const bool prof = g_k34_prof_on != 0;
  const long long pt0 = prof ? clock64() : 0;

  // ---- K3: merge the sorted per-CTA lists of the scoring kernel (every slice does it: ~1 us) ----
  const uint64_t* in = partial + (size_t)b * parts * kp;
  const uint32_t n_keys = parts * kp;
  if ((n_keys & 1u) == 0 && (reinterpret_cast<uintptr_t>(in) & 15u) == 0) {
    // 128-bit loads, eight per thread issued before the first store: one L2 round trip for the usual 13-38 KB
    const uint4* in4 = reinterpret_cast<const uint4*>(in);
    uint4* st4 = reinterpret_cast<uint4*>(stg);
    const uint32_t n16 = n_keys / 2;
    for (uint32_t i0 = threadIdx.x; i0 < n16; i0 += K4S_THREADS * 8) {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const uint32_t i = i0 + u * K4S_THREADS;
        v[u] = i < n16 ? __ldg(in4 + i) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const uint32_t i = i0 + u * K4S_THREADS;
        if (i < n16) st4[i] = v[u];
      }
    }
  } else {
    for (uint32_t i = threadIdx.x; i < n_keys; i += K4S_THREADS) stg[i] = in[i];
  }
  // (2) the query's keyword list: fetched now by the idle lanes' loads, read by the in-place K5 from shared memory
  if (fuse_k5 && io.kwc && io.a.mode == 0 && threadIdx.x < RAG_MAX_KEYWORDS) {
    const uint32_t nk = min(io.kwc[b], (uint32_t)RAG_MAX_KEYWORDS);
    if (threadIdx.x == 0) s_kwc = nk;
    if (threadIdx.x < nk) s_kw[threadIdx.x] = io.kw[(size_t)b * io.a.kw_stride + threadIdx.x];
  }
  __syncthreads();
  const long long pt0a = prof ? clock64() : 0;
  // warp w merges lists w, w+W, w+2W, ... at most 31 at a time, together with its running result
  // (ping-pong between two per-warp buffers; every warp runs the same number of rounds)
  uint64_t* wl2 = wl + (size_t)K4S_WARPS * kp;
  uint64_t* res = wl;  // where the per-warp results end up
  {
    uint64_t* cur_l = wl + (size_t)warp * kp;
    uint64_t* nxt_l = wl2 + (size_t)warp * kp;
    for (uint32_t i = lane; i < kp; i += 32) cur_l[i] = 0ull;
    __syncwarp();
    for (uint32_t base = 0; base < parts; base += K4S_WARPS * 31) {
      // lane 0 walks the running list, lane l >= 1 walks list base + warp + (l-1)*W
      const uint32_t li = base + warp + (uint32_t)(lane - 1) * K4S_WARPS;
      const uint64_t* lst = lane == 0 ? cur_l : (li < parts ? stg + (size_t)li * kp : nullptr);
      uint32_t head = 0;
      uint64_t cur = lst ? lst[0] : 0ull;
      for (uint32_t r = 0; r < kp; r++) {
        const uint64_t m = warp_max_u64(cur);
        if (lane == 0) nxt_l[r] = m;
        if (m != 0ull && cur == m) {  // keys are unique (the row is part of the key)
          head++;
          cur = head < kp ? lst[head] : 0ull;
        }
      }
      __syncwarp();
      uint64_t* t = cur_l; cur_l = nxt_l; nxt_l = t;
      res = res == wl ? wl2 : wl;
    }
  }
  __syncthreads();
  if (warp == 0) {
    warp_merge_lists(res, K4S_WARPS, (int)kp, (int)kp, s_cand, lane);
    if (slice == 0)
      for (uint32_t i = lane; i < RAG_MAX_CANDIDATES; i += 32) cand_out[(size_t)b * RAG_MAX_CANDIDATES + i] = i < kp ? s_cand[i] : 0ull;
  }
  __syncthreads();

  // the candidates' row metadata (fusion key, contentType, Memory columns): loaded now, used by the query's last CTA after the chains
  k4_row_meta pre = k4_row_meta();
  if (threadIdx.x < kp && s_cand[threadIdx.x] != 0ull) pre = k4_load_row_meta(M, rag_key_row(s_cand[threadIdx.x]));
  const long long pt1 = prof ? clock64() : 0;
  // ---- K4: exact sums, one warp per candidate -----------------------------------------------------
  double* my_scratch = scratch + (size_t)b * (2 * RAG_MAX_CANDIDATES + 2);
  const float* q = Q + (size_t)b * ld;
  // work item t: t < 2K' → sum (t & 1) of candidate t >> 1 (0 = dot, 1 = ||x||^2);  t == 2K' → ||q||^2
  for (uint32_t t = slice * K4S_WARPS + warp; t <= 2 * kp; t += nslices * K4S_WARPS) {
    if (t == 2 * kp) {
      const double nq = warp_chain1<BF16, 2>(nullptr, 0, ld, q, lane, s_prod[warp]);
      if (lane == 0) my_scratch[2 * RAG_MAX_CANDIDATES] = nq;
      continue;
    }
    const uint64_t key = s_cand[t >> 1];
    if (key == 0ull) continue;  // warp-uniform
    const double v = (t & 1) ? warp_chain1<BF16, 1>(X, rag_key_row(key), ld, q, lane, s_prod[warp])
                             : warp_chain1<BF16, 0>(X, rag_key_row(key), ld, q, lane, s_prod[warp]);
    if (lane == 0) my_scratch[t] = v;
  }

  // ---- the last slice of this query to arrive finishes it ------------------------------------------
  __threadfence();
  __syncthreads();
  const long long pt2 = prof ? clock64() : 0;
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&ticket[b], 1u);
    s_last = (t == nslices - 1) ? 1 : 0;
    if (s_last) ticket[b] = 0u;  // ready for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  const long long pt3 = prof ? clock64() : 0;
  __threadfence();
  const volatile double* vs = my_scratch;
  const double nq = vs[2 * RAG_MAX_CANDIDATES];
  const double eps = E.eps_q ? E.eps + (double)E.eps_q[b] * E.eps_q_mul : E.eps;
  for (uint32_t j0 = 0; j0 < kp || j0 == 0; j0 += K4S_THREADS) {  // kp <= 128 <= K4S_THREADS: one pass
    const uint32_t j = j0 + threadIdx.x;
    const uint64_t key = j < kp ? s_cand[j] : 0ull;
    const bool valid = key != 0ull;
    const uint32_t row = valid ? rag_key_row(key) : 0u;
    const double score = valid ? __ddiv_rn(vs[2 * j], __dmul_rn(__dsqrt_rn(nq), __dsqrt_rn(vs[2 * j + 1]))) : 0.0;
    k4_finalize(j, kp, k, valid, row, score, nq, s_cand[kp - 1], eps, E.floor, E.min_score, key_has_qnorm, M, s_score, s_row, &s_kth,
                local + (size_t)b * k, local_cnt + b, s_recs, j0 == 0 ? &pre : nullptr);
  }
  // ---- K5 in place: filter + fusion of this query by warp 0 — one launch less on the batch-1 latency path.
  //      Sharded (pv.nranks > 1): the same warp first exchanges this query's records with the peer ranks through
  //      the mailboxes; every query's last CTA is resident (the grid is B x K'/4 <= 32 x 32 CTAs), so a rank's
  //      wait for its peers cannot starve them. The product scratch is free by now and hosts K5's working set.
  const long long pt4 = prof ? clock64() : 0;
  if (fuse_k5) {
    if (pv.nranks > 1) __threadfence();  // sharded: the records above are read back through L2 by the exchange
    __syncthreads();
    static_assert(sizeof(rag_k5::fuse_smem) <= sizeof(s_prod), "K5 working set must fit the product scratch");
    if (warp == 0) {
      rag_k5::fuse_smem& fs = *reinterpret_cast<rag_k5::fuse_smem*>(&s_prod[0][0]);
      rag_k5::k5_io io2 = io;
      if (io.kwc && io.a.mode == 0) {   // the keyword list staged at the top of the kernel: K5 indexes [b * stride + i] / [b]
        io2.kw = s_kw - (size_t)b * io.a.kw_stride;
        io2.kwc = &s_kwc - b;
      }
      if (pv.nranks > 1) rag_k5::k5_fuse_body<false>(fs, rag_k5::p2p_exchange(pv, local, io.a.B, b, io.a.k, lane), io2, b, lane);
      else rag_k5::k5_fuse_body<true>(fs, s_recs, io2, b, lane);   // one GPU: straight from shared memory
    }
  }
  if (prof && threadIdx.x == 0) {
    atomicAdd(&g_k34_prof[0], (unsigned long long)(pt0a - pt0));
    atomicAdd(&g_k34_prof[6], (unsigned long long)(pt1 - pt0a));
    atomicAdd(&g_k34_prof[1], (unsigned long long)(pt2 - pt1));
    atomicAdd(&g_k34_prof[2], (unsigned long long)(pt3 - pt2));
    atomicAdd(&g_k34_prof[3], (unsigned long long)(pt4 - pt3));
    atomicAdd(&g_k34_prof[4], (unsigned long long)(clock64() - pt4));
    atomicAdd(&g_k34_prof[5], 1ull);
  }
}

}  // namespace

void k34_small_prof(int enable, FILE* dump) {
  if (enable >= 0) {
    unsigned long long z[8] = {0};
    cudaMemcpyToSymbol(g_k34_prof_on, &enable, sizeof(int));
    cudaMemcpyToSymbol(g_k34_prof, z, sizeof(z));
  }
  if (dump) {
    unsigned long long h[8];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(h, g_k34_prof, sizeof(h));
    if (h[5])
      fprintf(dump, "[k34_small prof] avg cycles in the LAST CTA of a query over %llu queries: staging the lists %.0f, K3 merge %.0f, exact chains %.0f, "
              "fence + ticket %.0f, rank + certify + records %.0f, K5 in place %.0f\n",
              h[5], (double)h[0] / h[5], (double)h[6] / h[5], (double)h[1] / h[5], (double)h[2] / h[5], (double)h[3] / h[5], (double)h[4] / h[5]);
  }
}

int k4_launch(rag_index* idx, uint32_t B, uint32_t kp, uint32_t k, rag_eps eps, int key_has_qnorm,
              int64_t now_ms, double decay, double bonus) {
  rag_prof_scope ps(idx, RAG_PROF_RESCORE);
  const uint32_t nw = (kp + 31) / 32;
  const size_t smem = (size_t)nw * K4_WARP_BYTES;
  const bool bf16 = idx->desc.dtype == RAG_BF16;
  const k4_meta M = {idx->desc.id_base, idx->ctype, idx->conf, idx->access, idx->last_ms, idx->row_keys, now_ms, decay, bonus};
  auto kern = bf16 ? k4_rescore_kernel<true> : k4_rescore_kernel<false>;
  RAG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * K4_WARP_BYTES)));
  kern<<<B, (nw + 1) * 32, smem, idx->stream>>>(idx->corpus, idx->ld, idx->cur->d_q, idx->cur->d_cand, kp, k, eps, key_has_qnorm,
                                           M, idx->cur->d_local, idx->cur->d_local_cnt);
  RAG_CUDA(cudaGetLastError());
  idx->launches++;
  return RAG_OK;
}

// K3 + K4 in one launch for small batches (see k34_small_kernel). Returns RAG_ERR_UNSUPPORTED-free:
// the caller asks k34_small_ok() first.
bool k34_small_fuses_exchange(const rag_index* idx) { return comm_uses_p2p(idx); }

bool k34_small_ok(const rag_index* idx, uint32_t B, uint32_t kp, uint32_t parts) {
  (void)idx;
  return B <= 32 && (size_t)parts * kp * 8 <= K4S_MAX_STAGE && kp <= RAG_MAX_CANDIDATES;
}

int k34_small_launch(rag_index* idx, uint32_t B, uint32_t kp, uint32_t parts, uint32_t k, rag_eps eps, int key_has_qnorm,
                     int64_t now_ms, double decay, double bonus, const rag_fuse_args* fuse) {
  rag_prof_scope ps(idx, RAG_PROF_RESCORE);
  rag_batch* bt = idx->cur;
  const bool bf16 = idx->desc.dtype == RAG_BF16;
  const k4_meta M = {idx->desc.id_base, idx->ctype, idx->conf, idx->access, idx->last_ms, idx->row_keys, now_ms, decay, bonus};
  const size_t smem = ((size_t)parts * kp + 2 * (size_t)K4S_WARPS * kp) * 8;
  auto kern = bf16 ? k34_small_kernel<true> : k34_small_kernel<false>;
  RAG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(K4S_MAX_STAGE + 2 * K4S_WARPS * RAG_MAX_CANDIDATES * 8)));
  const uint32_t slices = (2 * kp + 1 + K4S_WARPS - 1) / K4S_WARPS;
  rag_p2p_view pv = rag_p2p_view();
  pv.nranks = 1;
  if (fuse && idx->nranks > 1) RAG_CHECK(comm_p2p_next(idx, B, k, &pv));  // the in-place K5 runs the exchange as well
  kern<<<dim3(B, slices), K4S_THREADS, smem, idx->stream>>>(idx->corpus, idx->ld, bt->d_q, bt->d_partial, parts, kp, k, eps,
                                                             key_has_qnorm, M, bt->d_k4s, bt->d_ticket, bt->d_cand, bt->d_local,
                                                             bt->d_local_cnt, fuse ? 1 : 0,
                                                             fuse ? k5_make_io(idx, fuse) : rag_k5::k5_io(), pv);
  RAG_CUDA(cudaGetLastError());
  idx->launches++;
  return RAG_OK;
}
