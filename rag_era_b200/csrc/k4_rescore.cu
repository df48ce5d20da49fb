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

namespace {
using namespace rag_exact;
constexpr int K4_WARP_BYTES = WARP_BYTES;

template <bool BF16>
__global__ void __launch_bounds__(RAG_MAX_CANDIDATES)
k4_rescore_kernel(const void* __restrict__ X, uint32_t ld, const float* __restrict__ Q,
                  const uint64_t* __restrict__ cand, uint32_t kp, uint32_t k, double eps,
                  int key_has_qnorm, uint64_t id_base, const uint8_t* __restrict__ ctype, const double* __restrict__ conf,
                  const int32_t* __restrict__ access, const int64_t* __restrict__ last_ms,
                  const uint64_t* __restrict__ row_keys, int64_t now_ms, double decay, double bonus,
                  rag_rec* __restrict__ local, uint32_t* __restrict__ local_cnt) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double s_score[RAG_MAX_CANDIDATES];
  __shared__ uint32_t s_row[RAG_MAX_CANDIDATES];
  __shared__ double s_nq;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t b = blockIdx.x;
  const uint32_t j = threadIdx.x;  // candidate index
  const uint64_t key = j < kp ? cand[(size_t)b * RAG_MAX_CANDIDATES + j] : 0ull;
  const bool valid = key != 0ull;
  // empty slots re-read the query's best candidate (a row this CTA fetches anyway) instead of all
  // hammering row 0 — thousands of CTAs on one L2 line set is a measurable hot spot
  const uint64_t key0 = cand[(size_t)b * RAG_MAX_CANDIDATES];
  const uint32_t row = valid ? rag_key_row(key) : (key0 != 0ull ? rag_key_row(key0) : 0u);

  unsigned char* wsm = smem + (size_t)warp * K4_WARP_BYTES;
  const chains c = warp_exact_sums<BF16>(X, ld, Q + (size_t)b * ld, row, wsm, lane);

  // similarity = dot / (norm(q) * norm(x)); NaN (zero norm) is defined as never selected
  double score = finish(c);
  const bool good = valid && score == score;
  s_score[j] = good ? score : -INFINITY;
  s_row[j] = row;
  if (j == 0) s_nq = c.nq;
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
  __shared__ double s_kth;
  if (good && rank + 1 == cnt) s_kth = score;
  __syncthreads();
  bool certified = true;
  const uint64_t last_key = cand[(size_t)b * RAG_MAX_CANDIDATES + kp - 1];
  if (last_key != 0ull && cnt > 0) {  // list full: rows outside the candidate set exist
    // K1/K2 keys hold dot/||x|| (||q|| cannot change the order); K1x keys hold the cosine itself
    const double t = key_has_qnorm ? (double)rag_key_score(last_key)
                                   : (double)rag_key_score(last_key) / sqrt(s_nq);
    certified = cnt == k && s_kth > t + eps;
  }
  const uint32_t flags = certified ? 0u : 1u;

  rag_rec* out = local + (size_t)b * k;
  if (good && rank < k) {
    rag_rec r;
    r.score = score;
    r.id = id_base + row;
    r.key = row_keys ? row_keys[row] : r.id;
    r.ctype = ctype ? ctype[row] : (uint32_t)RAG_CT_DOCUMENT;
    r.flags = flags;
    r.fresh = 0.0;
    r.conf_pad = 0.0;
    if (r.ctype == RAG_CT_MEMORY && conf && access && last_ms) {
      // calculateFreshnessScore — src/lib/memory/freshness.ts:43-55 (exp/log: <=1 ulp libm variance)
      const double hours = (double)(now_ms - last_ms[row]) / 3600000.0;
      const double dec = exp(__dmul_rn(-decay, hours));
      const double fb = __dmul_rn(log((double)access[row] + 1.0), bonus);
      const double sc = __dmul_rn(__dmul_rn(conf[row], dec), __dadd_rn(1.0, fb));
      r.fresh = fmax(0.0, fmin(1.0, sc));
    }
    out[rank] = r;
  }
  // empty tail + flags on every slot so the merge sees them even for empty shards
  for (uint32_t i = cnt + threadIdx.x; i < k; i += blockDim.x) {
    rag_rec e;
    e.score = -INFINITY; e.id = ~0ull; e.key = ~0ull; e.fresh = 0.0; e.ctype = 0; e.flags = flags; e.conf_pad = 0.0;
    out[i] = e;
  }
  if (threadIdx.x == 0) local_cnt[b] = cnt;
}

}  // namespace

int k4_launch(rag_index* idx, uint32_t B, uint32_t kp, uint32_t k, double eps, int key_has_qnorm,
              int64_t now_ms, double decay, double bonus) {
  rag_prof_scope ps(idx, RAG_PROF_RESCORE);
  const uint32_t nw = (kp + 31) / 32;
  const size_t smem = (size_t)nw * K4_WARP_BYTES;
  const bool bf16 = idx->desc.dtype == RAG_BF16;
  auto kern = bf16 ? k4_rescore_kernel<true> : k4_rescore_kernel<false>;
  RAG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * K4_WARP_BYTES)));
  kern<<<B, nw * 32, smem, idx->stream>>>(idx->corpus, idx->ld, idx->cur->d_q, idx->cur->d_cand, kp, k, eps,
                                           key_has_qnorm,
                                           idx->desc.id_base, idx->ctype, idx->conf, idx->access,
                                           idx->last_ms, idx->row_keys, now_ms, decay, bonus,
                                           idx->cur->d_local, idx->cur->d_local_cnt);
  RAG_CUDA(cudaGetLastError());
  idx->launches++;
  return RAG_OK;
}
