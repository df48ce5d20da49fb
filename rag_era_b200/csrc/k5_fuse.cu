// k5_fuse.cu — K5: final merge of the ranks' exact top-k lists, min-cosine filter,
// Reciprocal Rank Fusion with the keyword list (+ optional freshness list), and the
// MemoryStore blend. One warp per query; every list here has <= 64 entries, so this is
// a latency-bound micro-kernel whose purpose is to keep the batched pipeline on the
// device (no D2H -> JS -> H2D hop between top-k and fusion).
//
// Reference semantics restated (all fp64, round-to-nearest, NO fma — nvcc would
// otherwise contract a*b+c, so every operation is an explicit _rn intrinsic):
//   filter   src/lib/hybrid-search.ts:308-314   drop score < minVectorScore, order kept
//   RRF      src/lib/hybrid-search.ts:129-208   Map in insertion order, vector pass then
//            keyword pass, `existing.score += rrf + bothBonus*existing.score`, stable sort
//   branch   src/lib/hybrid-search.ts:333,346-354  no keyword hits → filtered vector list,
//            raw cosine scores, source 'vector'
//   memory   src/lib/memory/store.ts:119-175    memory rows only, cos >= minRelevance,
//            cos*0.7 + fresh*0.3, stable sort desc, slice(limit)
#include "common.cuh"

namespace {

constexpr int MAXM = 8 * RAG_MAX_TOPK;                                   // merge inputs (8 ranks)
constexpr int MAXE = RAG_MAX_TOPK + RAG_MAX_KEYWORDS + RAG_MAX_FRESH;    // fused entries

struct fuse_smem {
  double m_score[MAXM];
  uint64_t m_id[MAXM];
  uint16_t m_src[MAXM];   // rank*k + slot
  // vector stage in rank order
  double v_score[RAG_MAX_TOPK];
  uint64_t v_id[RAG_MAX_TOPK];
  uint64_t v_key[RAG_MAX_TOPK];
  double v_fresh[RAG_MAX_TOPK];
  uint8_t v_ct[RAG_MAX_TOPK];
  // fusion map (insertion order)
  uint64_t e_key[MAXE];
  double e_score[MAXE];
  uint8_t e_src[MAXE];
  uint8_t e_ct[MAXE];
  uint64_t f_key[RAG_MAX_FRESH];
};

// one sequential RRF pass over `n` keys; all lanes execute, lane 0 mutates the map.
// first_pass: vector pass semantics (:147-166), else keyword pass semantics (:169-188).
__device__ __forceinline__ void rrf_pass(fuse_smem& s, uint32_t& n_entries, const uint64_t* keys,
                                         const uint8_t* cts, uint32_t n, double weight, double kconst,
                                         double bonus, bool first_pass, uint8_t new_src, uint8_t new_ct,
                                         int lane) {
  for (uint32_t r = 0; r < n; r++) {
    const uint64_t key = keys[r];
    const double rrf = __ddiv_rn(weight, __dadd_rn(__dadd_rn(kconst, (double)r), 1.0));
    int found = -1;
    for (uint32_t base = 0; base < n_entries; base += 32) {
      const uint32_t i = base + lane;
      const unsigned hit = __ballot_sync(0xFFFFFFFFu, i < n_entries && s.e_key[i] == key);
      if (hit) { found = (int)base + __ffs(hit) - 1; break; }
    }
    if (lane == 0) {
      if (found >= 0) {
        const double e = s.e_score[found];
        s.e_score[found] = first_pass ? __dadd_rn(e, rrf)
                                      : __dadd_rn(e, __dadd_rn(rrf, __dmul_rn(bonus, e)));
        s.e_src[found] = RAG_SRC_BOTH;
      } else {
        s.e_key[n_entries] = key;
        s.e_score[n_entries] = rrf;
        s.e_src[n_entries] = new_src;
        s.e_ct[n_entries] = cts ? cts[r] : new_ct;
      }
    }
    if (found < 0) n_entries++;
    __syncwarp();
  }
}

// stable sort by score desc over the map (insertion index breaks ties) and emit
__device__ __forceinline__ void emit_sorted(const fuse_smem& s, uint32_t n, uint64_t* o_key, double* o_score,
                                            uint8_t* o_src, uint8_t* o_ct, int lane) {
  for (uint32_t i = lane; i < n; i += 32) {
    const double si = s.e_score[i];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < n; j++) {
      const double sj = s.e_score[j];
      rank += (sj > si || (sj == si && j < i)) ? 1u : 0u;
    }
    o_key[rank] = s.e_key[i]; o_score[rank] = si; o_src[rank] = s.e_src[i]; o_ct[rank] = s.e_ct[i];
  }
}

// one record through L2 (ld.global.cg): mailbox records are written by PEER GPUs over NVLink, which this
// SM's L1 knows nothing about
__device__ __forceinline__ rag_rec rec_load(const rag_rec* r) {
  union { rag_rec rec; uint4 q[3]; } u;
  const uint4* p = reinterpret_cast<const uint4*>(r);
  u.q[0] = __ldcg(p);
  u.q[1] = __ldcg(p + 1);
  u.q[2] = __ldcg(p + 2);
  return u.rec;
}

// C1 fused into K5 (see comm.cu): store this rank's k records of query b into every rank's mailbox, raise the
// flags, wait for every rank's flag. Returns the base of the gathered records [rank][B][k] (local mailbox).
__device__ __forceinline__ const rag_rec* p2p_exchange(const rag_p2p_view& pv, const rag_rec* local, uint32_t B, uint32_t b,
                                                       uint32_t k, int lane) {
  const uint64_t half = (uint64_t)(pv.step & 1u) * pv.half_bytes;
  const uint4* src = reinterpret_cast<const uint4*>(local + (size_t)b * k);
  const uint32_t n16 = k * (uint32_t)(sizeof(rag_rec) / 16);
  for (uint32_t g = 0; g < pv.nranks; g++) {
    uint4* dst = reinterpret_cast<uint4*>(pv.base[g] + half) + ((size_t)pv.rank * B + b) * k * (sizeof(rag_rec) / 16);
    for (uint32_t i = lane; i < n16; i += 32) dst[i] = src[i];
  }
  __threadfence_system();  // this lane's stores are visible system-wide before the flags go up
  __syncwarp();
  if ((uint32_t)lane < pv.nranks) {
    uint32_t* f = reinterpret_cast<uint32_t*>(pv.base[lane] + half + pv.flags_off) + (size_t)pv.rank * pv.flag_stride + b;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(pv.step) : "memory");
    const uint32_t* w = reinterpret_cast<const uint32_t*>(pv.base[pv.rank] + half + pv.flags_off) + (size_t)lane * pv.flag_stride + b;
    const long long t0 = clock64();
    uint32_t seen;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(w) : "memory");
      if (seen != pv.step && clock64() - t0 > 40000000000ll) __trap();  // ~20 s: a peer never arrived
    } while (seen != pv.step);
  }
  __syncwarp();
  return reinterpret_cast<const rag_rec*>(pv.base[pv.rank] + half);
}

__global__ void __launch_bounds__(32)
k5_fuse_kernel(const rag_rec* __restrict__ recs, rag_p2p_view pv, rag_fuse_args a, const uint64_t* __restrict__ kw,
               const uint32_t* __restrict__ kwc, uint64_t* __restrict__ o_key, double* __restrict__ o_score,
               uint8_t* __restrict__ o_src, uint8_t* __restrict__ o_ct, uint32_t* __restrict__ o_cnt,
               uint8_t* __restrict__ o_rrf, uint64_t* __restrict__ v_ids, double* __restrict__ v_scores,
               uint32_t* __restrict__ v_cnt, uint8_t* __restrict__ o_cert, double* __restrict__ o_aux0,
               double* __restrict__ o_aux1) {
  __shared__ fuse_smem s;
  const int lane = threadIdx.x;
  const uint32_t b = blockIdx.x;
  const uint32_t k = a.k;

  // ---- 0. sharded: exchange the ranks' lists through the peer mailboxes -----------------
  if (pv.nranks > 1) recs = p2p_exchange(pv, recs, a.B, b, k, lane);

  // ---- 1. gather the ranks' exact top-k lists and merge on (score desc, id asc) -------
  const uint32_t m = a.nranks * k;
  uint32_t uncert = 0;
  for (uint32_t i = lane; i < m; i += 32) {
    const uint32_t g = i / k, slot = i % k;
    const rag_rec r = rec_load(recs + ((size_t)g * a.B + b) * k + slot);
    s.m_score[i] = r.score; s.m_id[i] = r.id; s.m_src[i] = (uint16_t)i;
    if (slot == 0) uncert |= r.flags & 1u;
  }
  uncert = __any_sync(0xFFFFFFFFu, uncert != 0);
  __syncwarp();
  uint32_t n_top = 0;
  for (uint32_t i = lane; i < m; i += 32) {
    const double si = s.m_score[i];
    const uint64_t ii = s.m_id[i];
    if (si == -INFINITY) continue;
    uint32_t rank = 0;
    for (uint32_t j = 0; j < m; j++) {
      const double sj = s.m_score[j];
      rank += (sj != -INFINITY && (sj > si || (sj == si && s.m_id[j] < ii))) ? 1u : 0u;
    }
    if (rank < k) {
      const uint32_t g = i / k, slot = i % k;
      const rag_rec r = rec_load(recs + ((size_t)g * a.B + b) * k + slot);
      s.v_score[rank] = si; s.v_id[rank] = ii; s.v_key[rank] = r.key;
      s.v_fresh[rank] = r.fresh; s.v_ct[rank] = (uint8_t)r.ctype;
      n_top++;
    }
  }
  n_top = __reduce_add_sync(0xFFFFFFFFu, n_top);
  __syncwarp();

  // ---- 2. min-cosine filter (hybrid-search.ts:308-314); list is sorted so survivors are a prefix
  uint32_t nv = n_top;
  if (a.mode == 0) {
    uint32_t keep = 0;
    for (uint32_t i = lane; i < n_top; i += 32) keep += (s.v_score[i] < a.min_score) ? 0u : 1u;
    nv = __reduce_add_sync(0xFFFFFFFFu, keep);
  }
  if (v_ids) {
    for (uint32_t i = lane; i < k; i += 32) {
      v_ids[(size_t)b * k + i] = i < nv ? s.v_id[i] : ~0ull;
      v_scores[(size_t)b * k + i] = i < nv ? s.v_score[i] : -INFINITY;
    }
    if (lane == 0) v_cnt[b] = nv;
  }
  if (lane == 0 && o_cert) o_cert[b] = uncert ? 0 : 1;

  uint64_t* ok = o_key + (size_t)b * a.out_cap;
  double* os = o_score + (size_t)b * a.out_cap;
  uint8_t* osrc = o_src + (size_t)b * a.out_cap;
  uint8_t* oct = o_ct + (size_t)b * a.out_cap;

  // ---- 3a. MemoryStore.retrieve blend (store.ts:119-175) -----------------------------
  if (a.mode == 1) {
    uint32_t n = 0;  // map reused: e_key = id, e_score = blended; insertion order = retriever rank
    for (uint32_t i = 0; i < nv; i++) {
      const bool take = s.v_ct[i] == RAG_CT_MEMORY && !(s.v_score[i] < a.mem_min_relevance);
      if (take) {
        if (lane == 0) {
          s.e_key[n] = s.v_id[i];
          s.e_score[n] = __dadd_rn(__dmul_rn(s.v_score[i], 0.7), __dmul_rn(s.v_fresh[i], 0.3));
          s.e_src[n] = (uint8_t)i; s.e_ct[n] = RAG_CT_MEMORY;
        }
        n++;
      }
    }
    __syncwarp();
    for (uint32_t i = lane; i < n; i += 32) {
      const double si = s.e_score[i];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < n; j++) {
        const double sj = s.e_score[j];
        rank += (sj > si || (sj == si && j < i)) ? 1u : 0u;
      }
      if (rank < a.mem_limit) {
        const uint32_t src = s.e_src[i];
        ok[rank] = s.e_key[i]; os[rank] = si; osrc[rank] = RAG_SRC_VECTOR; oct[rank] = RAG_CT_MEMORY;
        o_aux0[(size_t)b * a.out_cap + rank] = s.v_score[src];
        o_aux1[(size_t)b * a.out_cap + rank] = s.v_fresh[src];
      }
    }
    if (lane == 0) { o_cnt[b] = n < a.mem_limit ? n : a.mem_limit; o_rrf[b] = 0; }
    return;
  }

  const uint32_t nk = (a.mode == 0 && kwc) ? kwc[b] : 0u;
  // ---- 3b. vector-only branch (hybrid-search.ts:346-354) -----------------------------
  if (nk == 0) {
    for (uint32_t i = lane; i < nv; i += 32) {
      ok[i] = s.v_id[i]; os[i] = s.v_score[i]; osrc[i] = RAG_SRC_VECTOR; oct[i] = s.v_ct[i];
    }
    if (lane == 0) { o_cnt[b] = nv; o_rrf[b] = 0; }
    return;
  }

  // ---- 3c. reciprocalRankFusion (hybrid-search.ts:129-208) ---------------------------
  uint32_t n = 0;
  rrf_pass(s, n, s.v_key, s.v_ct, nv, a.rrf.vector_weight, a.rrf.k, a.rrf.both_bonus, true,
           RAG_SRC_VECTOR, RAG_CT_DOCUMENT, lane);
  rrf_pass(s, n, kw + (size_t)b * a.kw_stride, nullptr, nk, a.rrf.keyword_weight, a.rrf.k, a.rrf.both_bonus,
           false, RAG_SRC_KEYWORD, RAG_CT_DOCUMENT, lane);
  if (a.fresh_limit > 0) {
    // north-star extension (SURVEY N-c4 ii): memory hits of the vector stage ranked by
    // freshness desc (ties → lower chunk id), fused like a keyword list
    uint32_t nf = 0;
    for (uint32_t i = lane; i < nv; i += 32) {
      if (s.v_ct[i] != RAG_CT_MEMORY) continue;
      uint32_t rank = 0;
      for (uint32_t j = 0; j < nv; j++) {
        if (s.v_ct[j] != RAG_CT_MEMORY) continue;
        rank += (s.v_fresh[j] > s.v_fresh[i] || (s.v_fresh[j] == s.v_fresh[i] && s.v_id[j] < s.v_id[i])) ? 1u : 0u;
      }
      if (rank < a.fresh_limit) { s.f_key[rank] = s.v_key[i]; nf++; }
    }
    nf = __reduce_add_sync(0xFFFFFFFFu, nf);
    __syncwarp();
    rrf_pass(s, n, s.f_key, nullptr, nf, a.fresh_weight, a.rrf.k, a.rrf.both_bonus, false,
             RAG_SRC_FRESHNESS, RAG_CT_MEMORY, lane);
  }
  emit_sorted(s, n, ok, os, osrc, oct, lane);
  if (lane == 0) { o_cnt[b] = n; o_rrf[b] = 1; }
}

// fusion only (rag_rrf_fuse): lists given as keys
__global__ void __launch_bounds__(32)
k5_rrf_only_kernel(rag_rrf_config cfg, const uint64_t* __restrict__ vk, const uint8_t* __restrict__ vct,
                   const uint32_t* __restrict__ vc, uint32_t vstride, const uint64_t* __restrict__ kw,
                   const uint32_t* __restrict__ kwc, uint32_t kstride, uint32_t out_cap,
                   uint64_t* __restrict__ o_key, double* __restrict__ o_score, uint8_t* __restrict__ o_src,
                   uint8_t* __restrict__ o_ct, uint32_t* __restrict__ o_cnt) {
  __shared__ fuse_smem s;
  const int lane = threadIdx.x;
  const uint32_t b = blockIdx.x;
  uint32_t n = 0;
  rrf_pass(s, n, vk + (size_t)b * vstride, vct ? vct + (size_t)b * vstride : nullptr, vc[b], cfg.vector_weight,
           cfg.k, cfg.both_bonus, true, RAG_SRC_VECTOR, RAG_CT_DOCUMENT, lane);
  rrf_pass(s, n, kw + (size_t)b * kstride, nullptr, kwc[b], cfg.keyword_weight, cfg.k, cfg.both_bonus, false,
           RAG_SRC_KEYWORD, RAG_CT_DOCUMENT, lane);
  emit_sorted(s, n, o_key + (size_t)b * out_cap, o_score + (size_t)b * out_cap, o_src + (size_t)b * out_cap,
              o_ct + (size_t)b * out_cap, lane);
  if (lane == 0) o_cnt[b] = n;
}

// calculateFreshnessScore over n rows (src/lib/memory/freshness.ts:43-55)
__global__ void k5_freshness_kernel(uint64_t n, const double* __restrict__ conf, const int32_t* __restrict__ acc,
                                    const int64_t* __restrict__ last, int64_t now_ms, double decay, double bonus,
                                    double* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double hours = (double)(now_ms - last[i]) / 3600000.0;
  const double dec = exp(__dmul_rn(-decay, hours));
  const double fb = __dmul_rn(log((double)acc[i] + 1.0), bonus);
  const double sc = __dmul_rn(__dmul_rn(conf[i], dec), __dadd_rn(1.0, fb));
  out[i] = fmax(0.0, fmin(1.0, sc));
}

}  // namespace

int k5_launch(rag_index* idx, const rag_fuse_args* a) {
  rag_prof_scope ps(idx, RAG_PROF_FUSE);
  rag_p2p_view pv;
  RAG_CHECK(comm_p2p_next(idx, a->B, a->k, &pv));  // nranks 1 unless the peer-to-peer exchange is active
  // peer-to-peer: K5 reads this rank's records and gathers them itself; NCCL fallback: already gathered
  const rag_rec* recs = (a->nranks > 1 && pv.nranks <= 1) ? idx->cur->d_gather : idx->cur->d_local;
  k5_fuse_kernel<<<a->B, 32, 0, idx->stream>>>(recs, pv, *a, idx->cur->d_kw, idx->cur->d_kwc, idx->cur->d_out_keys,
                                                idx->cur->d_out_scores, idx->cur->d_out_src, idx->cur->d_out_ct, idx->cur->d_out_cnt,
                                                idx->cur->d_out_rrf, idx->cur->d_vec_ids, idx->cur->d_vec_scores, idx->cur->d_vec_cnt,
                                                idx->cur->d_cert, idx->cur->d_aux0, idx->cur->d_aux1);
  RAG_CUDA(cudaGetLastError());
  idx->launches++;
  return RAG_OK;
}

int k5_rrf_only_launch(rag_index* idx, uint32_t B, const rag_rrf_config* cfg, const uint64_t* d_vec_keys,
                       const uint8_t* d_vec_ct, const uint32_t* d_vec_cnt, uint32_t vec_stride,
                       const uint64_t* d_kw, const uint32_t* d_kwc, uint32_t kw_stride, uint32_t out_cap) {
  k5_rrf_only_kernel<<<B, 32, 0, idx->stream>>>(*cfg, d_vec_keys, d_vec_ct, d_vec_cnt, vec_stride, d_kw, d_kwc,
                                                 kw_stride, out_cap, idx->cur->d_out_keys, idx->cur->d_out_scores,
                                                 idx->cur->d_out_src, idx->cur->d_out_ct, idx->cur->d_out_cnt);
  RAG_CUDA(cudaGetLastError());
  idx->launches++;
  return RAG_OK;
}

int k5_freshness_launch(rag_index* idx, uint64_t n, const double* d_conf, const int32_t* d_acc,
                        const int64_t* d_last, int64_t now_ms, double decay, double bonus, double* d_out) {
  if (n == 0) return RAG_OK;
  const uint32_t threads = 256;
  k5_freshness_kernel<<<(uint32_t)((n + threads - 1) / threads), threads, 0, idx->stream>>>(
      n, d_conf, d_acc, d_last, now_ms, decay, bonus, d_out);
  RAG_CUDA(cudaGetLastError());
  idx->launches++;
  return RAG_OK;
}
