// k5_body.cuh — the per-query body of K5 (merge of the ranks' exact top-k lists, min-cosine filter,
// Reciprocal Rank Fusion / MemoryStore blend / vector-only branch), shared by k5_fuse_kernel and by the
// small-batch K3+K4 kernel, whose last CTA of a query runs it in place on a single GPU (one launch less on
// the batch-1 latency path). One warp per query. Reference lines are cited in k5_fuse.cu.
#pragma once
#include "common.cuh"

namespace rag_k5 {

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
  uint64_t k_key[RAG_MAX_KEYWORDS];  // the query's keyword list, staged with one coalesced load
  // the lists concatenated in pass order (vector, keyword, freshness): key and RRF contribution of every occurrence
  uint64_t c_key[MAXE];
  double c_val[MAXE];
};

// reciprocalRankFusion (src/lib/hybrid-search.ts:129-208) over the vector list (s.v_key / s.v_ct, nv entries), the keyword
// list (s.k_key, nk) and the optional freshness list (s.f_key, nf), by one warp. The reference walks the lists one key at a
// time through an insertion-ordered Map; what it computes per key depends only on that key's OWN occurrences, in pass order:
//   first occurrence      -> new entry {score: w/(k+rank+1), source: its list}                              (:156-166, :178-187)
//   later, vector list    -> score += rrf; source = 'both'                                                  (:152-155)
//   later, other lists    -> score += rrf + bothBonus*score (the score BEFORE this add); source = 'both'    (:173-177)
// and the Map's order is the order of first occurrences. So: every occurrence's contribution is computed side by side (one
// fp64 division each), every lane takes one element of the concatenation, finds out whether it is its key's first occurrence
// and, if so, replays that key's occurrences in order with exactly the reference's operations; the entry index is the number
// of first occurrences before it. ~10 instructions per element of the concatenation instead of a ~130-cycle dependent step
// per key with one lane mutating the map. Returns the number of entries; fills s.e_key / e_score / e_src / e_ct.
__device__ __forceinline__ uint32_t rrf_fuse_lists(fuse_smem& s, uint32_t nv, uint32_t nk, uint32_t nf, double w_vec, double w_kw,
                                                   double w_fresh, double kconst, double bonus, int lane) {
  const uint32_t T = nv + nk + nf;
  for (uint32_t j = lane; j < T; j += 32) {
    const uint32_t list = j < nv ? 0u : (j < nv + nk ? 1u : 2u);
    const uint32_t r = list == 0 ? j : (list == 1 ? j - nv : j - nv - nk);   // rank within its list
    const double w = list == 0 ? w_vec : (list == 1 ? w_kw : w_fresh);
    s.c_key[j] = list == 0 ? s.v_key[r] : (list == 1 ? s.k_key[r] : s.f_key[r]);
    s.c_val[j] = __ddiv_rn(w, __dadd_rn(__dadd_rn(kconst, (double)r), 1.0));
  }
  __syncwarp();
  uint32_t n = 0;
  for (uint32_t base = 0; base < T; base += 32) {
    const uint32_t e = base + lane;
    const bool live = e < T;
    const uint64_t key = live ? s.c_key[e] : 0ull;
    bool first = live, both = false;
    double score = 0.0;
    for (uint32_t j = 0; j < T; j++) {
      const uint64_t kj = s.c_key[j];   // uniform address: one broadcast load
      const double cj = s.c_val[j];
      if (live && kj == key) {
        if (j < e) first = false;
        else if (j == e) score = cj;
        else {
          score = j < nv ? __dadd_rn(score, cj) : __dadd_rn(score, __dadd_rn(cj, __dmul_rn(bonus, score)));
          both = true;
        }
      }
    }
    const unsigned fm = __ballot_sync(0xFFFFFFFFu, first);
    if (first) {
      const uint32_t at = n + __popc(fm & ((1u << lane) - 1u));
      const uint32_t list = e < nv ? 0u : (e < nv + nk ? 1u : 2u);
      s.e_key[at] = key;
      s.e_score[at] = score;
      s.e_src[at] = both ? (uint8_t)RAG_SRC_BOTH : (list == 0 ? (uint8_t)RAG_SRC_VECTOR : (list == 1 ? (uint8_t)RAG_SRC_KEYWORD : (uint8_t)RAG_SRC_FRESHNESS));
      s.e_ct[at] = list == 0 ? s.v_ct[e] : (list == 1 ? (uint8_t)RAG_CT_DOCUMENT : (uint8_t)RAG_CT_MEMORY);
    }
    n += __popc(fm);
  }
  __syncwarp();
  return n;
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
// SM's L1 knows nothing about. LOCAL: the records sit in this CTA's shared memory (the fused single-GPU tail).
template <bool LOCAL>
__device__ __forceinline__ rag_rec rec_load(const rag_rec* r) {
  union { rag_rec rec; uint4 q[3]; } u;
  const uint4* p = reinterpret_cast<const uint4*>(r);
  if (LOCAL) {
    u.q[0] = p[0]; u.q[1] = p[1]; u.q[2] = p[2];
  } else {
    u.q[0] = __ldcg(p);
    u.q[1] = __ldcg(p + 1);
    u.q[2] = __ldcg(p + 2);
  }
  return u.rec;
}

// C1 fused into K5 (see comm.cu): store this rank's k records of query b into every rank's mailbox, raise the
// flags, wait for every rank's flag. Returns the base of the gathered records [rank][B][k] (local mailbox).
// One full warp. Lane g raises the flag in rank g's mailbox and waits for rank g's flag in its own. A peer that
// does not arrive within pv.timeout_cycles is given up on: the host-mapped status word is raised (the call then
// fails with RAG_ERR_TIMEOUT) and the merge runs on whatever the mailbox holds — no trap, no hang.
__device__ __forceinline__ const rag_rec* p2p_exchange(const rag_p2p_view& pv, const rag_rec* local, uint32_t B, uint32_t b,
                                                       uint32_t k, int lane) {
  const uint64_t half = (uint64_t)(pv.step & 1u) * pv.half_bytes;
  const uint4* src = reinterpret_cast<const uint4*>(local + (size_t)b * k);
  const uint32_t n16 = k * (uint32_t)(sizeof(rag_rec) / 16);
  // the records were written by other warps / CTAs of this kernel (or an earlier one): read them through L2
  for (uint32_t i = lane; i < n16; i += 32) {
    const uint4 v = __ldcg(src + i);
    for (uint32_t g = 0; g < pv.nranks; g++)
      (reinterpret_cast<uint4*>(pv.base[g] + half) + ((size_t)pv.rank * B + b) * n16)[i] = v;
  }
  __threadfence_system();  // this lane's stores are visible system-wide before the flags go up
  __syncwarp();
  if ((uint32_t)lane < pv.nranks) {
    uint32_t* f = reinterpret_cast<uint32_t*>(pv.base[lane] + half + pv.flags_off) + (size_t)pv.rank * pv.flag_stride + b;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(pv.flag) : "memory");
    const uint32_t* w = reinterpret_cast<const uint32_t*>(pv.base[pv.rank] + half + pv.flags_off) + (size_t)lane * pv.flag_stride + b;
    const long long t0 = clock64();
    uint32_t seen;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(w) : "memory");
      if (seen == pv.flag) break;
      if (clock64() - t0 > pv.timeout_cycles) {
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(pv.status), "r"(1u) : "memory");
        break;
      }
    }
  }
  __syncwarp();
  return reinterpret_cast<const rag_rec*>(pv.base[pv.rank] + half);
}

// everything K5 reads and writes besides the records
struct k5_io {
  rag_fuse_args a;
  const uint64_t* kw;
  const uint32_t* kwc;
  uint64_t* o_key;
  double* o_score;
  uint8_t* o_src;
  uint8_t* o_ct;
  uint32_t* o_cnt;
  uint8_t* o_rrf;
  uint64_t* v_ids;
  double* v_scores;
  uint32_t* v_cnt;
  uint8_t* o_cert;
  double* o_aux0;
  double* o_aux1;
  unsigned long long* counters;  // [2] device-side totals since the last rag_certified_totals: certified, queries
};

// recs: [nranks][B][k] records; executed by one full warp for query b. LOCAL: `recs` points at this query's k records in
// shared memory (nranks == 1, the fused tail of the small-batch kernel) instead of the global [nranks][B][k] array.
template <bool LOCAL = false>
__device__ __forceinline__ void k5_fuse_body(fuse_smem& s, const rag_rec* recs, const k5_io io, uint32_t b, int lane) {
  const rag_fuse_args a = io.a;
  const uint32_t k = a.k;
  const uint64_t* kw = io.kw;
  const uint32_t* kwc = io.kwc;
  uint64_t* o_key = io.o_key;
  double* o_score = io.o_score;
  uint8_t* o_src = io.o_src;
  uint8_t* o_ct = io.o_ct;
  uint32_t* o_cnt = io.o_cnt;
  uint8_t* o_rrf = io.o_rrf;
  uint64_t* v_ids = io.v_ids;
  double* v_scores = io.v_scores;
  uint32_t* v_cnt = io.v_cnt;
  uint8_t* o_cert = io.o_cert;
  double* o_aux0 = io.o_aux0;
  double* o_aux1 = io.o_aux1;

  // ---- 1. gather the ranks' exact top-k lists and merge on (score desc, id asc) -------
  const uint32_t m = a.nranks * k;
  uint32_t uncert = 0;
  for (uint32_t i = lane; i < m; i += 32) {
    const uint32_t g = i / k, slot = i % k;
    const rag_rec r = rec_load<LOCAL>(LOCAL ? recs + slot : recs + ((size_t)g * a.B + b) * k + slot);
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
      const rag_rec r = rec_load<LOCAL>(LOCAL ? recs + slot : recs + ((size_t)g * a.B + b) * k + slot);
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
  if (lane == 0 && io.counters) {
    atomicAdd(io.counters + 1, 1ull);
    if (!uncert) atomicAdd(io.counters, 1ull);
  }

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

  const uint32_t nk = (a.mode == 0 && kwc) ? min(kwc[b], (uint32_t)RAG_MAX_KEYWORDS) : 0u;
  // the keyword keys in one coalesced load (the fusion pass walks them one by one: from global memory every step
  // would be a dependent L2 round trip)
  for (uint32_t i = lane; i < nk; i += 32) s.k_key[i] = kw[(size_t)b * a.kw_stride + i];
  __syncwarp();
  // ---- 3b. vector-only branch (hybrid-search.ts:346-354) -----------------------------
  if (nk == 0) {
    for (uint32_t i = lane; i < nv; i += 32) {
      ok[i] = s.v_id[i]; os[i] = s.v_score[i]; osrc[i] = RAG_SRC_VECTOR; oct[i] = s.v_ct[i];
    }
    if (lane == 0) { o_cnt[b] = nv; o_rrf[b] = 0; }
    return;
  }

  // ---- 3c. reciprocalRankFusion (hybrid-search.ts:129-208) ---------------------------
  uint32_t nf = 0;
  if (a.fresh_limit > 0) {
    // north-star extension (SURVEY N-c4 ii): memory hits of the vector stage ranked by
    // freshness desc (ties → lower chunk id), fused like a keyword list
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
  }
  const uint32_t n = rrf_fuse_lists(s, nv, nk, nf, a.rrf.vector_weight, a.rrf.keyword_weight, a.fresh_weight, a.rrf.k, a.rrf.both_bonus, lane);
  emit_sorted(s, n, ok, os, osrc, oct, lane);
  if (lane == 0) { o_cnt[b] = n; o_rrf[b] = 1; }
}

}  // namespace rag_k5

// host: the k5_io of the current batch (k5_fuse.cu)
rag_k5::k5_io k5_make_io(const rag_index* idx, const rag_fuse_args* a);
