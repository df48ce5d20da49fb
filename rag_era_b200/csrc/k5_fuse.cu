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
#include "k5_body.cuh"

#include <algorithm>

namespace {
using namespace rag_k5;

// One warp per query; the grid is capped (a CTA walks queries b, b + gridDim.x, ...) so that every CTA is resident:
// with the exchange inside, CTA b of one rank waits for CTA b of its peers, and both ranks walk the queries in the
// same order, so the lowest unfinished query of every rank is always being served.
__global__ void __launch_bounds__(32)
k5_fuse_kernel(const rag_rec* __restrict__ recs, rag_p2p_view pv, k5_io io) {
  __shared__ fuse_smem s;
  const int lane = threadIdx.x;
  for (uint32_t b = blockIdx.x; b < io.a.B; b += gridDim.x) {
    // sharded: exchange the ranks' lists through the peer mailboxes
    const rag_rec* r = pv.nranks > 1 ? p2p_exchange(pv, recs, io.a.B, b, io.a.k, lane) : recs;
    k5_fuse_body(s, r, io, b, lane);
    __syncwarp();
  }
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
  const uint32_t nv = min(vc[b], (uint32_t)RAG_MAX_TOPK), nk = min(kwc[b], (uint32_t)RAG_MAX_KEYWORDS);
  for (uint32_t i = lane; i < nv; i += 32) {
    s.v_key[i] = vk[(size_t)b * vstride + i];
    s.v_ct[i] = vct ? vct[(size_t)b * vstride + i] : (uint8_t)RAG_CT_DOCUMENT;
  }
  for (uint32_t i = lane; i < nk; i += 32) s.k_key[i] = kw[(size_t)b * kstride + i];
  __syncwarp();
  const uint32_t n = rrf_fuse_lists(s, nv, nk, 0, cfg.vector_weight, cfg.keyword_weight, 0.0, cfg.k, cfg.both_bonus, lane);
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

rag_k5::k5_io k5_make_io(const rag_index* idx, const rag_fuse_args* a) {
  const rag_batch* bt = idx->cur;
  return rag_k5::k5_io{*a, bt->d_kw, bt->d_kwc, bt->d_out_keys, bt->d_out_scores, bt->d_out_src, bt->d_out_ct, bt->d_out_cnt,
                       bt->d_out_rrf, bt->d_vec_ids, bt->d_vec_scores, bt->d_vec_cnt, bt->d_cert, bt->d_aux0, bt->d_aux1,
                       idx->d_counters};
}

int k5_launch(rag_index* idx, const rag_fuse_args* a) {
  rag_prof_scope ps(idx, RAG_PROF_FUSE);
  rag_p2p_view pv;
  RAG_CHECK(comm_p2p_next(idx, a->B, a->k, &pv));  // nranks 1 unless the peer-to-peer exchange is active
  // peer-to-peer: K5 reads this rank's records and gathers them itself; NCCL fallback: already gathered
  const rag_rec* recs = (a->nranks > 1 && pv.nranks <= 1) ? idx->cur->d_gather : idx->cur->d_local;
  const uint32_t grid = std::min<uint32_t>(a->B, (uint32_t)idx->sm_count * 16u);
  k5_fuse_kernel<<<grid, 32, 0, idx->stream>>>(recs, pv, k5_make_io(idx, a));
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
