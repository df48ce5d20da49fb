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

namespace {
using namespace rag_k5;

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
k5_fuse_kernel(const rag_rec* __restrict__ recs, rag_p2p_view pv, k5_io io) {
  __shared__ fuse_smem s;
  const int lane = threadIdx.x;
  const uint32_t b = blockIdx.x;
  // sharded: exchange the ranks' lists through the peer mailboxes
  if (pv.nranks > 1) recs = p2p_exchange(pv, recs, io.a.B, b, io.a.k, lane);
  k5_fuse_body(s, recs, io, b, lane);
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

rag_k5::k5_io k5_make_io(const rag_index* idx, const rag_fuse_args* a) {
  const rag_batch* bt = idx->cur;
  return rag_k5::k5_io{*a, bt->d_kw, bt->d_kwc, bt->d_out_keys, bt->d_out_scores, bt->d_out_src, bt->d_out_ct, bt->d_out_cnt,
                       bt->d_out_rrf, bt->d_vec_ids, bt->d_vec_scores, bt->d_vec_cnt, bt->d_cert, bt->d_aux0, bt->d_aux1};
}

int k5_launch(rag_index* idx, const rag_fuse_args* a) {
  rag_prof_scope ps(idx, RAG_PROF_FUSE);
  rag_p2p_view pv;
  RAG_CHECK(comm_p2p_next(idx, a->B, a->k, &pv));  // nranks 1 unless the peer-to-peer exchange is active
  // peer-to-peer: K5 reads this rank's records and gathers them itself; NCCL fallback: already gathered
  const rag_rec* recs = (a->nranks > 1 && pv.nranks <= 1) ? idx->cur->d_gather : idx->cur->d_local;
  k5_fuse_kernel<<<a->B, 32, 0, idx->stream>>>(recs, pv, k5_make_io(idx, a));
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
