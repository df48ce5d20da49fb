// api.cu — the C ABI of libragera.so (include/ragera.h): handle lifetime, staging, the
// kernel pipeline  score+select (K1 | K2 | K1x) → merge (K3) → exact rescore (K4) →
// [exchange of the ranks' local top-k: inside K5 over peer mailboxes, or ncclAllGather] →
// merge/filter/fuse (K5), certification-driven escalation, and the
// measurement helpers. Host code only; every kernel lives in its own .cu.
//
// There is no CPU implementation of any step in this library: without an sm_100 device
// rag_index_create fails with RAG_ERR_NO_DEVICE and nothing else can be called.
#include "common.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <new>
#include <vector>

namespace {

thread_local char g_err[1024] = "";

constexpr uint32_t kProfSpans = 8192;
// RAG_PATH_AUTO (measured on 1M x 1536, tools/k2_small_probe.py): the tensor kernel is HBM-bound at small
// batches too, and a bf16 operand halves the bytes — it wins from 4 queries (0.61 vs 0.93 ms); fp32 rows
// read as tf32 win from 8 (1.3 vs 1.5 ms). Below that the stream kernels (K1 / K1m) keep fp32 selection.
constexpr uint32_t kTensorMinBatchBf16 = 4;
constexpr uint32_t kTensorMinBatchTf32 = 8;
constexpr uint64_t kShadowStreamMinBytes = 128ull << 20;  // AUTO streams the fp16 shadow for one query once the shadow outgrows L2

size_t elem_size(const rag_index* idx) { return idx->desc.dtype == RAG_BF16 ? 2 : 4; }

// grow a device buffer to at least `need` bytes (contents are not preserved)
template <typename T>
int grow_dev(T** p, size_t* cap, size_t need, bool zero) {
  if (need <= *cap && *p) return RAG_OK;
  if (*p) { RAG_CUDA(cudaFree(*p)); *p = nullptr; *cap = 0; }
  if (need == 0) need = 16;
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, need);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return rag_set_error(RAG_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", need, cudaGetErrorString(e));
  }
  if (zero) {
    // cudaMemset runs on the legacy default stream, which the library's non-blocking stream does NOT
    // wait for: finish it here, or a later async copy into this buffer can be overwritten by the zeros
    RAG_CUDA(cudaMemset(q, 0, need));
    RAG_CUDA(cudaDeviceSynchronize());
  }
  *p = (T*)q;
  *cap = need;
  return RAG_OK;
}

int grow_pinned(uint8_t** p, size_t* cap, size_t need) {
  if (need <= *cap && *p) return RAG_OK;
  if (*p) { RAG_CUDA(cudaFreeHost(*p)); *p = nullptr; *cap = 0; }
  if (need == 0) need = 16;
  void* q = nullptr;
  cudaError_t e = cudaHostAlloc(&q, need, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return rag_set_error(RAG_ERR_NOMEM, "cudaHostAlloc(%zu bytes) failed: %s", need, cudaGetErrorString(e));
  }
  *p = (uint8_t*)q;
  *cap = need;
  return RAG_OK;
}

void free_batch(rag_batch* b) {
  cudaFree(b->d_q); cudaFree(b->d_qb); cudaFree(b->d_rho_q); cudaFree(b->d_in); cudaFree(b->d_sel); cudaFree(b->d_partial);
  cudaFree(b->d_cand); cudaFree(b->d_local); cudaFree(b->d_gather); cudaFree(b->d_local_cnt); cudaFree(b->d_out);
  cudaFree(b->d_k4s); cudaFree(b->d_ticket); cudaFree(b->d_gin);
  if (b->h_in) cudaFreeHost(b->h_in);
  if (b->h_out) cudaFreeHost(b->h_out);
  if (b->h_q) cudaFreeHost(b->h_q);
  *b = rag_batch();
}

// ---- carved per-call layouts ------------------------------------------------------
struct out_layout {
  size_t keys, scores, src, ct, cnt, rrf, vids, vscores, vcnt, cert, aux0, aux1, total, total_no_aux;
};
out_layout layout_out(uint32_t B, uint32_t cap, uint32_t k) {
  out_layout L;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 15) & ~(size_t)15; return at; };
  L.cnt = take((size_t)B * 4);
  L.rrf = take(B);
  L.vcnt = take((size_t)B * 4);
  L.cert = take(B);
  L.keys = take((size_t)B * cap * 8);
  L.scores = take((size_t)B * cap * 8);
  L.src = take((size_t)B * cap);
  L.ct = take((size_t)B * cap);
  L.vids = take((size_t)B * k * 8);
  L.vscores = take((size_t)B * k * 8);
  L.total_no_aux = o;
  L.aux0 = take((size_t)B * cap * 8);
  L.aux1 = take((size_t)B * cap * 8);
  L.total = o;
  return L;
}
// carve the output block at `base`: the device block, or — small batches — its pinned host mirror, which the kernels
// can write directly (pinned allocations are device-accessible under UVA; a few hundred bytes of posted PCIe writes
// instead of a D2H copy node)
void bind_out_at(rag_batch* b, const out_layout& L, uint8_t* base) {
  b->d_out_cnt = (uint32_t*)(base + L.cnt);
  b->d_out_rrf = base + L.rrf;
  b->d_vec_cnt = (uint32_t*)(base + L.vcnt);
  b->d_cert = base + L.cert;
  b->d_out_keys = (uint64_t*)(base + L.keys);
  b->d_out_scores = (double*)(base + L.scores);
  b->d_out_src = base + L.src;
  b->d_out_ct = base + L.ct;
  b->d_vec_ids = (uint64_t*)(base + L.vids);
  b->d_vec_scores = (double*)(base + L.vscores);
  b->d_aux0 = (double*)(base + L.aux0);
  b->d_aux1 = (double*)(base + L.aux1);
}
void bind_out(rag_batch* b, const out_layout& L) { bind_out_at(b, L, b->d_out); }

// ---- plan: which kernel scores, how many candidates, what error bound ---------------
struct plan {
  int path;          // rag_path (never AUTO)
  uint32_t kp;       // K'
  double eps;        // selection-error bound used by K4's certification (the per-batch part)
  bool eps_per_query;  // tensor path, rigorous bound: eps[b] = eps + rho_q[b] * eps_q_mul
  double eps_q_mul;
  int key_has_qnorm; // K1x keys hold the cosine, K1/K2 keys hold dot/||x||
};

// host mirror of the rows' rounding residual (aux_build raises it on the device)
int refresh_rho_x(rag_index* idx) {
  if (!idx->rho_x_stale || !idx->d_rho_x) return RAG_OK;
  uint32_t bits = 0;
  RAG_CUDA(cudaMemcpyAsync(&bits, idx->d_rho_x, 4, cudaMemcpyDeviceToHost, idx->stream));
  RAG_CUDA(cudaStreamSynchronize(idx->stream));
  memcpy(&idx->rho_x, &bits, 4);
  idx->rho_x_stale = false;
  return RAG_OK;
}

int make_plan(rag_index* idx, uint32_t B, uint32_t k, uint32_t req_path, uint32_t slack, double eps, uint32_t flags,
              plan* p) {
  int path = (int)req_path;
  const bool has_f16_shadow = idx->desc.dtype == RAG_F32 && idx->shadow && idx->shadow_f16 && idx->inv_norm;
  if (path == RAG_PATH_AUTO) {
    static const bool shadow_stream_auto = !(getenv("RAGERA_SHADOW_STREAM") && atoi(getenv("RAGERA_SHADOW_STREAM")) == 0);
    path = (k2_available(idx) && B >= (idx->shadow ? kTensorMinBatchBf16 : kTensorMinBatchTf32)) ? RAG_PATH_TENSOR
                                                                                              : RAG_PATH_STREAM;
    // a single query on an index that carries the fp16 shadow: stream the shadow — half the bytes of the fp32 rows.
    // Only where bytes are what costs: the shadow pass keeps K' = 32 candidates (its bound is the rows' fp16 residual), and
    // rescoring 32 instead of 16 rows outweighs the bytes saved while the corpus sits in L2 (measured at C1, 10k rows:
    // p50 52.6 us on the shadow against 49.0 us on the fp32 rows; at 10M rows 4.45 against 8.52 ms).
    if (path == RAG_PATH_STREAM && B == 1 && has_f16_shadow && shadow_stream_auto && idx->rows * (uint64_t)idx->ld * 2 >= kShadowStreamMinBytes)
      path = RAG_PATH_SHADOW_STREAM;
  }
  if (path != RAG_PATH_STREAM && path != RAG_PATH_TENSOR && path != RAG_PATH_EXACT && path != RAG_PATH_SHADOW_STREAM)
    return rag_set_error(RAG_ERR_INVALID, "unknown rag_path %d", path);
  if (path == RAG_PATH_SHADOW_STREAM && !has_f16_shadow)
    return rag_set_error(RAG_ERR_UNSUPPORTED, "shadow stream path needs an fp32 index created with RAG_INDEX_F16_SHADOW");
  if (path == RAG_PATH_TENSOR && !k2_available(idx))
    return rag_set_error(RAG_ERR_UNSUPPORTED, "tensor path (K2) is not available for this index");
  const bool stat_eps = (flags & RAG_SEARCH_STAT_EPS) != 0;
  uint32_t s = slack;
  // the tensor path selects on 16-bit / tf32 scores: a wider window keeps (nearly) every query certifiable in one
  // pass — an uncertified query costs a whole extra corpus pass on the stream path. With bf16 operands (a bf16 corpus
  // or a bf16 shadow) the rigorous bound carries bf16's 8-bit rounding of the query (~1.7e-3 at D=1536) and, for a
  // shadow, of every row as well: widest window.
  const bool bf16_ops = idx->shadow && !idx->shadow_f16;
  if (s == 0) s = path == RAG_PATH_TENSOR ? ((!stat_eps && eps <= 0.0 && bf16_ops) ? 48u : std::max(22u, k))
                  : (path == RAG_PATH_SHADOW_STREAM ? std::max(22u, k) : 6u);
  uint32_t kp = std::min<uint32_t>(path == RAG_PATH_TENSOR ? 48u : (uint32_t)RAG_MAX_CANDIDATES, k + s);
  if (kp < k) return rag_set_error(RAG_ERR_UNSUPPORTED, "tensor path supports k <= 48 (k=%u)", k);
  p->path = path;
  p->kp = kp;
  // K1x keys hold the cosine; so do K2's on the 16-bit path (its query operand is q/||q||); K1 and K2-tf32 keys hold dot/||x||
  p->key_has_qnorm = path == RAG_PATH_EXACT || (path == RAG_PATH_TENSOR && idx->shadow != nullptr);
  p->eps_per_query = false;
  p->eps_q_mul = 0.0;
  const double u24 = 5.9604644775390625e-08;  // 2^-24
  if (eps > 0.0) p->eps = eps;
  else if (path == RAG_PATH_SHADOW_STREAM) {
    // K1 over x~ = fp16(x/||x||): key = fl32(q.x~), in cosine units q.x~/||q||. With u_q = q/||q||, e_x = x~ - x/||x||:
    //   u_q.x~ - cos = u_q.e_x  =>  |.| <= ||e_x|| <= rho_x  (Cauchy-Schwarz; rho_x measured when the rows were loaded)
    // + the fp32 accumulation of the dot product (the stream path's bound, relative to ||q|| ||x~|| <= ||q|| (1 + rho_x))
    RAG_CHECK(refresh_rho_x(idx));
    const double rx = (double)idx->rho_x * 1.001;
    p->eps = rx + 2.5 * ((double)idx->ld / 64.0 + 8.0) * u24 * (1.0 + rx);
  } else if (path == RAG_PATH_STREAM)
    // fp32: <= ld/64 + 8 roundings per sum (2 accumulators per lane, 5 shuffle levels), twice
    // (dot and norm), with a safety factor — 4.8e-6 at D = 1536
    p->eps = 2.5 * ((double)idx->ld / 64.0 + 8.0) * u24;
  else if (path == RAG_PATH_TENSOR && stat_eps)
    // STATISTICAL (round 1): measured per-score error sigma ~ 0.0022/sqrt(D) for bf16 operands (5.6e-5 at
    // D=1536, max 2.8e-4 over 5e5 pairs — tests/test_gpu_tensor.py); ~11 sigma, NOT a proof. tf32 is 4x finer.
    p->eps = (idx->shadow ? 0.024 : 0.006) / sqrt((double)idx->ld);
  else if (path == RAG_PATH_TENSOR) {
    // RIGOROUS (default). In cosine units K2's key is acc (* inv_x) with acc ~ q~.x~ (tensor core, fp32 accumulate)
    // over the operands q~, x~ of gen.cu's table. With u_q = q/||q||, u_x = x/||x||, e_q = q~ - u_q, e_x = x~ - u_x:
    //   q~.x~ - u_q.u_x = e_q.x~ + u_q.e_x  =>  |.| <= ||e_q|| ||x~|| + ||e_x||  <=  rho_q (1 + rho_x) + rho_x   (Cauchy-Schwarz)
    //   rho_q = ||e_q|| measured per query (q_operand_launch), rho_x = max over rows of ||e_x|| measured when the rows
    //   were loaded (aux_build; 0 for a bf16 corpus, whose rows are the operand).
    // + accumulation: every one of the ld products enters an fp32 sum once, each add rounds (or truncates) at most
    //   2^-23 relative to a partial sum that is <= sum|q~_i x~_i| <= ||q~|| ||x~||           => ld * 2^-23
    // + the two normalisations: fp32 sum of ld squares (ld/32 sequential fmas per lane + 5 shuffle adds) and rsqrtf
    //   (2 ulp), halved by the square root, once for the query and once for the row; + the fp32 roundings of the
    //   scaled element / of the key itself. The residuals are computed in fp32 and inflated.
    RAG_CHECK(refresh_rho_x(idx));
    const double rx = (double)idx->rho_x * 1.001;
    const double acc = (double)idx->ld * 2.0 * u24 * 1.02;
    const double nrm = 2.0 * (0.5 * ((double)idx->ld / 32.0 + 8.0) * u24 + 4.0 * u24) + 1.0e-6;
    p->eps = rx + acc + nrm;
    p->eps_per_query = true;
    p->eps_q_mul = (1.0 + rx) * 1.001;
  } else
    p->eps = 2.0e-7;  // one fp32 rounding of the exact cosine
  return RAG_OK;
}

bool next_plan(rag_index* idx, uint32_t k, const plan& cur, plan* nxt) {
  *nxt = cur;
  if (cur.path == RAG_PATH_TENSOR || cur.path == RAG_PATH_SHADOW_STREAM) {
    make_plan(idx, 1, k, RAG_PATH_STREAM, 0, 0.0, 0, nxt);
    return true;
  }
  if (cur.path == RAG_PATH_STREAM) {
    make_plan(idx, 1, k, RAG_PATH_EXACT, 0, 0.0, 0, nxt);
    return true;
  }
  if (cur.kp < RAG_MAX_CANDIDATES) {  // exact path, wider candidate window (long runs of exact ties)
    nxt->kp = RAG_MAX_CANDIDATES;
    return true;
  }
  return false;
}

struct fresh_cfg { int64_t now_ms; double decay, bonus; };

// ---- buffers for one batch ----------------------------------------------------------
int ensure_queries(rag_index* idx, rag_batch* bt, uint32_t B) {
  return grow_dev(&bt->d_q, &bt->c_q, (size_t)B * idx->ld * sizeof(float), true);
}

int ensure_inputs(rag_index* idx, rag_batch* bt, uint32_t B, uint32_t kw_stride) {
  (void)idx;
  const size_t keys = ((size_t)B * kw_stride * 8 + 15) & ~(size_t)15;
  const size_t need = keys + (size_t)B * 4;
  RAG_CHECK(grow_dev(&bt->d_in, &bt->c_in, need, true));
  RAG_CHECK(grow_pinned(&bt->h_in, &bt->c_hin, need));
  bt->d_kw = (uint64_t*)bt->d_in;
  bt->d_kwc = (uint32_t*)(bt->d_in + keys);
  return RAG_OK;
}

int ensure_work(rag_index* idx, rag_batch* bt, uint32_t B, uint32_t k, uint32_t out_cap, out_layout* L) {
  RAG_CHECK(grow_dev(&bt->d_cand, &bt->c_cand, (size_t)B * RAG_MAX_CANDIDATES * 8, false));
  RAG_CHECK(grow_dev(&bt->d_local, &bt->c_local, (size_t)B * k * sizeof(rag_rec), false));
  RAG_CHECK(grow_dev(&bt->d_local_cnt, &bt->c_lcnt, (size_t)B * 4, false));
  if (B <= 32) {
    RAG_CHECK(grow_dev(&bt->d_k4s, &bt->c_k4s, (size_t)B * (2 * RAG_MAX_CANDIDATES + 2) * 8, false));
    RAG_CHECK(grow_dev(&bt->d_ticket, &bt->c_ticket, (size_t)32 * 4, true));
  }
  if (idx->nranks > 1) {
    RAG_CHECK(comm_p2p_ensure(idx, B, k));  // collective when the mailboxes grow; may fall back to NCCL
    if (!comm_uses_p2p(idx))
      RAG_CHECK(grow_dev(&bt->d_gather, &bt->c_gather, (size_t)idx->nranks * B * k * sizeof(rag_rec), false));
  }
  *L = layout_out(B, out_cap, k);
  RAG_CHECK(grow_dev(&bt->d_out, &bt->c_out, L->total, false));
  RAG_CHECK(grow_pinned(&bt->h_out, &bt->c_hout, L->total));
  bind_out(bt, *L);
  return RAG_OK;
}

int stage_queries(rag_index* idx, rag_batch* bt, const float* queries, uint32_t B) {
  RAG_CHECK(ensure_queries(idx, bt, B));
  if (idx->ld == idx->dim)
    RAG_CUDA(cudaMemcpyAsync(bt->d_q, queries, (size_t)B * idx->dim * 4, cudaMemcpyHostToDevice, idx->stream));
  else
    RAG_CUDA(cudaMemcpy2DAsync(bt->d_q, (size_t)idx->ld * 4, queries, (size_t)idx->dim * 4, (size_t)idx->dim * 4, B,
                               cudaMemcpyHostToDevice, idx->stream));
  return RAG_OK;
}

int stage_keywords(rag_index* idx, rag_batch* bt, uint32_t B, const uint64_t* kw_keys, const uint32_t* kw_counts,
                   uint32_t kw_stride) {
  RAG_CHECK(ensure_inputs(idx, bt, B, kw_stride));
  const size_t keys = ((size_t)B * kw_stride * 8 + 15) & ~(size_t)15;
  uint32_t* hc = (uint32_t*)(bt->h_in + keys);
  if (kw_stride && kw_keys) memcpy(bt->h_in, kw_keys, (size_t)B * kw_stride * 8);
  for (uint32_t b = 0; b < B; b++) {
    uint32_t c = (kw_counts && kw_keys) ? kw_counts[b] : 0u;
    if (c > kw_stride) return rag_set_error(RAG_ERR_INVALID, "kw_counts[%u]=%u exceeds keyword_limit=%u", b, c, kw_stride);
    hc[b] = c;
  }
  RAG_CUDA(cudaMemcpyAsync(bt->d_in, bt->h_in, keys + (size_t)B * 4, cudaMemcpyHostToDevice, idx->stream));
  bt->staged_kw_stride = kw_stride;
  return RAG_OK;
}

// ---- the pipeline (asynchronous on idx->stream; operates on idx->cur) -----------------
int run_pipeline(rag_index* idx, uint32_t B, uint32_t k, const plan& p, const fresh_cfg& fc, rag_fuse_args fa) {
  rag_batch* bt = idx->cur;
  uint32_t parts = 0;
  const bool stream = p.path == RAG_PATH_STREAM || p.path == RAG_PATH_SHADOW_STREAM;
  if (stream) RAG_CHECK(k1_plan(idx, B, p.kp, &parts));
  else if (p.path == RAG_PATH_TENSOR) RAG_CHECK(k2_plan(idx, B, p.kp, &parts));
  else RAG_CHECK(k1x_plan(idx, B, p.kp, &parts));
  RAG_CHECK(grow_dev(&bt->d_partial, &bt->c_partial, (size_t)B * parts * p.kp * 8, false));
  // The score the caller filters at (hybridSearch keeps r.score >= minVectorScore, hybrid-search.ts:308-314; MemoryStore
  // keeps relevance >= minRelevance, store.ts:151): rows below it can never reach the result, and filtering commutes with
  // taking the best k by the same score. K4 uses it to certify (a query whose non-candidates all lie below the filter is
  // exact whatever its k-th score is — the "nothing relevant" query no longer costs an escalation pass), and the tensor
  // path (16-bit operands: its keys are cosines) starts its thresholds there instead of at -inf, which removes the
  // start-up transient of the selection lists. The floor sits a margin below the filter; K4 checks rigorously, per query,
  // that floor + eps[b] < min — a query for which it does not hold is uncertified and escalates like any other.
  double min_eff = fa.mode == 0 ? fa.min_score : (fa.mode == 1 ? fa.mem_min_relevance : -INFINITY);
  if (!(min_eff > -INFINITY) || !(min_eff < INFINITY)) min_eff = -INFINITY;   // NaN / infinite: no filter to lean on
  double floor = -INFINITY;
  static const bool k2_floor_on = !(getenv("RAGERA_K2_FLOOR") && atoi(getenv("RAGERA_K2_FLOOR")) == 0);
  if (k2_floor_on && p.path == RAG_PATH_TENSOR && p.key_has_qnorm && min_eff > -INFINITY) {
    const double rho_q_typ = (idx->shadow && !idx->shadow_f16) ? 4.0e-3 : 5.0e-4;   // bf16 / fp16 rounding of a unit vector
    floor = min_eff - 2.0 * (p.eps + (p.eps_per_query ? rho_q_typ * p.eps_q_mul : 0.0)) - 1.0e-6;
    floor = (double)(float)floor;   // K2 compares in fp32: K4 must reason about the very value K2 used
  }
  if (stream) RAG_CHECK(k1_launch(idx, B, p.kp, parts, p.path == RAG_PATH_SHADOW_STREAM));
  else if (p.path == RAG_PATH_TENSOR) RAG_CHECK(k2_launch(idx, B, p.kp, parts, (float)floor));
  else RAG_CHECK(k1x_launch(idx, B, p.kp, parts));
  fa.B = B;
  fa.k = k;
  fa.nranks = (uint32_t)idx->nranks;
  // small batches on one GPU: the last CTA of each query in the K3+K4 kernel runs K5 in place
  static const bool k5_in_place = !(getenv("RAGERA_FUSE_K5") && atoi(getenv("RAGERA_FUSE_K5")) == 0);
  bool fused = false;
  const rag_eps eps = {p.eps, p.eps_per_query ? idx->cur->d_rho_q : nullptr, p.eps_q_mul, floor, min_eff};
  if (k34_small_ok(idx, B, p.kp, parts)) {
    // one GPU, or sharded with the peer-to-peer exchange (which then runs inside the same kernel)
    fused = k5_in_place && (idx->nranks == 1 || k34_small_fuses_exchange(idx));
    RAG_CHECK(k34_small_launch(idx, B, p.kp, parts, k, eps, p.key_has_qnorm, fc.now_ms, fc.decay, fc.bonus, fused ? &fa : nullptr));
  } else {
    RAG_CHECK(k3_launch(idx, B, p.kp, parts));
    RAG_CHECK(k4_launch(idx, B, p.kp, k, eps, p.key_has_qnorm, fc.now_ms, fc.decay, fc.bonus));
  }
  if (fused) return RAG_OK;
  RAG_CHECK(comm_allgather_local(idx, B, k));
  RAG_CHECK(k5_launch(idx, &fa));
  return RAG_OK;
}

int fetch_out(rag_index* idx, rag_batch* bt, const out_layout& L, bool with_aux) {
  RAG_CUDA(cudaMemcpyAsync(bt->h_out, bt->d_out, with_aux ? L.total : L.total_no_aux, cudaMemcpyDeviceToHost,
                           idx->stream));
  RAG_CUDA(cudaStreamSynchronize(idx->stream));
  RAG_CHECK(comm_check_status(idx));  // sharded: a peer that never arrived at the exchange
  return RAG_OK;
}

// writer(user_b, batch, layout, local_b): copy one query's result from the pinned mirror to the caller
typedef std::function<void(uint32_t, const rag_batch*, const out_layout&, uint32_t)> writer_fn;

// Re-run the queries whose result could not be certified on successively stronger paths.
int escalate(rag_index* idx, std::vector<uint32_t> sel, uint32_t k, plan p, const fresh_cfg& fc,
             const rag_fuse_args& fa, uint32_t out_cap, bool with_aux, const writer_fn& write, const rag_batch* src_view = nullptr) {
  const rag_batch* src = src_view ? src_view : &idx->main;  // where the batch's queries / keyword lists sit on the device
  rag_batch* bt = &idx->esc;
  int rc = RAG_OK;
  while (!sel.empty()) {
    plan nxt;
    if (!next_plan(idx, k, p, &nxt)) break;  // stays flagged as uncertified
    p = nxt;
    const uint32_t n = (uint32_t)sel.size();
    idx->cur = bt;
    out_layout L;
    if ((rc = ensure_queries(idx, bt, n)) != RAG_OK) break;
    if ((rc = ensure_inputs(idx, bt, n, fa.kw_stride)) != RAG_OK) break;
    if ((rc = ensure_work(idx, bt, n, k, out_cap, &L)) != RAG_OK) break;
    if ((rc = grow_dev(&bt->d_sel, &bt->c_sel, (size_t)n * 4, false)) != RAG_OK) break;
    cudaError_t e = cudaMemcpyAsync(bt->d_sel, sel.data(), (size_t)n * 4, cudaMemcpyHostToDevice, idx->stream);
    if (e != cudaSuccess) { rc = rag_set_error(RAG_ERR_CUDA, "escalation H2D failed: %s", cudaGetErrorString(e)); break; }
    if ((rc = gather_batch_launch(idx, src, bt, n, fa.mode == 0 ? fa.kw_stride : 0)) != RAG_OK) break;
    if ((rc = run_pipeline(idx, n, k, p, fc, fa)) != RAG_OK) break;
    if ((rc = fetch_out(idx, bt, L, with_aux)) != RAG_OK) break;
    const uint8_t* cert = bt->h_out + L.cert;
    std::vector<uint32_t> still;
    for (uint32_t i = 0; i < n; i++) {
      write(sel[i], bt, L, i);
      if (!cert[i]) still.push_back(sel[i]);
    }
    sel.swap(still);
  }
  idx->cur = &idx->main;
  return rc;
}

// ---- small-batch latency path as ONE graph launch ---------------------------------------------------------------
// A batch-1 search over a small corpus (the reference's production size: C1, 10k chunks) is launch-bound: two H2D copies,
// K1, the fused K3+K4+K5 kernel and the D2H are five API calls and four stream dependencies. The sequence is captured
// once per call shape and replayed with a single cudaGraphLaunch: the inputs are first copied into pinned buffers with
// fixed addresses (the caller's pointers change from call to call), every kernel argument is part of the cache key.
struct graph_key {
  uint32_t B = 0, k = 0, kp = 0, parts = 0, kw_stride = 0, out_cap = 0, fresh_limit = 0;
  int path = 0, mode = 0, key_has_qnorm = 0;
  double eps = 0, min_score = 0, rrf_k = 0, rrf_vw = 0, rrf_kw = 0, rrf_bonus = 0, fresh_weight = 0, mem_min = 0;
  uint32_t mem_limit = 0;
  uint64_t rows = 0;
  const void* ptr[10] = {nullptr};
  size_t d2h = 0;
  bool same(const graph_key& o) const {
    if (B != o.B || k != o.k || kp != o.kp || parts != o.parts || kw_stride != o.kw_stride || out_cap != o.out_cap ||
        fresh_limit != o.fresh_limit || path != o.path || mode != o.mode || key_has_qnorm != o.key_has_qnorm || eps != o.eps ||
        min_score != o.min_score || rrf_k != o.rrf_k || rrf_vw != o.rrf_vw || rrf_kw != o.rrf_kw || rrf_bonus != o.rrf_bonus ||
        fresh_weight != o.fresh_weight || mem_min != o.mem_min || mem_limit != o.mem_limit || rows != o.rows || d2h != o.d2h)
      return false;
    for (int i = 0; i < 10; i++)
      if (ptr[i] != o.ptr[i]) return false;
    return true;
  }
};
}  // namespace
struct rag_graph {
  cudaGraphExec_t exec = nullptr;
  graph_key key;
  uint64_t launches = 0;     // kernel launches one replay stands for
  uint32_t recaptures = 0;   // consecutive misses: a caller that changes shape on every call gets the plain path
  bool disabled = false;
};
namespace {
bool graphs_enabled() {
  static const bool on = !(getenv("RAGERA_GRAPH") && atoi(getenv("RAGERA_GRAPH")) == 0);
  return on;
}

// Is this call the latency shape? one GPU, batch <= 32 on the stream path with the fused tail, no per-kernel profiling,
// and nothing in the pipeline that reads the wall clock of the call (freshness needs now_ms, which changes every call).
bool graph_shape_ok(const rag_index* idx, uint32_t B, const plan& p, const rag_fuse_args& fa) {
  return graphs_enabled() && idx->nranks == 1 && !idx->prof_on && B <= 32 && (p.path == RAG_PATH_STREAM || p.path == RAG_PATH_SHADOW_STREAM) && !p.eps_per_query &&
         (fa.mode == 2 || (fa.mode == 0 && fa.fresh_limit == 0)) && !(idx->graph && idx->graph->disabled);
}

// Runs stage→pipeline→fetch for idx->main as one graph launch. queries/kw are the caller's host arrays.
// Graph nodes: ONE H2D (queries and keyword lists packed in one pinned block → one device block), K1, the fused
// K3+K4+K5 kernel. There is no D2H node: the fusion kernel writes the result block straight into pinned host memory.
int run_graphed(rag_index* idx, const float* queries, uint32_t B, uint32_t k, const plan& p, const fresh_cfg& fc, rag_fuse_args fa,
                const uint64_t* kw_keys, const uint32_t* kw_counts, uint32_t kw_stride, uint32_t out_cap, out_layout* L, bool* ran,
                rag_batch* view_out) {
  *ran = false;
  rag_batch* bt = &idx->main;
  idx->cur = bt;
  // 1. every allocation up front (none may happen inside a capture)
  RAG_CHECK(ensure_work(idx, bt, B, k, out_cap, L));
  uint32_t parts = 0;
  RAG_CHECK(k1_plan(idx, B, p.kp, &parts));
  RAG_CHECK(grow_dev(&bt->d_partial, &bt->c_partial, (size_t)B * parts * p.kp * 8, false));
  if (!k34_small_ok(idx, B, p.kp, parts)) return RAG_OK;  // not the fused tail: plain path
  const bool with_kw = fa.mode == 0;
  const size_t qb = (size_t)B * idx->ld * 4;
  const size_t kwb = with_kw ? (((size_t)B * kw_stride * 8 + 15) & ~(size_t)15) : 0;
  const size_t in_bytes = qb + kwb + (with_kw ? (size_t)B * 4 : 0);
  if (in_bytes > bt->c_hq || !bt->h_q) {
    if (bt->h_q) RAG_CUDA(cudaFreeHost(bt->h_q));
    bt->h_q = nullptr;
    bt->c_hq = 0;
    const size_t cap = std::max<size_t>(in_bytes, 64 * 1024);
    RAG_CUDA(cudaHostAlloc((void**)&bt->h_q, cap, cudaHostAllocDefault));
    memset(bt->h_q, 0, cap);  // the padding columns of the query rows stay zero
    bt->c_hq = cap;
  }
  RAG_CHECK(grow_dev(&bt->d_gin, &bt->c_gin, std::max<size_t>(in_bytes, 64 * 1024), true));
  // 2. inputs into the pinned block (fixed address: the caller's pointers change from call to call)
  uint8_t* hin = reinterpret_cast<uint8_t*>(bt->h_q);
  if (idx->ld == idx->dim) memcpy(hin, queries, qb);
  else
    for (uint32_t b = 0; b < B; b++) memcpy(hin + (size_t)b * idx->ld * 4, queries + (size_t)b * idx->dim, (size_t)idx->dim * 4);
  if (with_kw) {
    uint32_t* hc = (uint32_t*)(hin + qb + kwb);
    if (kw_stride && kw_keys) memcpy(hin + qb, kw_keys, (size_t)B * kw_stride * 8);
    for (uint32_t b = 0; b < B; b++) {
      const uint32_t c = (kw_counts && kw_keys) ? kw_counts[b] : 0u;
      if (c > kw_stride) return rag_set_error(RAG_ERR_INVALID, "kw_counts[%u]=%u exceeds keyword_limit=%u", b, c, kw_stride);
      hc[b] = c;
    }
  }
  // 3. the batch as the kernels see it: inputs in the packed device block, outputs in the pinned mirror
  rag_batch view = *bt;
  view.d_q = reinterpret_cast<float*>(bt->d_gin);
  view.d_kw = reinterpret_cast<uint64_t*>(bt->d_gin + qb);
  view.d_kwc = reinterpret_cast<uint32_t*>(bt->d_gin + qb + kwb);
  view.staged_kw_stride = kw_stride;
  bind_out_at(&view, *L, bt->h_out);
  // 4. the key: everything a captured node bakes in
  graph_key key;
  key.B = B; key.k = k; key.kp = p.kp; key.parts = parts; key.kw_stride = with_kw ? kw_stride : 0; key.out_cap = out_cap;
  key.fresh_limit = fa.fresh_limit; key.path = p.path; key.mode = fa.mode; key.key_has_qnorm = p.key_has_qnorm; key.eps = p.eps;
  key.min_score = fa.min_score; key.rrf_k = fa.rrf.k; key.rrf_vw = fa.rrf.vector_weight; key.rrf_kw = fa.rrf.keyword_weight;
  key.rrf_bonus = fa.rrf.both_bonus; key.fresh_weight = fa.fresh_weight; key.mem_min = fa.mem_min_relevance; key.mem_limit = fa.mem_limit;
  key.rows = idx->rows;
  const void* ptrs[10] = {bt->d_gin, bt->h_q, bt->h_out, bt->d_partial, bt->d_cand, bt->d_local, bt->d_k4s, bt->d_ticket, idx->row_keys, idx->ctype};
  for (int i = 0; i < 10; i++) key.ptr[i] = ptrs[i];
  key.d2h = in_bytes;
  if (!idx->graph) idx->graph = new (std::nothrow) rag_graph();
  rag_graph* g = idx->graph;
  if (!g) return RAG_OK;
  if (!g->exec || !g->key.same(key)) {
    if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; }
    if (++g->recaptures > 8) { g->disabled = true; return RAG_OK; }  // shape changes on every call: capturing costs more than it saves
    const uint64_t l0 = idx->launches;
    RAG_CUDA(cudaStreamBeginCapture(idx->stream, cudaStreamCaptureModeThreadLocal));
    int rc = RAG_OK;
    cudaError_t e = cudaMemcpyAsync(bt->d_gin, bt->h_q, in_bytes, cudaMemcpyHostToDevice, idx->stream);
    if (e != cudaSuccess) rc = rag_set_error(RAG_ERR_CUDA, "graph capture (H2D): %s", cudaGetErrorString(e));
    idx->cur = &view;
    if (rc == RAG_OK) rc = run_pipeline(idx, B, k, p, fc, fa);
    idx->cur = bt;
    cudaGraph_t graph = nullptr;
    e = cudaStreamEndCapture(idx->stream, &graph);
    if (rc == RAG_OK && e != cudaSuccess) rc = rag_set_error(RAG_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
    if (rc == RAG_OK) {
      e = cudaGraphInstantiate(&g->exec, graph, 0);
      if (e != cudaSuccess) rc = rag_set_error(RAG_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    }
    if (graph) cudaGraphDestroy(graph);
    if (rc != RAG_OK) { g->exec = nullptr; cudaGetLastError(); return rc; }
    g->key = key;
    g->launches = idx->launches - l0;
    idx->launches = l0;
  } else {
    g->recaptures = 0;
  }
  bt->staged_B = 0;   // the staged pool (rag_stage_batch) is not what this call used
  bt->win_count = 0;
  RAG_CUDA(cudaGraphLaunch(g->exec, idx->stream));
  idx->launches += g->launches;
  RAG_CUDA(cudaStreamSynchronize(idx->stream));
  *view_out = view;  // an escalation of uncertified queries gathers its inputs from here
  *ran = true;
  return RAG_OK;
}

bool small_prof_on() {
  static const bool on = getenv("RAGERA_SMALL_PROF") && atoi(getenv("RAGERA_SMALL_PROF")) != 0;
  return on;
}

int check_handle(const rag_index* idx) {
  if (!idx) return rag_set_error(RAG_ERR_INVALID, "null index handle");
  return RAG_OK;
}

int ensure_meta(rag_index* idx) {
  if (idx->ctype) return RAG_OK;
  const size_t cap = idx->desc.capacity_rows;
  size_t dummy = 0;
  RAG_CHECK(grow_dev(&idx->ctype, &dummy, cap, true)); dummy = 0;
  RAG_CHECK(grow_dev(&idx->conf, &dummy, cap * 8, true)); dummy = 0;
  RAG_CHECK(grow_dev(&idx->access, &dummy, cap * 4, true)); dummy = 0;
  RAG_CHECK(grow_dev(&idx->last_ms, &dummy, cap * 8, true));
  return RAG_OK;
}

void fresh_defaults(double* decay, double* bonus) {
  if (*decay == 0.0) *decay = 0.05;  // DEFAULT_FRESHNESS_CONFIG, src/lib/memory/freshness.ts:20-23
  if (*bonus == 0.0) *bonus = 0.1;
}

}  // namespace

int rag_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ---- profiling spans -------------------------------------------------------------------
static void prof_drain(rag_index* idx) {
  if (idx->prof_used == 0) return;
  cudaStreamSynchronize(idx->stream);
  for (uint32_t i = 0; i < idx->prof_used; i++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, idx->prof_spans[i].a, idx->prof_spans[i].b) == cudaSuccess) {
      idx->prof_ms[idx->prof_spans[i].cls] += ms;
      idx->prof_cnt[idx->prof_spans[i].cls] += 1;
    }
  }
  cudaGetLastError();
  idx->prof_used = 0;
}

int rag_prof_begin(rag_index* idx, int cls) {
  if (!idx->prof_spans) return -1;
  if (idx->prof_used == idx->prof_cap) prof_drain(idx);
  const int t = (int)idx->prof_used++;
  idx->prof_spans[t].cls = cls;
  cudaEventRecord(idx->prof_spans[t].a, idx->stream);
  return t;
}
void rag_prof_end(rag_index* idx, int token) { cudaEventRecord(idx->prof_spans[token].b, idx->stream); }

extern "C" {

int rag_version(void) { return RAGERA_VERSION; }
const char* rag_last_error(void) { return g_err; }

int rag_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int rag_index_create(const rag_index_desc* d, rag_index** out) {
  if (!d || !out) return rag_set_error(RAG_ERR_INVALID, "rag_index_create: null argument");
  *out = nullptr;
  if (d->dim == 0 || d->dim > 8192) return rag_set_error(RAG_ERR_INVALID, "dim must be in 1..8192");
  if (d->capacity_rows == 0 || d->capacity_rows >= 0xFFFFFFFFull)
    return rag_set_error(RAG_ERR_INVALID, "capacity_rows must be in 1..2^32-2 per shard");
  if (d->dtype != RAG_F32 && d->dtype != RAG_BF16) return rag_set_error(RAG_ERR_INVALID, "unknown dtype %u", d->dtype);
  const int ndev = rag_device_count();
  if (ndev == 0) return rag_set_error(RAG_ERR_NO_DEVICE, "no CUDA device visible; libragera has no CPU fallback");
  if (d->device < 0 || d->device >= ndev) return rag_set_error(RAG_ERR_INVALID, "device %d out of range (0..%d)", d->device, ndev - 1);
  int major = 0;
  RAG_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d->device));
  if (major != 10)
    return rag_set_error(RAG_ERR_NO_DEVICE, "device %d is sm_%d0: libragera is built for sm_100a only (no fallback)",
                         d->device, major);
  RAG_CUDA(cudaSetDevice(d->device));

  rag_index* idx = new (std::nothrow) rag_index();
  if (!idx) return rag_set_error(RAG_ERR_NOMEM, "out of host memory");
  idx->desc = *d;
  idx->device = d->device;
  idx->dim = d->dim;
  idx->ld = (d->dim + 255u) & ~255u;
  idx->cur = &idx->main;
  if (small_prof_on()) { k1_small_prof(1, nullptr); k34_small_prof(1, nullptr); }
  int rc = RAG_OK;
  do {
    cudaError_t e;
    if ((e = cudaDeviceGetAttribute(&idx->sm_count, cudaDevAttrMultiProcessorCount, d->device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&idx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&idx->ev0)) != cudaSuccess || (e = cudaEventCreate(&idx->ev1)) != cudaSuccess) {
      rc = rag_set_error(RAG_ERR_CUDA, "rag_index_create: %s", cudaGetErrorString(e));
      break;
    }
    size_t cap = 0;
    const size_t bytes = (size_t)d->capacity_rows * idx->ld * elem_size(idx);
    char* corpus = nullptr;
    if ((rc = grow_dev(&corpus, &cap, bytes, false)) != RAG_OK) break;
    idx->corpus = corpus;
    if (idx->ld != idx->dim) {  // padding columns must read as zero
      if ((e = cudaMemsetAsync(idx->corpus, 0, bytes, idx->stream)) != cudaSuccess) {
        rc = rag_set_error(RAG_ERR_CUDA, "corpus memset: %s", cudaGetErrorString(e));
        break;
      }
    }
    if (d->dtype == RAG_BF16) {
      idx->shadow = (__nv_bfloat16*)idx->corpus;
    } else if (d->flags & (RAG_INDEX_BF16_SHADOW | RAG_INDEX_F16_SHADOW)) {
      idx->shadow_f16 = (d->flags & RAG_INDEX_F16_SHADOW) != 0;
      cap = 0;
      if ((rc = grow_dev(&idx->shadow, &cap, (size_t)d->capacity_rows * idx->ld * 2, false)) != RAG_OK) break;
    }
    // 1/||x|| per row of the tensor-path operand (the bf16 rows, or the fp32 rows read as tf32)
    cap = 0;
    if ((rc = grow_dev(&idx->inv_norm, &cap, (size_t)d->capacity_rows * 4, true)) != RAG_OK) break;
    cap = 0;
    if ((rc = grow_dev(&idx->d_rho_x, &cap, 16, true)) != RAG_OK) break;
    cap = 0;
    if ((rc = grow_dev(&idx->d_counters, &cap, 16, true)) != RAG_OK) break;
    if ((e = cudaStreamSynchronize(idx->stream)) != cudaSuccess) {
      rc = rag_set_error(RAG_ERR_CUDA, "rag_index_create: %s", cudaGetErrorString(e));
      break;
    }
  } while (0);
  if (rc != RAG_OK) {
    rag_index_destroy(idx);
    return rc;
  }
  *out = idx;
  return RAG_OK;
}

void rag_index_destroy(rag_index* idx) {
  if (!idx) return;
  { RAG_LOCK(idx); }  // a call still in flight on another thread finishes first (calls after destroy are the caller's bug)
  cudaSetDevice(idx->device);
  if (small_prof_on()) { k1_small_prof(-1, stderr); k34_small_prof(-1, stderr); }
  if (idx->stream) cudaStreamSynchronize(idx->stream);
  rag_comm_destroy(idx);
  k2_destroy(idx);
  if (idx->graph) {
    if (idx->graph->exec) cudaGraphExecDestroy(idx->graph->exec);
    delete idx->graph;
  }
  free_batch(&idx->main);
  free_batch(&idx->esc);
  if (idx->shadow && (void*)idx->shadow != idx->corpus) cudaFree(idx->shadow);
  cudaFree(idx->corpus); cudaFree(idx->inv_norm); cudaFree(idx->d_rho_x); cudaFree(idx->d_counters); cudaFree(idx->ctype); cudaFree(idx->conf);
  cudaFree(idx->access); cudaFree(idx->last_ms); cudaFree(idx->row_keys);
  if (idx->prof_spans) {
    for (uint32_t i = 0; i < idx->prof_cap; i++) { cudaEventDestroy(idx->prof_spans[i].a); cudaEventDestroy(idx->prof_spans[i].b); }
    delete[] idx->prof_spans;
  }
  if (idx->ev0) cudaEventDestroy(idx->ev0);
  if (idx->ev1) cudaEventDestroy(idx->ev1);
  if (idx->stream) cudaStreamDestroy(idx->stream);
  cudaGetLastError();
  delete idx;
}

uint64_t rag_index_rows(const rag_index* idx) { return idx ? idx->rows : 0; }

int rag_index_upload(rag_index* idx, uint64_t row0, uint64_t nrows, const void* host_rows) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  if (nrows == 0) return RAG_OK;
  if (!host_rows) return rag_set_error(RAG_ERR_INVALID, "rag_index_upload: null rows");
  if (row0 > idx->rows) return rag_set_error(RAG_ERR_INVALID, "rag_index_upload: row0=%llu leaves a gap (rows=%llu)",
                                             (unsigned long long)row0, (unsigned long long)idx->rows);
  if (row0 + nrows > idx->desc.capacity_rows)
    return rag_set_error(RAG_ERR_INVALID, "rag_index_upload: %llu rows exceed capacity %llu",
                         (unsigned long long)(row0 + nrows), (unsigned long long)idx->desc.capacity_rows);
  RAG_CUDA(cudaSetDevice(idx->device));
  const size_t es = elem_size(idx);
  char* dst = (char*)idx->corpus + (size_t)row0 * idx->ld * es;
  // 2-D copies are limited in height; upload in slabs
  const uint64_t slab = 1u << 20;
  for (uint64_t r = 0; r < nrows; r += slab) {
    const uint64_t h = std::min(slab, nrows - r);
    RAG_CUDA(cudaMemcpy2DAsync(dst + (size_t)r * idx->ld * es, (size_t)idx->ld * es,
                               (const char*)host_rows + (size_t)r * idx->dim * es, (size_t)idx->dim * es,
                               (size_t)idx->dim * es, h, cudaMemcpyHostToDevice, idx->stream));
  }
  idx->rows = std::max(idx->rows, row0 + nrows);
  if (idx->row_keys) RAG_CHECK(iota_u64_launch(idx, idx->row_keys + row0, nrows, idx->desc.id_base + row0));
  RAG_CHECK(aux_build_launch(idx, row0, nrows));
  RAG_CUDA(cudaStreamSynchronize(idx->stream));
  return RAG_OK;
}

int rag_index_generate(rag_index* idx, const rag_gen_desc* gen, uint64_t nrows) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  if (!gen || gen->n_clusters == 0 || gen->total_rows == 0)
    return rag_set_error(RAG_ERR_INVALID, "rag_index_generate: bad generator description");
  if (nrows > idx->desc.capacity_rows) return rag_set_error(RAG_ERR_INVALID, "rag_index_generate: nrows exceeds capacity");
  RAG_CUDA(cudaSetDevice(idx->device));
  RAG_CHECK(gen_corpus_launch(idx, gen, nrows));
  idx->rows = nrows;
  RAG_CHECK(ensure_meta(idx));
  RAG_CHECK(gen_meta_launch(idx, gen, nrows));
  if (idx->row_keys) RAG_CHECK(iota_u64_launch(idx, idx->row_keys, nrows, idx->desc.id_base));
  RAG_CHECK(aux_build_launch(idx, 0, nrows));
  RAG_CUDA(cudaStreamSynchronize(idx->stream));
  return RAG_OK;
}

int rag_index_set_row_meta(rag_index* idx, uint64_t row0, uint64_t nrows, const uint8_t* content_type,
                           const double* confidence, const int32_t* access_count, const int64_t* last_access_ms) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  if (row0 + nrows > idx->desc.capacity_rows) return rag_set_error(RAG_ERR_INVALID, "rag_index_set_row_meta: range exceeds capacity");
  if (nrows == 0) return RAG_OK;
  RAG_CUDA(cudaSetDevice(idx->device));
  RAG_CHECK(ensure_meta(idx));
  if (content_type) RAG_CUDA(cudaMemcpyAsync(idx->ctype + row0, content_type, nrows, cudaMemcpyHostToDevice, idx->stream));
  if (confidence) RAG_CUDA(cudaMemcpyAsync(idx->conf + row0, confidence, nrows * 8, cudaMemcpyHostToDevice, idx->stream));
  if (access_count) RAG_CUDA(cudaMemcpyAsync(idx->access + row0, access_count, nrows * 4, cudaMemcpyHostToDevice, idx->stream));
  if (last_access_ms) RAG_CUDA(cudaMemcpyAsync(idx->last_ms + row0, last_access_ms, nrows * 8, cudaMemcpyHostToDevice, idx->stream));
  RAG_CUDA(cudaStreamSynchronize(idx->stream));
  return RAG_OK;
}

int rag_index_set_row_keys(rag_index* idx, uint64_t row0, uint64_t nrows, const uint64_t* keys) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  if (row0 + nrows > idx->desc.capacity_rows) return rag_set_error(RAG_ERR_INVALID, "rag_index_set_row_keys: range exceeds capacity");
  if (nrows == 0) return RAG_OK;
  if (!keys) return rag_set_error(RAG_ERR_INVALID, "rag_index_set_row_keys: null keys");
  RAG_CUDA(cudaSetDevice(idx->device));
  if (!idx->row_keys) {
    size_t cap = 0;
    RAG_CHECK(grow_dev(&idx->row_keys, &cap, (size_t)idx->desc.capacity_rows * 8, false));
    RAG_CHECK(iota_u64_launch(idx, idx->row_keys, idx->desc.capacity_rows, idx->desc.id_base));  // default key = chunk id
  }
  RAG_CUDA(cudaMemcpyAsync(idx->row_keys + row0, keys, nrows * 8, cudaMemcpyHostToDevice, idx->stream));
  RAG_CUDA(cudaStreamSynchronize(idx->stream));
  return RAG_OK;
}

int rag_index_read_row_meta(rag_index* idx, uint64_t row0, uint64_t nrows, uint8_t* content_type, double* confidence,
                            int32_t* access_count, int64_t* last_access_ms, uint64_t* keys) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  if (row0 + nrows > idx->rows) return rag_set_error(RAG_ERR_INVALID, "rag_index_read_row_meta: range exceeds rows");
  if (nrows == 0) return RAG_OK;
  RAG_CUDA(cudaSetDevice(idx->device));
  // metadata never set: every row is a document with zeroed Memory columns; keys default to the chunk id
  if (!idx->ctype) {
    if (content_type) memset(content_type, RAG_CT_DOCUMENT, nrows);
    if (confidence) memset(confidence, 0, nrows * 8);
    if (access_count) memset(access_count, 0, nrows * 4);
    if (last_access_ms) memset(last_access_ms, 0, nrows * 8);
  } else {
    if (content_type) RAG_CUDA(cudaMemcpyAsync(content_type, idx->ctype + row0, nrows, cudaMemcpyDeviceToHost, idx->stream));
    if (confidence) RAG_CUDA(cudaMemcpyAsync(confidence, idx->conf + row0, nrows * 8, cudaMemcpyDeviceToHost, idx->stream));
    if (access_count) RAG_CUDA(cudaMemcpyAsync(access_count, idx->access + row0, nrows * 4, cudaMemcpyDeviceToHost, idx->stream));
    if (last_access_ms) RAG_CUDA(cudaMemcpyAsync(last_access_ms, idx->last_ms + row0, nrows * 8, cudaMemcpyDeviceToHost, idx->stream));
  }
  if (keys) {
    if (idx->row_keys) RAG_CUDA(cudaMemcpyAsync(keys, idx->row_keys + row0, nrows * 8, cudaMemcpyDeviceToHost, idx->stream));
    else for (uint64_t r = 0; r < nrows; r++) keys[r] = idx->desc.id_base + row0 + r;
  }
  RAG_CUDA(cudaStreamSynchronize(idx->stream));
  return RAG_OK;
}

int rag_index_read_rows(rag_index* idx, uint64_t row0, uint64_t nrows, void* host_rows) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  if (row0 + nrows > idx->rows) return rag_set_error(RAG_ERR_INVALID, "rag_index_read_rows: range exceeds rows");
  if (nrows == 0) return RAG_OK;
  if (!host_rows) return rag_set_error(RAG_ERR_INVALID, "rag_index_read_rows: null buffer");
  RAG_CUDA(cudaSetDevice(idx->device));
  const size_t es = elem_size(idx);
  const uint64_t slab = 1u << 20;
  for (uint64_t r = 0; r < nrows; r += slab) {
    const uint64_t h = std::min(slab, nrows - r);
    RAG_CUDA(cudaMemcpy2DAsync((char*)host_rows + (size_t)r * idx->dim * es, (size_t)idx->dim * es,
                               (const char*)idx->corpus + (size_t)(row0 + r) * idx->ld * es, (size_t)idx->ld * es,
                               (size_t)idx->dim * es, h, cudaMemcpyDeviceToHost, idx->stream));
  }
  RAG_CUDA(cudaStreamSynchronize(idx->stream));
  return RAG_OK;
}

int rag_generate_queries(rag_index* idx, const rag_gen_desc* gen, uint64_t b0, uint32_t B, float* host_out) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  if (!gen || !host_out || B == 0 || gen->total_rows == 0 || gen->n_clusters == 0)
    return rag_set_error(RAG_ERR_INVALID, "rag_generate_queries: bad argument");
  RAG_CUDA(cudaSetDevice(idx->device));
  float* d = nullptr;
  size_t cap = 0;
  RAG_CHECK(grow_dev(&d, &cap, (size_t)B * idx->dim * 4, false));
  int rc = gen_queries_launch(idx, gen, b0, B, d);
  if (rc == RAG_OK) {
    cudaError_t e = cudaMemcpyAsync(host_out, d, (size_t)B * idx->dim * 4, cudaMemcpyDeviceToHost, idx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(idx->stream);
    if (e != cudaSuccess) rc = rag_set_error(RAG_ERR_CUDA, "rag_generate_queries: %s", cudaGetErrorString(e));
  }
  cudaFree(d);
  return rc;
}

// ---- search --------------------------------------------------------------------------------
int rag_search(rag_index* idx, const float* queries, uint32_t B, const rag_search_opts* o, rag_topk_out* out) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  if (!queries || !o || !out || !out->ids || !out->scores || !out->counts || B == 0)
    return rag_set_error(RAG_ERR_INVALID, "rag_search: null argument or empty batch");
  if (o->k == 0 || o->k > RAG_MAX_TOPK) return rag_set_error(RAG_ERR_INVALID, "k must be in 1..%d", RAG_MAX_TOPK);
  if (idx->rows == 0) return rag_set_error(RAG_ERR_STATE, "search on an empty index");
  RAG_CUDA(cudaSetDevice(idx->device));
  const uint32_t k = o->k;
  plan p;
  RAG_CHECK(make_plan(idx, B, k, o->path, o->slack, o->epsilon, o->flags, &p));
  rag_batch* bt = &idx->main;
  idx->cur = bt;
  out_layout L;
  rag_fuse_args fa = {};
  fa.mode = 2;
  fa.out_cap = k;
  const fresh_cfg fc = {0, 0.05, 0.1};
  bool graphed = false;
  rag_batch gview;
  if (graph_shape_ok(idx, B, p, fa)) RAG_CHECK(run_graphed(idx, queries, B, k, p, fc, fa, nullptr, nullptr, 0, k, &L, &graphed, &gview));
  if (!graphed) {
    RAG_CHECK(stage_queries(idx, bt, queries, B));
    RAG_CHECK(ensure_work(idx, bt, B, k, k, &L));
    RAG_CHECK(run_pipeline(idx, B, k, p, fc, fa));
    RAG_CHECK(fetch_out(idx, bt, L, false));
  }

  writer_fn write = [&](uint32_t ub, const rag_batch* src, const out_layout& SL, uint32_t lb) {
    const uint32_t cnt = ((const uint32_t*)(src->h_out + SL.cnt))[lb];
    const uint64_t* keys = (const uint64_t*)(src->h_out + SL.keys) + (size_t)lb * k;
    const double* sc = (const double*)(src->h_out + SL.scores) + (size_t)lb * k;
    for (uint32_t i = 0; i < k; i++) {
      out->ids[(size_t)ub * k + i] = i < cnt ? keys[i] : ~0ull;
      out->scores[(size_t)ub * k + i] = i < cnt ? sc[i] : -INFINITY;
    }
    out->counts[ub] = cnt;
    if (out->certified) out->certified[ub] = (src->h_out + SL.cert)[lb];
  };
  std::vector<uint32_t> sel;
  const uint8_t* cert = bt->h_out + L.cert;
  for (uint32_t b = 0; b < B; b++) {
    write(b, bt, L, b);
    if (!cert[b]) sel.push_back(b);
  }
  if (!sel.empty() && !(o->flags & RAG_SEARCH_NO_ESCALATE)) RAG_CHECK(escalate(idx, sel, k, p, fc, fa, k, false, write, graphed ? &gview : nullptr));
  return RAG_OK;
}

// ---- hybrid search ---------------------------------------------------------------------------
static int hybrid_args(const rag_index* idx, const rag_hybrid_opts* o, rag_fuse_args* fa, fresh_cfg* fc,
                       uint32_t* out_cap) {
  if (o->vector_top_k == 0 || o->vector_top_k > RAG_MAX_TOPK)
    return rag_set_error(RAG_ERR_INVALID, "vector_top_k must be in 1..%d", RAG_MAX_TOPK);
  if (o->keyword_limit > RAG_MAX_KEYWORDS) return rag_set_error(RAG_ERR_INVALID, "keyword_limit must be <= %d", RAG_MAX_KEYWORDS);
  if (o->fresh_limit > RAG_MAX_FRESH) return rag_set_error(RAG_ERR_INVALID, "fresh_limit must be <= %d", RAG_MAX_FRESH);
  if (idx->rows == 0) return rag_set_error(RAG_ERR_STATE, "search on an empty index");
  *fa = rag_fuse_args();
  fa->mode = 0;
  fa->kw_stride = o->keyword_limit;
  fa->min_score = o->min_vector_score;
  fa->rrf = o->rrf;
  fa->fresh_limit = o->fresh_limit;
  fa->fresh_weight = o->fresh_weight;
  *out_cap = o->vector_top_k + o->keyword_limit + o->fresh_limit;
  fa->out_cap = *out_cap;
  fc->now_ms = o->now_ms;
  fc->decay = o->time_decay_factor;
  fc->bonus = o->frequency_bonus;
  fresh_defaults(&fc->decay, &fc->bonus);
  return RAG_OK;
}

static void write_fused(const rag_hybrid_opts* o, rag_fused_out* out, uint32_t out_cap, uint32_t ub, const rag_batch* src,
                        const out_layout& SL, uint32_t lb) {
  const uint32_t k = o->vector_top_k;
  const uint32_t cnt = ((const uint32_t*)(src->h_out + SL.cnt))[lb];
  const uint64_t* keys = (const uint64_t*)(src->h_out + SL.keys) + (size_t)lb * out_cap;
  const double* sc = (const double*)(src->h_out + SL.scores) + (size_t)lb * out_cap;
  const uint8_t* s8 = src->h_out + SL.src + (size_t)lb * out_cap;
  const uint8_t* c8 = src->h_out + SL.ct + (size_t)lb * out_cap;
  const size_t base = (size_t)ub * out->capacity;
  for (uint32_t i = 0; i < out->capacity; i++) {
    const bool live = i < cnt;
    out->keys[base + i] = live ? keys[i] : ~0ull;
    out->scores[base + i] = live ? sc[i] : -INFINITY;
    if (out->source) out->source[base + i] = live ? s8[i] : 0;
    if (out->content_type) out->content_type[base + i] = live ? c8[i] : 0;
  }
  out->counts[ub] = cnt;
  if (out->used_rrf) out->used_rrf[ub] = (src->h_out + SL.rrf)[lb];
  if (out->certified) out->certified[ub] = (src->h_out + SL.cert)[lb];
  if (out->vec_ids && out->vec_scores && out->vec_counts) {
    const uint32_t vc = ((const uint32_t*)(src->h_out + SL.vcnt))[lb];
    const uint64_t* vi = (const uint64_t*)(src->h_out + SL.vids) + (size_t)lb * k;
    const double* vs = (const double*)(src->h_out + SL.vscores) + (size_t)lb * k;
    for (uint32_t i = 0; i < k; i++) {
      out->vec_ids[(size_t)ub * k + i] = i < vc ? vi[i] : ~0ull;
      out->vec_scores[(size_t)ub * k + i] = i < vc ? vs[i] : -INFINITY;
    }
    out->vec_counts[ub] = vc;
  }
}

static int check_fused_out(const rag_hybrid_opts* o, const rag_fused_out* out, uint32_t out_cap) {
  if (!out || !out->keys || !out->scores || !out->counts) return rag_set_error(RAG_ERR_INVALID, "rag_fused_out: null arrays");
  if (out->capacity < out_cap)
    return rag_set_error(RAG_ERR_INVALID, "rag_fused_out.capacity=%u < vector_top_k+keyword_limit+fresh_limit=%u",
                         out->capacity, out_cap);
  (void)o;
  return RAG_OK;
}

int rag_hybrid_search(rag_index* idx, const float* queries, uint32_t B, const rag_hybrid_opts* o,
                      const uint64_t* kw_keys, const uint32_t* kw_counts, rag_fused_out* out) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  if (!queries || !o || B == 0) return rag_set_error(RAG_ERR_INVALID, "rag_hybrid_search: null argument or empty batch");
  rag_fuse_args fa;
  fresh_cfg fc;
  uint32_t out_cap = 0;
  RAG_CHECK(hybrid_args(idx, o, &fa, &fc, &out_cap));
  RAG_CHECK(check_fused_out(o, out, out_cap));
  RAG_CUDA(cudaSetDevice(idx->device));
  const uint32_t k = o->vector_top_k;
  plan p;
  RAG_CHECK(make_plan(idx, B, k, o->path, o->slack, o->epsilon, o->flags, &p));
  rag_batch* bt = &idx->main;
  idx->cur = bt;
  out_layout L;
  bool graphed = false;
  rag_batch gview;
  if (graph_shape_ok(idx, B, p, fa))   // the latency shape: one graph launch for H2D + K1 + K3/K4/K5 (results land in pinned memory)
    RAG_CHECK(run_graphed(idx, queries, B, k, p, fc, fa, kw_keys, kw_counts, o->keyword_limit, out_cap, &L, &graphed, &gview));
  if (!graphed) {
    RAG_CHECK(stage_queries(idx, bt, queries, B));
    RAG_CHECK(stage_keywords(idx, bt, B, kw_keys, kw_counts, o->keyword_limit));
    bt->staged_B = B;
    bt->win_first = 0;
    bt->win_count = B;
    RAG_CHECK(ensure_work(idx, bt, B, k, out_cap, &L));
    RAG_CHECK(run_pipeline(idx, B, k, p, fc, fa));
    RAG_CHECK(fetch_out(idx, bt, L, false));
  }
  writer_fn write = [&](uint32_t ub, const rag_batch* src, const out_layout& SL, uint32_t lb) {
    write_fused(o, out, out_cap, ub, src, SL, lb);
  };
  std::vector<uint32_t> sel;
  const uint8_t* cert = bt->h_out + L.cert;
  for (uint32_t b = 0; b < B; b++) {
    write(b, bt, L, b);
    if (!cert[b]) sel.push_back(b);
  }
  if (!sel.empty() && !(o->flags & RAG_SEARCH_NO_ESCALATE))
    RAG_CHECK(escalate(idx, sel, k, p, fc, fa, out_cap, false, write, graphed ? &gview : nullptr));
  return RAG_OK;
}

// ---- staged form -----------------------------------------------------------------------------
int rag_stage_batch(rag_index* idx, const float* queries, uint32_t B, const uint64_t* kw_keys,
                    const uint32_t* kw_counts, uint32_t keyword_limit) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  if (!queries || B == 0) return rag_set_error(RAG_ERR_INVALID, "rag_stage_batch: null queries or empty batch");
  if (keyword_limit > RAG_MAX_KEYWORDS) return rag_set_error(RAG_ERR_INVALID, "keyword_limit must be <= %d", RAG_MAX_KEYWORDS);
  RAG_CUDA(cudaSetDevice(idx->device));
  rag_batch* bt = &idx->main;
  idx->cur = bt;
  RAG_CHECK(stage_queries(idx, bt, queries, B));
  RAG_CHECK(stage_keywords(idx, bt, B, kw_keys, kw_counts, keyword_limit));
  bt->staged_B = B;
  bt->win_first = 0;
  bt->win_count = B;
  RAG_CUDA(cudaStreamSynchronize(idx->stream));
  return RAG_OK;
}

int rag_stage_window(rag_index* idx, uint32_t first, uint32_t count) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  rag_batch* bt = &idx->main;
  if (count == 0 || (uint64_t)first + count > bt->staged_B)
    return rag_set_error(RAG_ERR_INVALID, "rag_stage_window: [%u,%u) is outside the %u staged queries", first, first + count, bt->staged_B);
  bt->win_first = first;
  bt->win_count = count;
  return RAG_OK;
}

int rag_hybrid_search_staged(rag_index* idx, uint32_t B, const rag_hybrid_opts* o) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  if (!o) return rag_set_error(RAG_ERR_INVALID, "rag_hybrid_search_staged: null options");
  rag_batch* bt = &idx->main;
  if (B == 0 || B != bt->win_count) return rag_set_error(RAG_ERR_STATE, "rag_hybrid_search_staged: B=%u but the staged window holds %u queries", B, bt->win_count);
  if (o->keyword_limit != bt->staged_kw_stride)
    return rag_set_error(RAG_ERR_STATE, "rag_hybrid_search_staged: keyword_limit=%u but staged with %u", o->keyword_limit, bt->staged_kw_stride);
  rag_fuse_args fa;
  fresh_cfg fc;
  uint32_t out_cap = 0;
  RAG_CHECK(hybrid_args(idx, o, &fa, &fc, &out_cap));
  RAG_CUDA(cudaSetDevice(idx->device));
  plan p;
  RAG_CHECK(make_plan(idx, B, o->vector_top_k, o->path, o->slack, o->epsilon, o->flags, &p));
  idx->cur = bt;
  out_layout L;
  RAG_CHECK(ensure_work(idx, bt, B, o->vector_top_k, out_cap, &L));
  // view of the window: inputs shifted, work/output buffers shared
  rag_batch view = *bt;
  view.d_q = bt->d_q + (size_t)bt->win_first * idx->ld;
  view.d_kw = bt->d_kw + (size_t)bt->win_first * bt->staged_kw_stride;
  view.d_kwc = bt->d_kwc + bt->win_first;
  idx->cur = &view;
  int rc = run_pipeline(idx, B, o->vector_top_k, p, fc, fa);
  bt->d_partial = view.d_partial;  // run_pipeline may have grown these through the view
  bt->c_partial = view.c_partial;
  bt->d_qb = view.d_qb;
  bt->c_qb = view.c_qb;
  bt->d_rho_q = view.d_rho_q;
  bt->c_rho_q = view.c_rho_q;
  idx->cur = bt;
  return rc;
}

int rag_fetch_fused(rag_index* idx, uint32_t B, const rag_hybrid_opts* o, rag_fused_out* out) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  if (!o) return rag_set_error(RAG_ERR_INVALID, "rag_fetch_fused: null options");
  rag_batch* bt = &idx->main;
  if (B == 0 || B != bt->win_count) return rag_set_error(RAG_ERR_STATE, "rag_fetch_fused: B does not match the staged window");
  const uint32_t out_cap = o->vector_top_k + o->keyword_limit + o->fresh_limit;
  RAG_CHECK(check_fused_out(o, out, out_cap));
  RAG_CUDA(cudaSetDevice(idx->device));
  const out_layout L = layout_out(B, out_cap, o->vector_top_k);
  if (!bt->d_out || L.total > bt->c_out) return rag_set_error(RAG_ERR_STATE, "rag_fetch_fused: no staged search has run");
  RAG_CHECK(fetch_out(idx, bt, L, false));
  for (uint32_t b = 0; b < B; b++) write_fused(o, out, out_cap, b, bt, L, b);
  return RAG_OK;
}

int rag_sync(rag_index* idx) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  RAG_CUDA(cudaSetDevice(idx->device));
  RAG_CUDA(cudaStreamSynchronize(idx->stream));
  return RAG_OK;
}

// ---- fusion only ---------------------------------------------------------------------------------
int rag_rrf_fuse(rag_index* idx, uint32_t B, const rag_rrf_config* cfg, const uint64_t* vec_keys,
                 const uint8_t* vec_ctype, const uint32_t* vec_counts, uint32_t vec_stride, const uint64_t* kw_keys,
                 const uint32_t* kw_counts, uint32_t kw_stride, rag_fused_out* out) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  if (!cfg || !out || !out->keys || !out->scores || !out->counts || B == 0 || !vec_counts || !kw_counts)
    return rag_set_error(RAG_ERR_INVALID, "rag_rrf_fuse: null argument or empty batch");
  if (vec_stride > RAG_MAX_TOPK || kw_stride > RAG_MAX_KEYWORDS)
    return rag_set_error(RAG_ERR_INVALID, "rag_rrf_fuse: list strides exceed %d / %d", RAG_MAX_TOPK, RAG_MAX_KEYWORDS);
  const uint32_t out_cap = std::max(1u, vec_stride + kw_stride);
  if (out->capacity < out_cap) return rag_set_error(RAG_ERR_INVALID, "rag_fused_out.capacity=%u < %u", out->capacity, out_cap);
  for (uint32_t b = 0; b < B; b++)
    if (vec_counts[b] > vec_stride || kw_counts[b] > kw_stride)
      return rag_set_error(RAG_ERR_INVALID, "rag_rrf_fuse: counts[%u] exceed the list stride", b);
  RAG_CUDA(cudaSetDevice(idx->device));
  rag_batch* bt = &idx->main;
  idx->cur = bt;
  bt->staged_B = bt->win_count = 0;  // the input block is re-carved below
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 15) & ~(size_t)15; return at; };
  const size_t o_vk = take((size_t)B * vec_stride * 8), o_kk = take((size_t)B * kw_stride * 8);
  const size_t o_vc = take((size_t)B * 4), o_kc = take((size_t)B * 4), o_vt = take((size_t)B * vec_stride);
  RAG_CHECK(grow_dev(&bt->d_in, &bt->c_in, o, true));
  RAG_CHECK(grow_pinned(&bt->h_in, &bt->c_hin, o));
  if (vec_stride && vec_keys) memcpy(bt->h_in + o_vk, vec_keys, (size_t)B * vec_stride * 8);
  if (kw_stride && kw_keys) memcpy(bt->h_in + o_kk, kw_keys, (size_t)B * kw_stride * 8);
  memcpy(bt->h_in + o_vc, vec_counts, (size_t)B * 4);
  memcpy(bt->h_in + o_kc, kw_counts, (size_t)B * 4);
  if (vec_ctype && vec_stride) memcpy(bt->h_in + o_vt, vec_ctype, (size_t)B * vec_stride);
  RAG_CUDA(cudaMemcpyAsync(bt->d_in, bt->h_in, o, cudaMemcpyHostToDevice, idx->stream));
  const out_layout L = layout_out(B, out_cap, 1);
  RAG_CHECK(grow_dev(&bt->d_out, &bt->c_out, L.total, false));
  RAG_CHECK(grow_pinned(&bt->h_out, &bt->c_hout, L.total));
  bind_out(bt, L);
  RAG_CHECK(k5_rrf_only_launch(idx, B, cfg, (const uint64_t*)(bt->d_in + o_vk), vec_ctype ? bt->d_in + o_vt : nullptr,
                               (const uint32_t*)(bt->d_in + o_vc), vec_stride, (const uint64_t*)(bt->d_in + o_kk),
                               (const uint32_t*)(bt->d_in + o_kc), kw_stride, out_cap));
  RAG_CHECK(fetch_out(idx, bt, L, false));
  for (uint32_t b = 0; b < B; b++) {
    const uint32_t cnt = ((const uint32_t*)(bt->h_out + L.cnt))[b];
    const size_t base = (size_t)b * out->capacity, sb = (size_t)b * out_cap;
    for (uint32_t i = 0; i < out->capacity; i++) {
      const bool live = i < cnt;
      out->keys[base + i] = live ? ((const uint64_t*)(bt->h_out + L.keys))[sb + i] : ~0ull;
      out->scores[base + i] = live ? ((const double*)(bt->h_out + L.scores))[sb + i] : -INFINITY;
      if (out->source) out->source[base + i] = live ? (bt->h_out + L.src)[sb + i] : 0;
      if (out->content_type) out->content_type[base + i] = live ? (bt->h_out + L.ct)[sb + i] : 0;
    }
    out->counts[b] = cnt;
    if (out->used_rrf) out->used_rrf[b] = 1;
    if (out->certified) out->certified[b] = 1;
  }
  return RAG_OK;
}

// ---- memory ------------------------------------------------------------------------------------
int rag_memory_retrieve(rag_index* idx, const float* queries, uint32_t B, const rag_memory_opts* o, rag_memory_out* out) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  if (!queries || !o || !out || !out->ids || !out->scores || !out->counts || B == 0)
    return rag_set_error(RAG_ERR_INVALID, "rag_memory_retrieve: null argument or empty batch");
  // similarityTopK: limit * 2 — src/lib/memory/store.ts:112 (or the caller's own value)
  const uint32_t k = o->similarity_top_k ? o->similarity_top_k : o->limit * 2;
  if (o->limit == 0 || o->limit > RAG_MAX_TOPK || k > RAG_MAX_TOPK)
    return rag_set_error(RAG_ERR_INVALID, "limit must be in 1..%d (similarityTopK = %u must not exceed %d)", RAG_MAX_TOPK / 2, k, RAG_MAX_TOPK);
  if (idx->rows == 0) return rag_set_error(RAG_ERR_STATE, "search on an empty index");
  RAG_CUDA(cudaSetDevice(idx->device));
  const uint32_t out_cap = o->limit;
  plan p;
  RAG_CHECK(make_plan(idx, B, k, o->path, 0, 0.0, o->flags, &p));
  rag_batch* bt = &idx->main;
  idx->cur = bt;
  bt->staged_B = bt->win_count = 0;
  out_layout L;
  RAG_CHECK(stage_queries(idx, bt, queries, B));
  RAG_CHECK(ensure_work(idx, bt, B, k, out_cap, &L));
  rag_fuse_args fa = {};
  fa.mode = 1;
  fa.out_cap = out_cap;
  fa.mem_limit = o->limit;
  fa.mem_min_relevance = o->min_relevance;
  fresh_cfg fc = {o->now_ms, o->time_decay_factor, o->frequency_bonus};
  fresh_defaults(&fc.decay, &fc.bonus);
  RAG_CHECK(run_pipeline(idx, B, k, p, fc, fa));
  RAG_CHECK(fetch_out(idx, bt, L, true));
  writer_fn write = [&](uint32_t ub, const rag_batch* src, const out_layout& SL, uint32_t lb) {
    const uint32_t cnt = ((const uint32_t*)(src->h_out + SL.cnt))[lb];
    const size_t sb = (size_t)lb * out_cap, base = (size_t)ub * o->limit;
    for (uint32_t i = 0; i < o->limit; i++) {
      const bool live = i < cnt;
      out->ids[base + i] = live ? ((const uint64_t*)(src->h_out + SL.keys))[sb + i] : ~0ull;
      out->scores[base + i] = live ? ((const double*)(src->h_out + SL.scores))[sb + i] : -INFINITY;
      if (out->relevance) out->relevance[base + i] = live ? ((const double*)(src->h_out + SL.aux0))[sb + i] : -INFINITY;
      if (out->freshness) out->freshness[base + i] = live ? ((const double*)(src->h_out + SL.aux1))[sb + i] : 0.0;
    }
    out->counts[ub] = cnt;
  };
  std::vector<uint32_t> sel;
  const uint8_t* cert = bt->h_out + L.cert;
  for (uint32_t b = 0; b < B; b++) {
    write(b, bt, L, b);
    if (!cert[b]) sel.push_back(b);
  }
  if (!sel.empty()) RAG_CHECK(escalate(idx, sel, k, p, fc, fa, out_cap, true, write));
  return RAG_OK;
}

int rag_freshness_scores(rag_index* idx, uint64_t n, const double* confidence, const int32_t* access_count,
                         const int64_t* last_access_ms, int64_t now_ms, double decay, double bonus, double* out_scores) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  if (n == 0) return RAG_OK;
  if (!confidence || !access_count || !last_access_ms || !out_scores)
    return rag_set_error(RAG_ERR_INVALID, "rag_freshness_scores: null argument");
  fresh_defaults(&decay, &bonus);
  RAG_CUDA(cudaSetDevice(idx->device));
  uint8_t* d = nullptr;
  size_t cap = 0;
  RAG_CHECK(grow_dev(&d, &cap, n * 28, false));
  double* d_conf = (double*)d;
  int64_t* d_last = (int64_t*)(d + n * 8);
  double* d_out = (double*)(d + n * 16);
  int32_t* d_acc = (int32_t*)(d + n * 24);
  int rc = RAG_OK;
  cudaError_t e;
  if ((e = cudaMemcpyAsync(d_conf, confidence, n * 8, cudaMemcpyHostToDevice, idx->stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(d_last, last_access_ms, n * 8, cudaMemcpyHostToDevice, idx->stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(d_acc, access_count, n * 4, cudaMemcpyHostToDevice, idx->stream)) != cudaSuccess)
    rc = rag_set_error(RAG_ERR_CUDA, "rag_freshness_scores H2D: %s", cudaGetErrorString(e));
  if (rc == RAG_OK) rc = k5_freshness_launch(idx, n, d_conf, d_acc, d_last, now_ms, decay, bonus, d_out);
  if (rc == RAG_OK) {
    e = cudaMemcpyAsync(out_scores, d_out, n * 8, cudaMemcpyDeviceToHost, idx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(idx->stream);
    if (e != cudaSuccess) rc = rag_set_error(RAG_ERR_CUDA, "rag_freshness_scores D2H: %s", cudaGetErrorString(e));
  }
  cudaFree(d);
  return rc;
}

// ---- diagnostics -----------------------------------------------------------------------------------
int rag_debug_tensor_scores(rag_index* idx, const float* queries, uint32_t B, float* out_scores) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  if (!queries || !out_scores || B == 0) return rag_set_error(RAG_ERR_INVALID, "rag_debug_tensor_scores: null argument");
  if (idx->rows == 0) return rag_set_error(RAG_ERR_STATE, "search on an empty index");
  RAG_CUDA(cudaSetDevice(idx->device));
  plan p;
  RAG_CHECK(make_plan(idx, B, 8, RAG_PATH_TENSOR, 0, 0.0, 0, &p));
  rag_batch* bt = &idx->main;
  idx->cur = bt;
  bt->staged_B = bt->win_count = 0;
  RAG_CHECK(stage_queries(idx, bt, queries, B));
  float* d = nullptr;
  size_t cap = 0;
  RAG_CHECK(grow_dev(&d, &cap, (size_t)B * idx->rows * 4, true));
  uint32_t parts = 0;
  int rc = k2_plan(idx, B, p.kp, &parts);
  if (rc == RAG_OK) rc = grow_dev(&bt->d_partial, &bt->c_partial, (size_t)B * parts * p.kp * 8, false);
  if (rc == RAG_OK) {
    k2_set_debug(idx, d);
    rc = k2_launch(idx, B, p.kp, parts);
    k2_set_debug(idx, nullptr);
  }
  if (rc == RAG_OK) {
    cudaError_t e = cudaMemcpyAsync(out_scores, d, (size_t)B * idx->rows * 4, cudaMemcpyDeviceToHost, idx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(idx->stream);
    if (e != cudaSuccess) rc = rag_set_error(RAG_ERR_CUDA, "rag_debug_tensor_scores: %s", cudaGetErrorString(e));
  }
  cudaFree(d);
  return rc;
}

// K2 + K3 only: the K' packed candidate keys per query next to the score matrix they were selected from
// (one launch, so both views come from the same accumulators) — lets a test check the fused selection
// exactly, ties included, without going through the fp64 rescoring
int rag_debug_tensor_candidates(rag_index* idx, const float* queries, uint32_t B, uint32_t kp, float* out_scores,
                                uint64_t* out_keys) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  if (!queries || !out_scores || !out_keys || B == 0 || kp == 0 || kp > RAG_MAX_CANDIDATES)
    return rag_set_error(RAG_ERR_INVALID, "rag_debug_tensor_candidates: bad argument");
  if (idx->rows == 0) return rag_set_error(RAG_ERR_STATE, "search on an empty index");
  if (!k2_available(idx)) return rag_set_error(RAG_ERR_UNSUPPORTED, "tensor path (K2) is not available for this index");
  RAG_CUDA(cudaSetDevice(idx->device));
  rag_batch* bt = &idx->main;
  idx->cur = bt;
  bt->staged_B = bt->win_count = 0;
  RAG_CHECK(stage_queries(idx, bt, queries, B));
  float* d = nullptr;
  size_t cap = 0;
  RAG_CHECK(grow_dev(&d, &cap, (size_t)B * idx->rows * 4, true));
  uint32_t parts = 0;
  int rc = k2_plan(idx, B, kp, &parts);
  if (rc == RAG_OK) rc = grow_dev(&bt->d_partial, &bt->c_partial, (size_t)B * parts * kp * 8, false);
  if (rc == RAG_OK) rc = grow_dev(&bt->d_cand, &bt->c_cand, (size_t)B * RAG_MAX_CANDIDATES * 8, false);
  if (rc == RAG_OK) {
    k2_set_debug(idx, d);
    rc = k2_launch(idx, B, kp, parts);
    k2_set_debug(idx, nullptr);
  }
  if (rc == RAG_OK) rc = k3_launch(idx, B, kp, parts);
  if (rc == RAG_OK) {
    cudaError_t e = cudaMemcpyAsync(out_scores, d, (size_t)B * idx->rows * 4, cudaMemcpyDeviceToHost, idx->stream);
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync(out_keys, (size_t)kp * 8, bt->d_cand, (size_t)RAG_MAX_CANDIDATES * 8, (size_t)kp * 8, B,
                            cudaMemcpyDeviceToHost, idx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(idx->stream);
    if (e != cudaSuccess) rc = rag_set_error(RAG_ERR_CUDA, "rag_debug_tensor_candidates: %s", cudaGetErrorString(e));
  }
  cudaFree(d);
  return rc;
}

double rag_index_row_residual(rag_index* idx) {
  if (!idx) return (double)rag_set_error(RAG_ERR_INVALID, "null index handle");
  RAG_LOCK(idx);
  if (cudaSetDevice(idx->device) != cudaSuccess) return (double)rag_set_error(RAG_ERR_CUDA, "cudaSetDevice failed");
  const int rc = refresh_rho_x(idx);
  return rc != RAG_OK ? (double)rc : (double)idx->rho_x;
}

// ---- measurement ---------------------------------------------------------------------------------
int rag_timer_start(rag_index* idx) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  RAG_CUDA(cudaSetDevice(idx->device));
  RAG_CUDA(cudaEventRecord(idx->ev0, idx->stream));
  return RAG_OK;
}

int rag_timer_stop(rag_index* idx, float* elapsed_ms) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  RAG_CUDA(cudaSetDevice(idx->device));
  RAG_CUDA(cudaEventRecord(idx->ev1, idx->stream));
  RAG_CUDA(cudaEventSynchronize(idx->ev1));
  float ms = 0.f;
  RAG_CUDA(cudaEventElapsedTime(&ms, idx->ev0, idx->ev1));
  if (elapsed_ms) *elapsed_ms = ms;
  return RAG_OK;
}

uint64_t rag_launch_count(const rag_index* idx) { return idx ? idx->launches : 0; }

// device-side totals of the fusion kernel since the last call: how many queries it finished and how many of them
// were certified by the scoring pass that produced them (escalated re-runs count as their own queries)
int rag_certified_totals(rag_index* idx, uint64_t* certified, uint64_t* queries) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  RAG_CUDA(cudaSetDevice(idx->device));
  unsigned long long h[2] = {0, 0};
  RAG_CUDA(cudaMemcpyAsync(h, idx->d_counters, sizeof(h), cudaMemcpyDeviceToHost, idx->stream));
  RAG_CUDA(cudaMemsetAsync(idx->d_counters, 0, sizeof(h), idx->stream));
  RAG_CUDA(cudaStreamSynchronize(idx->stream));
  if (certified) *certified = h[0];
  if (queries) *queries = h[1];
  return RAG_OK;
}

int rag_profile_enable(rag_index* idx, int on) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  RAG_CUDA(cudaSetDevice(idx->device));
  if (on && !idx->prof_spans) {
    idx->prof_spans = new (std::nothrow) rag_prof_span[kProfSpans];
    if (!idx->prof_spans) return rag_set_error(RAG_ERR_NOMEM, "out of host memory");
    idx->prof_cap = 0;
    for (uint32_t i = 0; i < kProfSpans; i++) {
      RAG_CUDA(cudaEventCreate(&idx->prof_spans[i].a));
      RAG_CUDA(cudaEventCreate(&idx->prof_spans[i].b));
      idx->prof_cap = i + 1;
    }
  }
  if (!on) prof_drain(idx);
  idx->prof_on = on != 0;
  return RAG_OK;
}

int rag_profile_read(rag_index* idx, float ms[RAG_PROF_CLASSES], uint32_t counts[RAG_PROF_CLASSES]) {
  RAG_CHECK(check_handle(idx));
  RAG_LOCK(idx);
  RAG_CUDA(cudaSetDevice(idx->device));
  prof_drain(idx);
  for (int i = 0; i < RAG_PROF_CLASSES; i++) {
    if (ms) ms[i] = idx->prof_ms[i];
    if (counts) counts[i] = idx->prof_cnt[i];
    idx->prof_ms[i] = 0.f;
    idx->prof_cnt[i] = 0;
  }
  return RAG_OK;
}

void* rag_host_alloc(uint64_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault) != cudaSuccess) {
    rag_set_error(RAG_ERR_NOMEM, "cudaHostAlloc(%llu) failed", (unsigned long long)bytes);
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void rag_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"
