// batcher.cu — §8f N4: a micro-batching front end for concurrent batch-1 callers.
//
// The reference serves one query per request (hybridSearch is called once per tool call /
// buildContext): every caller owns a whole corpus scan. On a B200 the scan is
// HBM-bound at batch 1 and tensor-bound at batch >= ~64, so requests that arrive within a few
// hundred microseconds of each other are worth two orders of magnitude more throughput when
// they share one pass. The batcher collects concurrent rag_batcher_submit() calls (any threads)
// for at most `max_wait_us` or until `max_batch` are waiting, runs ONE rag_hybrid_search over
// them on its own worker thread (the only thread that touches the rag_index, as the handle
// contract requires), and hands every caller exactly the result it would have got alone:
// batching never changes a result (tests/test_gpu_batcher.py).
//
// One batcher = one call-site class (fixed HybridSearchOptions: search_knowledge, deep_search, ...),
// because vectorTopK / keywordLimit / RRF config are per-launch parameters.
// Host code only.
#include "common.cuh"

#include <string.h>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

struct rag_batcher_req {
  const float* q;
  const uint64_t* kw;
  uint32_t kwc;
  rag_fused_out* out;
  int rc = RAG_OK;
  bool done = false;
  std::string err;
  std::chrono::steady_clock::time_point t_enq;
};

struct rag_batcher {
  rag_index* idx = nullptr;
  rag_batcher_desc desc;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::deque<rag_batcher_req*> queue;
  bool stop = false;
  std::thread worker;
  uint64_t n_batches = 0, n_queries = 0, max_seen = 0;
  // batch staging (worker only)
  float* h_q = nullptr;  // pinned [max_batch][dim]
  std::vector<uint64_t> kw_keys;
  std::vector<uint32_t> kw_counts;
  std::vector<uint64_t> keys, vec_ids;
  std::vector<double> scores, vec_scores;
  std::vector<uint8_t> source, ctype, used_rrf, certified;
  std::vector<uint32_t> counts, vec_counts;
};

namespace {

void run_batch(rag_batcher* b, std::vector<rag_batcher_req*>& reqs) {
  const uint32_t B = (uint32_t)reqs.size(), dim = b->idx->dim;
  const rag_hybrid_opts& o = b->desc.opts;
  const uint32_t kl = o.keyword_limit, k = o.vector_top_k, cap = k + kl + o.fresh_limit;
  for (uint32_t i = 0; i < B; i++) {
    memcpy(b->h_q + (size_t)i * dim, reqs[i]->q, (size_t)dim * sizeof(float));
    const uint32_t c = reqs[i]->kw ? std::min(reqs[i]->kwc, kl) : 0u;
    if (c) memcpy(b->kw_keys.data() + (size_t)i * kl, reqs[i]->kw, (size_t)c * 8);
    b->kw_counts[i] = c;
  }
  rag_fused_out out = {cap, b->keys.data(), b->scores.data(), b->source.data(), b->ctype.data(), b->counts.data(),
                       b->used_rrf.data(), b->vec_ids.data(), b->vec_scores.data(), b->vec_counts.data(), b->certified.data()};
  const int rc = rag_hybrid_search(b->idx, b->h_q, B, &o, b->kw_keys.data(), b->kw_counts.data(), &out);
  const std::string err = rc == RAG_OK ? std::string() : std::string(rag_last_error());
  for (uint32_t i = 0; i < B; i++) {
    rag_batcher_req* r = reqs[i];
    r->rc = rc;
    r->err = err;
    if (rc == RAG_OK) {
      rag_fused_out* d = r->out;
      const uint32_t n = std::min(d->capacity, cap);
      memcpy(d->keys, b->keys.data() + (size_t)i * cap, (size_t)n * 8);
      memcpy(d->scores, b->scores.data() + (size_t)i * cap, (size_t)n * 8);
      if (d->source) memcpy(d->source, b->source.data() + (size_t)i * cap, n);
      if (d->content_type) memcpy(d->content_type, b->ctype.data() + (size_t)i * cap, n);
      d->counts[0] = b->counts[i];
      if (d->used_rrf) d->used_rrf[0] = b->used_rrf[i];
      if (d->certified) d->certified[0] = b->certified[i];
      if (d->vec_ids && d->vec_scores && d->vec_counts) {
        memcpy(d->vec_ids, b->vec_ids.data() + (size_t)i * k, (size_t)k * 8);
        memcpy(d->vec_scores, b->vec_scores.data() + (size_t)i * k, (size_t)k * 8);
        d->vec_counts[0] = b->vec_counts[i];
      }
    }
  }
}

void worker_main(rag_batcher* b) {
  std::vector<rag_batcher_req*> batch;
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(b->mu);
      b->cv_work.wait(lk, [&] { return b->stop || !b->queue.empty(); });
      if (b->stop && b->queue.empty()) return;
      // the OLDEST waiting request opens the collection window (requests that queued up behind a running batch have
      // already waited: they go out at once); leave early once the batch is full
      const auto deadline = b->queue.front()->t_enq + std::chrono::microseconds(b->desc.max_wait_us);
      b->cv_work.wait_until(lk, deadline, [&] { return b->stop || b->queue.size() >= b->desc.max_batch; });
      batch.clear();
      while (!b->queue.empty() && batch.size() < b->desc.max_batch) {
        batch.push_back(b->queue.front());
        b->queue.pop_front();
      }
    }
    run_batch(b, batch);
    {
      std::lock_guard<std::mutex> lk(b->mu);
      for (rag_batcher_req* r : batch) r->done = true;
      b->n_batches++;
      b->n_queries += batch.size();
      b->max_seen = std::max<uint64_t>(b->max_seen, batch.size());
    }
    b->cv_done.notify_all();
  }
}

}  // namespace

extern "C" {

int rag_batcher_create(rag_index* idx, const rag_batcher_desc* d, rag_batcher** out) {
  if (!idx || !d || !out) return rag_set_error(RAG_ERR_INVALID, "rag_batcher_create: null argument");
  *out = nullptr;
  if (d->max_batch == 0 || d->max_batch > 4096) return rag_set_error(RAG_ERR_INVALID, "max_batch must be in 1..4096");
  const rag_hybrid_opts& o = d->opts;
  if (o.vector_top_k == 0 || o.vector_top_k > RAG_MAX_TOPK || o.keyword_limit > RAG_MAX_KEYWORDS || o.fresh_limit > RAG_MAX_FRESH)
    return rag_set_error(RAG_ERR_INVALID, "rag_batcher_create: bad hybrid options");
  rag_batcher* b = new (std::nothrow) rag_batcher();
  if (!b) return rag_set_error(RAG_ERR_NOMEM, "out of host memory");
  b->idx = idx;
  b->desc = *d;
  const uint32_t B = d->max_batch, k = o.vector_top_k, cap = k + o.keyword_limit + o.fresh_limit;
  b->h_q = (float*)rag_host_alloc((uint64_t)B * idx->dim * sizeof(float));
  if (!b->h_q) { delete b; return RAG_ERR_NOMEM; }
  b->kw_keys.assign((size_t)B * std::max(1u, o.keyword_limit), 0);
  b->kw_counts.assign(B, 0);
  b->keys.resize((size_t)B * cap); b->scores.resize((size_t)B * cap);
  b->source.resize((size_t)B * cap); b->ctype.resize((size_t)B * cap);
  b->counts.resize(B); b->used_rrf.resize(B); b->certified.resize(B);
  b->vec_ids.resize((size_t)B * k); b->vec_scores.resize((size_t)B * k); b->vec_counts.resize(B);
  b->worker = std::thread(worker_main, b);
  *out = b;
  return RAG_OK;
}

// Blocking; callable from any number of threads. `query` is one [dim] fp32 vector, `kw_keys` the
// request's keyword hits as fusion keys in rank order; `out` is shaped for ONE query (capacity >=
// vector_top_k + keyword_limit + fresh_limit; counts/used_rrf/certified/vec_counts have 1 entry).
int rag_batcher_submit(rag_batcher* b, const float* query, const uint64_t* kw_keys, uint32_t kw_count, rag_fused_out* out) {
  if (!b || !query || !out || !out->keys || !out->scores || !out->counts)
    return rag_set_error(RAG_ERR_INVALID, "rag_batcher_submit: null argument");
  const rag_hybrid_opts& o = b->desc.opts;
  if (out->capacity < o.vector_top_k + o.keyword_limit + o.fresh_limit)
    return rag_set_error(RAG_ERR_INVALID, "rag_batcher_submit: rag_fused_out.capacity too small");
  if (kw_count > o.keyword_limit) return rag_set_error(RAG_ERR_INVALID, "rag_batcher_submit: kw_count exceeds keyword_limit");
  rag_batcher_req r;
  r.q = query;
  r.kw = kw_keys;
  r.kwc = kw_count;
  r.out = out;
  {
    std::unique_lock<std::mutex> lk(b->mu);
    if (b->stop) return rag_set_error(RAG_ERR_STATE, "rag_batcher_submit: batcher is shutting down");
    r.t_enq = std::chrono::steady_clock::now();
    b->queue.push_back(&r);
    b->cv_work.notify_one();
    b->cv_done.wait(lk, [&] { return r.done; });
  }
  if (r.rc != RAG_OK) return rag_set_error(r.rc, "%s", r.err.c_str());
  return RAG_OK;
}

int rag_batcher_stats(rag_batcher* b, uint64_t* batches, uint64_t* queries, uint64_t* largest_batch) {
  if (!b) return rag_set_error(RAG_ERR_INVALID, "null batcher");
  std::lock_guard<std::mutex> lk(b->mu);
  if (batches) *batches = b->n_batches;
  if (queries) *queries = b->n_queries;
  if (largest_batch) *largest_batch = b->max_seen;
  return RAG_OK;
}

void rag_batcher_destroy(rag_batcher* b) {
  if (!b) return;
  {
    std::lock_guard<std::mutex> lk(b->mu);
    b->stop = true;
  }
  b->cv_work.notify_all();
  if (b->worker.joinable()) b->worker.join();
  rag_host_free(b->h_q);
  delete b;
}

}  // extern "C"
