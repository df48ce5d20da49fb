// batcher.cu — §8f N4: a micro-batching front end for concurrent batch-1 callers.
//
// The reference serves one query per request (hybridSearch is called once per tool call /
// buildContext): every caller owns a whole corpus scan. On a B200 the scan is
// HBM-bound at batch 1 and tensor-bound at batch >= ~64, so requests that arrive within a few
// hundred microseconds of each other are worth two orders of magnitude more throughput when
// they share one pass. The batcher collects concurrent rag_batcher_submit() calls (any threads)
// into an OPEN batch, runs ONE rag_hybrid_search per batch, and hands every caller exactly the
// result it would have got alone: batching never changes a result (tests/test_gpu_batcher.py).
//
// Pipeline (round 2; the first version did everything on one worker thread and lost 38% of the
// tensor path's throughput at 1024 request threads — profiles/r02_batcher_load.md):
//   * the SUBMITTERS stage their own inputs: a request takes a slot of the open batch under the
//     queue lock, copies its query (6 KB at D=1536) and keyword list into the batch's pinned block
//     outside the lock, and copies its own result out afterwards — the 6 MB of a 1024-query batch
//     move on as many cores as there are callers instead of on the worker;
//   * TWO workers take turns: one holds the turn while it closes the open batch and runs it on the
//     GPU; the other meanwhile wakes the previous batch's callers. The open batch keeps collecting
//     for as long as the GPU is busy (that is what sizes the batches under load); when the GPU is
//     idle it goes out once the arrivals pause (20-50 us without a new request), at the latest `max_wait_us` after
//     its first request;
//   * callers sleep on 64 wake groups (16 consecutive slots each), not on one condition variable:
//     completing a batch is at most 64 short broadcasts instead of 1024 threads fighting for one mutex;
//   * four batch buffers circulate (open, on the GPU, being read out, spare); a submitter that
//     finds none free waits (back-pressure).
// The workers are the only threads that touch the rag_index; calls on the handle are serialised
// by the library anyway (per-handle mutex).
//
// Two ways in: rag_batcher_submit blocks its thread until the answer is there (a server with a thread per request);
// rag_batcher_submit_async takes a slot, stages the inputs and returns — a worker copies the result into the caller's
// arrays and calls `done` — so ONE thread (Node's event loop) can keep thousands of requests in flight without parking a
// pool thread per request (libuv's default pool is 4 threads: blocking submits could never form a batch above 4).
//
// One batcher = one call-site class (fixed HybridSearchOptions: search_knowledge, deep_search, ...),
// because vectorTopK / keywordLimit / RRF config are per-launch parameters.
// Host code only.
#include "common.cuh"

#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {
constexpr uint32_t kWakeGroups = 64;
constexpr uint32_t kGroupSlots = 16;   // consecutive slots that share a wake group: a 64-query batch is woken by 4 broadcasts
constexpr int kBuffers = 4;
constexpr int kWorkers = 2;
}  // namespace

struct rag_batcher_req {
  bool done = false;  // guarded by the wake group's mutex
};
// a request that nobody waits for: where its result goes and whom to tell (rag_batcher_submit_async)
struct rag_async_req {
  rag_batcher_done_fn done = nullptr;   // null = the slot belongs to a blocking caller
  void* user = nullptr;
  rag_fused_out out;
};

// One batch: inputs written by the submitters (slot by slot), outputs written by rag_hybrid_search, read by the submitters.
struct rag_batch_buf {
  float* h_q = nullptr;  // pinned [max_batch][dim]
  std::vector<uint64_t> kw_keys;
  std::vector<uint32_t> kw_counts;
  std::vector<uint64_t> keys, vec_ids;
  std::vector<double> scores, vec_scores;
  std::vector<uint8_t> source, ctype, used_rrf, certified;
  std::vector<uint32_t> counts, vec_counts;
  std::vector<rag_batcher_req*> reqs;
  std::vector<rag_async_req> areqs;
  uint32_t count = 0;                  // slots handed out (under rag_batcher::mu)
  std::atomic<uint32_t> staged{0};     // slots whose inputs are in place
  std::atomic<uint32_t> readers{0};    // callers that have not copied their result out yet
  std::chrono::steady_clock::time_point t_first, t_last;   // arrival of the first / latest request (under rag_batcher::mu)
  int rc = RAG_OK;
  std::string err;
};

struct rag_wake_group {
  std::mutex mu;
  std::condition_variable cv;
};

struct rag_batcher {
  rag_index* idx = nullptr;
  rag_batcher_desc desc;
  std::mutex mu;                        // open / full / free_bufs / stop / stats
  std::condition_variable cv_work;      // the worker holding the turn: a first request, a full batch, stop
  std::condition_variable cv_free;      // submitters waiting for a batch buffer
  std::mutex turn;                      // one worker at a time closes a batch and runs it
  rag_batch_buf bufs[kBuffers];
  rag_batch_buf* open = nullptr;        // the batch that takes new requests
  std::deque<rag_batch_buf*> full;      // batches that filled up while the GPU was busy
  std::vector<rag_batch_buf*> free_bufs;
  rag_wake_group groups[kWakeGroups];
  bool stop = false;
  std::atomic<int> waking{0};           // batches whose callers are being woken right now
  std::atomic<uint32_t> active{0};      // rag_batcher_submit calls between entry and return: destroy waits for them
  std::thread workers[kWorkers];
  uint64_t n_batches = 0, n_queries = 0, max_seen = 0;
};

namespace {

// counts a submitter in and out, whichever way it returns
struct active_call {
  rag_batcher* b;
  explicit active_call(rag_batcher* b_) : b(b_) { b->active.fetch_add(1, std::memory_order_acq_rel); }
  ~active_call() { b->active.fetch_sub(1, std::memory_order_acq_rel); }
};

void run_batch(rag_batcher* b, rag_batch_buf* bb) {
  const uint32_t B = bb->count;
  const rag_hybrid_opts& o = b->desc.opts;
  const uint32_t cap = o.vector_top_k + o.keyword_limit + o.fresh_limit;
  // every submitter's copy has landed (they are a few microseconds each)
  while (bb->staged.load(std::memory_order_acquire) != B) std::this_thread::yield();
  rag_fused_out out = {cap, bb->keys.data(), bb->scores.data(), bb->source.data(), bb->ctype.data(), bb->counts.data(),
                       bb->used_rrf.data(), bb->vec_ids.data(), bb->vec_scores.data(), bb->vec_counts.data(), bb->certified.data()};
  bb->rc = rag_hybrid_search(b->idx, bb->h_q, B, &o, bb->kw_keys.data(), bb->kw_counts.data(), &out);
  bb->err = bb->rc == RAG_OK ? std::string() : std::string(rag_last_error());
}

// one request's result: from the batch's arrays into the caller's (shaped for ONE query)
void copy_out(const rag_batcher* b, const rag_batch_buf* bb, uint32_t slot, rag_fused_out* out) {
  const rag_hybrid_opts& o = b->desc.opts;
  const uint32_t k = o.vector_top_k, cap = k + o.keyword_limit + o.fresh_limit;
  const uint32_t n = std::min(out->capacity, cap);
  memcpy(out->keys, bb->keys.data() + (size_t)slot * cap, (size_t)n * 8);
  memcpy(out->scores, bb->scores.data() + (size_t)slot * cap, (size_t)n * 8);
  if (out->source) memcpy(out->source, bb->source.data() + (size_t)slot * cap, n);
  if (out->content_type) memcpy(out->content_type, bb->ctype.data() + (size_t)slot * cap, n);
  out->counts[0] = bb->counts[slot];
  if (out->used_rrf) out->used_rrf[0] = bb->used_rrf[slot];
  if (out->certified) out->certified[0] = bb->certified[slot];
  if (out->vec_ids && out->vec_scores && out->vec_counts) {
    memcpy(out->vec_ids, bb->vec_ids.data() + (size_t)slot * k, (size_t)k * 8);
    memcpy(out->vec_scores, bb->vec_scores.data() + (size_t)slot * k, (size_t)k * 8);
    out->vec_counts[0] = bb->vec_counts[slot];
  }
}

// a request is through with the batch's arrays; the last one hands the buffer back
void release_reader(rag_batcher* b, rag_batch_buf* bb) {
  if (bb->readers.fetch_sub(1, std::memory_order_acq_rel) != 1) return;
  bb->staged.store(0, std::memory_order_relaxed);
  {
    std::lock_guard<std::mutex> lk(b->mu);
    bb->count = 0;
    b->free_bufs.push_back(bb);
  }
  b->cv_free.notify_all();  // every waiting submitter: the buffer becomes the open batch and has room for all of them
}

// wake the batch's blocking callers group by group (the flag of every request set under its group's mutex), then serve
// the requests nobody waits for: result into the caller's arrays, `done` called from here (a worker thread)
void complete_batch(rag_batcher* b, rag_batch_buf* bb) {
  const uint32_t B = bb->count;
  bb->readers.store(B, std::memory_order_release);
  uint32_t n_async = 0;
  for (uint32_t g = 0; g < kWakeGroups && g * kGroupSlots < B; g++) {
    bool any = false;
    {
      std::lock_guard<std::mutex> lk(b->groups[g].mu);
      for (uint32_t s0 = g * kGroupSlots; s0 < B; s0 += kWakeGroups * kGroupSlots)
        for (uint32_t s = s0; s < s0 + kGroupSlots && s < B; s++) {
          if (bb->areqs[s].done) { n_async++; continue; }
          bb->reqs[s]->done = true;
          any = true;
        }
    }
    if (any) b->groups[g].cv.notify_all();
  }
  if (!n_async) return;
  const int rc = bb->rc;
  const std::string err = bb->err;   // the buffer may be recycled by the time the last callback runs
  for (uint32_t s = 0; s < B && n_async; s++) {
    rag_async_req a = bb->areqs[s];
    if (!a.done) continue;
    n_async--;
    if (rc == RAG_OK) copy_out(b, bb, s, &a.out);
    bb->areqs[s].done = nullptr;
    release_reader(b, bb);            // after the last of these `bb` belongs to the submitters again
    a.done(a.user, rc, rc == RAG_OK ? "" : err.c_str());
  }
}

void worker_main(rag_batcher* b) {
  for (;;) {
    rag_batch_buf* mine = nullptr;
    {
      std::unique_lock<std::mutex> my_turn(b->turn);
      {
        std::unique_lock<std::mutex> lk(b->mu);
        b->cv_work.wait(lk, [&] { return b->stop || !b->full.empty() || (b->open && b->open->count > 0); });
        if (b->full.empty() && !(b->open && b->open->count > 0)) return;  // stop, and nothing left to run
        // The open batch's FIRST request opens the collection window (at most max_wait_us; a batch that has been
        // collecting behind a busy GPU is past it and goes out at once). Inside the window the batch goes out as soon as
        // the arrivals pause for `gap`: callers that were answered together come back together, within tens of
        // microseconds of each other, and splitting them over two corpus passes costs a whole pass (a batch below ~256
        // queries is HBM-bound: the pass costs the same for 30 queries as for 250) — measured with 64 / 256 closed-loop
        // callers: 39k / 79k requests/s when the window simply expired after 100 us, half-full.
        const auto gap = std::chrono::microseconds(std::min<uint32_t>(50u, std::max<uint32_t>(20u, b->desc.max_wait_us / 8)));
        for (;;) {
          if (b->stop || !b->full.empty() || b->open->count >= b->desc.max_batch) break;
          const auto now = std::chrono::steady_clock::now();
          const auto cap = b->open->t_first + std::chrono::microseconds(b->desc.max_wait_us);
          if (now >= cap) break;
          // a finished batch's callers are still being woken: they are about to come back, the pause rule waits for them
          const bool burst = b->waking.load(std::memory_order_acquire) > 0;
          const auto until = std::min(cap, (burst ? now : b->open->t_last) + gap);
          if (!burst && now >= until) break;
          b->cv_work.wait_until(lk, until);
        }
        if (!b->full.empty()) {
          mine = b->full.front();
          b->full.pop_front();
        } else {
          mine = b->open;
          b->open = nullptr;
        }
      }
      run_batch(b, mine);
      b->waking.fetch_add(1, std::memory_order_release);
    }  // the turn passes on: the other worker closes and launches the next batch while this one wakes its callers
    {
      std::lock_guard<std::mutex> lk(b->mu);
      b->n_batches++;
      b->n_queries += mine->count;
      b->max_seen = std::max<uint64_t>(b->max_seen, mine->count);
    }
    complete_batch(b, mine);
    b->waking.fetch_sub(1, std::memory_order_release);
  }
}

}  // namespace

extern "C" {

void rag_batcher_destroy(rag_batcher* b);

int rag_batcher_create(rag_index* idx, const rag_batcher_desc* d, rag_batcher** out) {
  if (!idx || !d || !out) return rag_set_error(RAG_ERR_INVALID, "rag_batcher_create: null argument");
  *out = nullptr;
  if (d->max_batch == 0 || d->max_batch > 4096) return rag_set_error(RAG_ERR_INVALID, "max_batch must be in 1..4096");
  // a row-sharded index needs every rank to issue the SAME sequence of calls (the exchange is collective): batches formed
  // from each rank's own arrivals would differ in size from rank to rank and the exchange would time out
  if (idx->nranks > 1)
    return rag_set_error(RAG_ERR_UNSUPPORTED, "rag_batcher_create: the index is one shard of %d; batch on ONE front end and broadcast the batch to the ranks", idx->nranks);
  const rag_hybrid_opts& o = d->opts;
  if (o.vector_top_k == 0 || o.vector_top_k > RAG_MAX_TOPK || o.keyword_limit > RAG_MAX_KEYWORDS || o.fresh_limit > RAG_MAX_FRESH)
    return rag_set_error(RAG_ERR_INVALID, "rag_batcher_create: bad hybrid options");
  rag_batcher* b = new (std::nothrow) rag_batcher();
  if (!b) return rag_set_error(RAG_ERR_NOMEM, "out of host memory");
  b->idx = idx;
  b->desc = *d;
  const uint32_t B = d->max_batch, k = o.vector_top_k, cap = k + o.keyword_limit + o.fresh_limit;
  for (int i = 0; i < kBuffers; i++) {
    rag_batch_buf& bb = b->bufs[i];
    bb.h_q = (float*)rag_host_alloc((uint64_t)B * idx->dim * sizeof(float));
    if (!bb.h_q) {
      for (int j = 0; j < i; j++) rag_host_free(b->bufs[j].h_q);
      delete b;
      return RAG_ERR_NOMEM;
    }
    bb.kw_keys.assign((size_t)B * std::max(1u, o.keyword_limit), 0);
    bb.kw_counts.assign(B, 0);
    bb.keys.resize((size_t)B * cap); bb.scores.resize((size_t)B * cap);
    bb.source.resize((size_t)B * cap); bb.ctype.resize((size_t)B * cap);
    bb.counts.resize(B); bb.used_rrf.resize(B); bb.certified.resize(B);
    bb.vec_ids.resize((size_t)B * k); bb.vec_scores.resize((size_t)B * k); bb.vec_counts.resize(B);
    bb.reqs.assign(B, nullptr);
    bb.areqs.assign(B, rag_async_req());
    b->free_bufs.push_back(&bb);
  }
  for (int w = 0; w < kWorkers; w++) b->workers[w] = std::thread(worker_main, b);
  *out = b;
  return RAG_OK;
}

}  // extern "C"

namespace {
// Under b->mu: a slot of the open batch (opening one from a free buffer if need be). `wait` = block while every buffer is
// in flight (back-pressure); otherwise RAG_ERR_BUSY. The slot's owner is recorded before the lock is dropped.
int take_slot(rag_batcher* b, bool wait, rag_batcher_req* r, const rag_async_req* a, rag_batch_buf** out_bb, uint32_t* out_slot) {
  std::unique_lock<std::mutex> lk(b->mu);
  for (;;) {
    if (b->stop) return rag_set_error(RAG_ERR_STATE, "rag_batcher_submit: batcher is shutting down");
    if (b->open && b->open->count == b->desc.max_batch) {  // filled up behind a busy GPU: queue it whole
      b->full.push_back(b->open);
      b->open = nullptr;
      b->cv_work.notify_all();
    }
    if (b->open) break;
    if (!b->free_bufs.empty()) {
      b->open = b->free_bufs.back();
      b->free_bufs.pop_back();
      break;
    }
    if (!wait) return rag_set_error(RAG_ERR_BUSY, "rag_batcher_submit_async: every batch buffer is in flight");
    b->cv_free.wait(lk);  // every buffer is in flight: back-pressure
  }
  rag_batch_buf* bb = b->open;
  const uint32_t slot = bb->count++;
  bb->reqs[slot] = r;
  bb->areqs[slot] = a ? *a : rag_async_req();
  bb->t_last = std::chrono::steady_clock::now();
  if (slot == 0) bb->t_first = bb->t_last;
  if (slot == 0 || bb->count == b->desc.max_batch) b->cv_work.notify_all();
  *out_bb = bb;
  *out_slot = slot;
  return RAG_OK;
}

// this request's inputs into its slot (outside the lock, in parallel with every other caller)
void stage_slot(rag_batcher* b, rag_batch_buf* bb, uint32_t slot, const float* query, const uint64_t* kw_keys, uint32_t kw_count) {
  const uint32_t kl = b->desc.opts.keyword_limit, dim = b->idx->dim;
  memcpy(bb->h_q + (size_t)slot * dim, query, (size_t)dim * sizeof(float));
  const uint32_t c = kw_keys ? kw_count : 0u;
  if (c) memcpy(bb->kw_keys.data() + (size_t)slot * kl, kw_keys, (size_t)c * 8);
  bb->kw_counts[slot] = c;
  bb->staged.fetch_add(1, std::memory_order_release);
}

int check_submit(const rag_batcher* b, const float* query, uint32_t kw_count, const rag_fused_out* out, const char* who) {
  if (!b || !query || !out || !out->keys || !out->scores || !out->counts) return rag_set_error(RAG_ERR_INVALID, "%s: null argument", who);
  const rag_hybrid_opts& o = b->desc.opts;
  if (out->capacity < o.vector_top_k + o.keyword_limit + o.fresh_limit) return rag_set_error(RAG_ERR_INVALID, "%s: rag_fused_out.capacity too small", who);
  if (kw_count > o.keyword_limit) return rag_set_error(RAG_ERR_INVALID, "%s: kw_count exceeds keyword_limit", who);
  return RAG_OK;
}
}  // namespace

extern "C" {

// Blocking; callable from any number of threads. `query` is one [dim] fp32 vector, `kw_keys` the
// request's keyword hits as fusion keys in rank order; `out` is shaped for ONE query (capacity >=
// vector_top_k + keyword_limit + fresh_limit; counts/used_rrf/certified/vec_counts have 1 entry).
int rag_batcher_submit(rag_batcher* b, const float* query, const uint64_t* kw_keys, uint32_t kw_count, rag_fused_out* out) {
  RAG_CHECK(check_submit(b, query, kw_count, out, "rag_batcher_submit"));
  active_call in_flight(b);
  rag_batcher_req r;
  rag_batch_buf* bb = nullptr;
  uint32_t slot = 0;
  RAG_CHECK(take_slot(b, true, &r, nullptr, &bb, &slot));
  stage_slot(b, bb, slot, query, kw_keys, kw_count);
  {
    rag_wake_group& g = b->groups[(slot / kGroupSlots) % kWakeGroups];
    std::unique_lock<std::mutex> lk(g.mu);
    g.cv.wait(lk, [&] { return r.done; });
  }
  const int rc = bb->rc;
  if (rc == RAG_OK) copy_out(b, bb, slot, out);
  else rag_set_error(rc, "%s", bb->err.c_str());
  release_reader(b, bb);
  return rc;
}

// Non-blocking: takes a slot, copies `query` / `kw_keys` into it (the caller may reuse them at once) and returns. `out`
// (shaped for one query, like rag_batcher_submit's) must stay valid until `done(user, rc, err)` is called — from a
// batcher worker thread, after the result has been copied into `out` (rc == RAG_OK) or with the batch's error (`err` is
// valid during the callback only). RAG_ERR_BUSY when every batch buffer is in flight: nothing was queued and `done` will
// not be called — retry later or fall back to the blocking call. A non-zero return never calls `done`.
int rag_batcher_submit_async(rag_batcher* b, const float* query, const uint64_t* kw_keys, uint32_t kw_count, rag_fused_out* out,
                             rag_batcher_done_fn done, void* user) {
  RAG_CHECK(check_submit(b, query, kw_count, out, "rag_batcher_submit_async"));
  if (!done) return rag_set_error(RAG_ERR_INVALID, "rag_batcher_submit_async: null completion callback");
  active_call in_flight(b);
  rag_async_req a;
  a.done = done;
  a.user = user;
  a.out = *out;
  rag_batch_buf* bb = nullptr;
  uint32_t slot = 0;
  RAG_CHECK(take_slot(b, false, nullptr, &a, &bb, &slot));
  stage_slot(b, bb, slot, query, kw_keys, kw_count);
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

// May be called while submitters are still inside rag_batcher_submit: requests that already hold a slot are run and
// answered, submitters still waiting for a buffer fail with RAG_ERR_STATE, and the buffers are freed only after the last
// of them has copied its result out and returned. No submit may START once destroy has been called (the handle dies).
void rag_batcher_destroy(rag_batcher* b) {
  if (!b) return;
  {
    std::lock_guard<std::mutex> lk(b->mu);
    b->stop = true;
  }
  b->cv_work.notify_all();
  b->cv_free.notify_all();
  for (int w = 0; w < kWorkers; w++)
    if (b->workers[w].joinable()) b->workers[w].join();
  while (b->active.load(std::memory_order_acquire) != 0) {  // woken callers copying their results out / leaving
    b->cv_free.notify_all();
    std::this_thread::sleep_for(std::chrono::microseconds(50));
  }
  for (int i = 0; i < kBuffers; i++) rag_host_free(b->bufs[i].h_q);
  delete b;
}

}  // extern "C"
