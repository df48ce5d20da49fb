// batcher_tsan.cc — TEST HARNESS: the micro-batcher (batcher.cu, host-only) compiled as plain C++ with -fsanitize=thread and
// driven by many concurrent submitters against a STUB rag_hybrid_search that derives every output from the query it was
// given. Checks: every submitter gets exactly its own result back (no cross-talk between slots of a batch), errors of a
// batch reach every waiter, batches really form (largest batch > 1), callers that outnumber the batch buffers wait and are
// all served (back-pressure), a lone caller is not left waiting, destroy with callers in flight answers them and frees nothing
// under them, one thread keeps hundreds of requests in flight through the asynchronous submit, and ThreadSanitizer reports no race. Run by tests/test_gpu_batcher.py::test_batcher_threads_under_tsan (CPU).
#include "common.cuh"
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>
static thread_local char g_err[1024];
int rag_set_error(int code, const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); return code; }
static std::atomic<int> g_calls{0}, g_fail_every{0}, g_gate_closed{0};
extern "C" {
const char* rag_last_error(void) { return g_err; }
void* rag_host_alloc(uint64_t bytes) { return malloc(bytes); }
void rag_host_free(void* p) { free(p); }
// the stub: result of query b = f(query[b][0], its keyword keys); sleeps like a corpus pass so that requests pile up
int rag_hybrid_search(rag_index* idx, const float* queries, uint32_t B, const rag_hybrid_opts* o, const uint64_t* kw_keys,
                      const uint32_t* kw_counts, rag_fused_out* out) {
  std::this_thread::sleep_for(std::chrono::microseconds(300));
  while (g_gate_closed.load()) std::this_thread::sleep_for(std::chrono::microseconds(100));  // phase 5 holds a pass on the "GPU"
  const int call = ++g_calls;
  if (g_fail_every && call % g_fail_every == 0) return rag_set_error(RAG_ERR_CUDA, "injected failure of batch %d", call);
  const uint32_t k = o->vector_top_k, cap = out->capacity;
  for (uint32_t b = 0; b < B; b++) {
    const uint64_t tag = (uint64_t)queries[(size_t)b * idx->dim];
    const uint32_t n = 1 + (uint32_t)(tag % k);
    for (uint32_t i = 0; i < n; i++) { out->keys[(size_t)b * cap + i] = tag * 1000 + i; out->scores[(size_t)b * cap + i] = (double)tag + 0.001 * i; }
    uint64_t kwsum = 0;
    for (uint32_t i = 0; i < kw_counts[b]; i++) kwsum += kw_keys[(size_t)b * o->keyword_limit + i];
    out->keys[(size_t)b * cap + n] = kwsum;  // one extra slot carries the keyword list's checksum
    out->counts[b] = n + 1;
    if (out->used_rrf) out->used_rrf[b] = kw_counts[b] ? 1 : 0;
    if (out->certified) out->certified[b] = 1;
    if (out->vec_ids && out->vec_counts) { out->vec_counts[b] = 1; out->vec_ids[(size_t)b * k] = tag; out->vec_scores[(size_t)b * k] = (double)tag; }
  }
  return RAG_OK;
}
}
int main() {
  rag_index idx;  // never touched by the batcher except for dim and as an opaque handle for the stub
  idx.dim = 16;
  rag_batcher_desc d;
  memset(&d, 0, sizeof d);
  d.max_batch = 8;
  d.max_wait_us = 200;
  d.opts.vector_top_k = 5;
  d.opts.keyword_limit = 4;
  rag_batcher* bt = nullptr;
  if (rag_batcher_create(&idx, &d, &bt) != RAG_OK) { printf("create failed: %s\n", g_err); return 1; }
  std::atomic<int> wrong{0}, failed{0}, done{0};
  auto submitter = [&](int tid, int n) {
    for (int r = 0; r < n; r++) {
      const uint64_t tag = (uint64_t)tid * 100000 + r + 1;
      float q[16] = {(float)tag};
      uint64_t kw[4] = {tag, tag + 1, 7, 9};
      const uint32_t kwc = (uint32_t)(tag % 5);  // 0..4
      uint64_t keys[9]; double scores[9]; uint8_t src[9], ct[9], rrf, cert; uint32_t cnt, vcnt; uint64_t vid[5]; double vs[5];
      rag_fused_out out = {9, keys, scores, src, ct, &cnt, &rrf, vid, vs, &vcnt, &cert};
      const int rc = rag_batcher_submit(bt, q, kw, kwc, &out);
      if (rc != RAG_OK) { failed++; continue; }
      const uint32_t n_exp = 1 + (uint32_t)(tag % 5);
      uint64_t kwsum = 0;
      for (uint32_t i = 0; i < kwc; i++) kwsum += kw[i];
      bool ok = cnt == n_exp + 1 && keys[n_exp] == kwsum && vid[0] == tag && rrf == (kwc ? 1 : 0);
      for (uint32_t i = 0; ok && i < n_exp; i++) ok = keys[i] == tag * 1000 + i && scores[i] == (double)tag + 0.001 * i;
      if (!ok) wrong++;
      done++;
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < 24; t++) th.emplace_back(submitter, t, 40);
  for (auto& t : th) t.join();
  uint64_t batches = 0, queries = 0, largest = 0;
  rag_batcher_stats(bt, &batches, &queries, &largest);
  printf("phase 1: done=%d wrong=%d failed=%d batches=%llu queries=%llu largest=%llu\n", done.load(), wrong.load(), failed.load(),
         (unsigned long long)batches, (unsigned long long)queries, (unsigned long long)largest);
  const bool p1 = done == 24 * 40 && wrong == 0 && failed == 0 && queries == 24 * 40 && largest > 1 && largest <= 8;
  // phase 2: every 3rd batch fails: all its waiters must see the error, the others their own results
  g_fail_every = 3;
  done = 0; wrong = 0; failed = 0;
  th.clear();
  for (int t = 0; t < 12; t++) th.emplace_back(submitter, 100 + t, 20);
  for (auto& t : th) t.join();
  printf("phase 2: done=%d wrong=%d failed=%d\n", done.load(), wrong.load(), failed.load());
  const bool p2 = done + failed == 12 * 20 && wrong == 0 && failed > 0 && done > 0;
  // phase 3: bad arguments are refused without touching the queue
  float q[16] = {1.f};
  uint64_t keys[2]; double scores[2]; uint32_t cnt;
  rag_fused_out small = {2, keys, scores, nullptr, nullptr, &cnt, nullptr, nullptr, nullptr, nullptr, nullptr};
  const bool p3 = rag_batcher_submit(bt, q, nullptr, 0, &small) == RAG_ERR_INVALID && rag_batcher_submit(bt, nullptr, nullptr, 0, &small) == RAG_ERR_INVALID;
  rag_batcher_destroy(bt);
  // phase 4: back-pressure — 96 callers against 4 buffers of 8 slots: submitters wait for a free buffer, full batches queue up
  // behind the running one, buffers are recycled by the last caller that reads its result
  g_fail_every = 0;
  done = 0; wrong = 0; failed = 0;
  d.max_wait_us = 50;
  if (rag_batcher_create(&idx, &d, &bt) != RAG_OK) { printf("create failed: %s\n", g_err); return 1; }
  th.clear();
  for (int t = 0; t < 96; t++) th.emplace_back(submitter, t, 20);   // tags stay below 2^24: they travel as one float
  for (auto& t : th) t.join();
  rag_batcher_stats(bt, &batches, &queries, &largest);
  printf("phase 4: done=%d wrong=%d failed=%d batches=%llu queries=%llu largest=%llu\n", done.load(), wrong.load(), failed.load(),
         (unsigned long long)batches, (unsigned long long)queries, (unsigned long long)largest);
  const bool p4 = done == 96 * 20 && wrong == 0 && failed == 0 && queries == 96 * 20 && largest == 8 && batches < 96 * 20 / 4;
  // a lone caller is answered after the collection window, not left waiting for company
  {
    const auto t0 = std::chrono::steady_clock::now();
    submitter(150, 3);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    printf("lone caller: 3 requests in %.2f ms\n", ms);
    if (ms > 500.0 || wrong != 0 || failed != 0) { printf("result: FAILED\n"); return 1; }
  }
  rag_batcher_destroy(bt);
  // phase 5: destroy while callers are inside submit. 40 callers against 32 slots with the stub held shut: one batch sits on
  // the "GPU", the rest hold slots or wait for a buffer. destroy must answer every request that holds a slot, fail the ones
  // still waiting for a buffer with RAG_ERR_STATE, and free nothing before the last caller has left.
  done = 0; wrong = 0; failed = 0;
  if (rag_batcher_create(&idx, &d, &bt) != RAG_OK) { printf("create failed: %s\n", g_err); return 1; }
  g_gate_closed = 1;
  std::atomic<int> entering{0};
  th.clear();
  for (int t = 0; t < 40; t++) th.emplace_back([&, t] { entering++; submitter(100 + t, 1); });   // tags below 2^24
  while (entering.load() != 40) std::this_thread::yield();
  std::this_thread::sleep_for(std::chrono::milliseconds(600));   // every caller is inside submit by now (the stub is shut; generous for a loaded box)
  std::thread closer([&] { rag_batcher_destroy(bt); });
  std::this_thread::sleep_for(std::chrono::milliseconds(20));
  g_gate_closed = 0;
  closer.join();
  for (auto& t : th) t.join();
  printf("phase 5: done=%d wrong=%d failed=%d\n", done.load(), wrong.load(), failed.load());
  const bool p5 = done + failed == 40 && wrong == 0 && done >= 8 && failed > 0;   // failed = the callers still waiting for a buffer
  // phase 6: ONE thread keeps hundreds of requests in flight through rag_batcher_submit_async (what Node's event loop does),
  // four blocking callers share the batches with it; every 5th batch fails. Results land in the caller's arrays before
  // `done` runs (on a worker thread); RAG_ERR_BUSY (all 32 slots in flight) queues nothing and calls nothing.
  done = 0; wrong = 0; failed = 0;
  g_fail_every = 5;
  d.max_wait_us = 200;
  if (rag_batcher_create(&idx, &d, &bt) != RAG_OK) { printf("create failed: %s\n", g_err); return 1; }
  struct async_req {
    uint64_t tag; uint32_t kwc; uint64_t kw[4];
    uint64_t keys[9]; double scores[9]; uint8_t src[9], ct[9], rrf, cert; uint32_t cnt, vcnt; uint64_t vid[5]; double vs[5];
    rag_fused_out out;
    std::atomic<int> state{0};   // 1 = done OK, 2 = done with an error
    std::atomic<int>* completions;
  };
  const int n_async = 400;
  std::vector<async_req> reqs(n_async);
  std::atomic<int> completions{0};
  int busy = 0, refused_wrong = 0;
  th.clear();
  for (int t = 0; t < 4; t++) th.emplace_back(submitter, 120 + t, 20);
  for (int i = 0; i < n_async; i++) {
    async_req& a = reqs[i];
    a.tag = 13000000ull + i;   // below 2^24
    a.kwc = (uint32_t)(a.tag % 5);
    a.kw[0] = a.tag; a.kw[1] = a.tag + 1; a.kw[2] = 7; a.kw[3] = 9;
    a.out = {9, a.keys, a.scores, a.src, a.ct, &a.cnt, &a.rrf, a.vid, a.vs, &a.vcnt, &a.cert};
    a.completions = &completions;
    float q[16] = {(float)a.tag};
    uint64_t kw_copy[4] = {a.kw[0], a.kw[1], a.kw[2], a.kw[3]};
    for (;;) {
      const int rc = rag_batcher_submit_async(bt, q, kw_copy, a.kwc, &a.out, [](void* user, int rc2, const char* err) {
        async_req* r = static_cast<async_req*>(user);
        if (rc2 != RAG_OK && (!err || !strstr(err, "injected failure"))) r->state = 3;   // the batch's own error text must arrive
        else r->state = rc2 == RAG_OK ? 1 : 2;
        (*r->completions)++;
      }, &a);
      if (rc == RAG_OK) break;
      if (rc != RAG_ERR_BUSY) { refused_wrong++; break; }
      busy++;
      std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
    q[0] = -1.f; kw_copy[0] = kw_copy[1] = 0;   // the inputs were copied: scribbling over them must not matter
  }
  while (completions.load() != n_async - refused_wrong) std::this_thread::sleep_for(std::chrono::microseconds(100));
  for (auto& t : th) t.join();
  int a_ok = 0, a_failed = 0, a_wrong = 0;
  for (async_req& a : reqs) {
    if (a.state == 2) { a_failed++; continue; }
    if (a.state != 1) { a_wrong++; continue; }
    const uint32_t n_exp = 1 + (uint32_t)(a.tag % 5);
    uint64_t kwsum = 0;
    for (uint32_t i = 0; i < a.kwc; i++) kwsum += a.kw[i];
    bool ok = a.cnt == n_exp + 1 && a.keys[n_exp] == kwsum && a.vid[0] == a.tag && a.rrf == (a.kwc ? 1 : 0);
    for (uint32_t i = 0; ok && i < n_exp; i++) ok = a.keys[i] == a.tag * 1000 + i && a.scores[i] == (double)a.tag + 0.001 * i;
    if (ok) a_ok++; else a_wrong++;
  }
  rag_batcher_stats(bt, &batches, &queries, &largest);
  printf("phase 6: async ok=%d failed=%d wrong=%d busy=%d | blocking done=%d wrong=%d failed=%d | batches=%llu largest=%llu\n", a_ok, a_failed, a_wrong, busy,
         done.load(), wrong.load(), failed.load(), (unsigned long long)batches, (unsigned long long)largest);
  const bool p6 = a_wrong == 0 && refused_wrong == 0 && a_ok + a_failed == n_async && a_ok > 0 && a_failed > 0 && busy > 0 && wrong == 0 &&
                  done + failed == 80 && largest == 8 && queries == (uint64_t)n_async + 80;
  // null callback / bad shapes are refused
  {
    float q1[16] = {1.f};
    const bool p6b = rag_batcher_submit_async(bt, q1, nullptr, 0, &reqs[0].out, nullptr, nullptr) == RAG_ERR_INVALID &&
                     rag_batcher_submit_async(bt, q1, nullptr, 9, &reqs[0].out, [](void*, int, const char*) {}, nullptr) == RAG_ERR_INVALID;
    if (!p6b) { printf("phase 6: bad arguments were accepted\nresult: FAILED\n"); return 1; }
  }
  rag_batcher_destroy(bt);
  printf("result: %s\n", (p1 && p2 && p3 && p4 && p5 && p6) ? "OK" : "FAILED");
  return (p1 && p2 && p3 && p4 && p5 && p6) ? 0 : 1;
}
