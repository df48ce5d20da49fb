// batcher_load.cc — closed-loop load on the micro-batcher through the C ABI: T request threads, each submitting
// one query at a time (what the reference's per-request hybridSearch does), against ONE index. Reports requests/s, latency
// percentiles and the batch sizes the batcher formed, next to one thread calling rag_hybrid_search directly (batch 1), and ONE
// thread keeping W requests in flight through rag_batcher_submit_async (the N-API addon's route on Node's event loop).
//   g++ -O2 -std=c++17 -I include tools/micro/batcher_load.cc -o tools/micro/batcher_load -L rag_era_b200 -lragera \
//       -Wl,-rpath,$PWD/rag_era_b200 -lpthread
//   ./tools/micro/batcher_load [rows=1000000] [dim=1536] [seconds=3] [f16 shadow: 1|0]
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "ragera.h"

#define CHECK(x) do { int rc_ = (x); if (rc_ != RAG_OK) { fprintf(stderr, "%s failed (%d): %s\n", #x, rc_, rag_last_error()); exit(2); } } while (0)

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
  const uint64_t rows = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1000000;
  const uint32_t dim = argc > 2 ? (uint32_t)atoi(argv[2]) : 1536;
  const double seconds = argc > 3 ? atof(argv[3]) : 3.0;
  const bool shadow = argc > 4 ? atoi(argv[4]) != 0 : true;
  const uint32_t k = 10, kl = 10, cap = k + kl, NQ = 4096;
  rag_index_desc d = {rows, dim, RAG_F32, 0, shadow ? RAG_INDEX_F16_SHADOW : 0u, 0};
  rag_index* idx = nullptr;
  CHECK(rag_index_create(&d, &idx));
  rag_gen_desc g;
  memset(&g, 0, sizeof g);
  g.seed = 0xC0FFEE; g.query_seed = 0xBEEF; g.meta_seed = 0xF00D; g.total_rows = rows;
  g.n_clusters = 4096; g.noise = 0.6f; g.query_noise = 0.5f;
  CHECK(rag_index_generate(idx, &g, rows));
  std::vector<float> Q((size_t)NQ * dim);
  CHECK(rag_generate_queries(idx, &g, 0, NQ, Q.data()));
  rag_hybrid_opts o;
  memset(&o, 0, sizeof o);
  o.vector_top_k = k; o.keyword_limit = kl; o.min_vector_score = 0.3;
  o.rrf.k = 60; o.rrf.vector_weight = 1; o.rrf.keyword_weight = 1; o.rrf.both_bonus = 0.1;
  std::vector<uint64_t> kw((size_t)NQ * kl);
  for (size_t i = 0; i < kw.size(); i++) kw[i] = (i * 2654435761ull) % rows;

  printf("batcher load: %llu x %u fp32%s, deep_search shape (vectorTopK 10, keywordLimit 10), %.1f s per line\n",
         (unsigned long long)rows, dim, shadow ? " + fp16 shadow" : "", seconds);
  // baseline: one thread, direct batch-1 calls
  {
    uint64_t keys[cap]; double scores[cap]; uint8_t src[cap], ct[cap], rrf[1], cert[1]; uint32_t counts[1];
    rag_fused_out out = {cap, keys, scores, src, ct, counts, rrf, nullptr, nullptr, nullptr, cert};
    const uint32_t one = kl;
    std::vector<double> lat;
    const double t0 = now_s();
    uint32_t i = 0;
    while (now_s() - t0 < seconds) {
      const double a = now_s();
      CHECK(rag_hybrid_search(idx, Q.data() + (size_t)(i % NQ) * dim, 1, &o, kw.data() + (size_t)(i % NQ) * kl, &one, &out));
      lat.push_back(now_s() - a);
      i++;
    }
    const double dt = now_s() - t0;
    std::sort(lat.begin(), lat.end());
    printf("direct, 1 thread, batch 1        : %9.0f req/s   p50 %8.3f ms  p99 %8.3f ms\n", lat.size() / dt, lat[lat.size() / 2] * 1e3, lat[lat.size() * 99 / 100] * 1e3);
  }
  for (uint32_t wait_us : {200u, 1000u}) {
    for (int T : {8, 64, 256, 1024}) {
      rag_batcher_desc bd;
      memset(&bd, 0, sizeof bd);
      bd.max_batch = 1024; bd.max_wait_us = wait_us; bd.opts = o;
      rag_batcher* bt = nullptr;
      CHECK(rag_batcher_create(idx, &bd, &bt));
      std::atomic<bool> stop{false};
      std::atomic<uint64_t> bad{0};
      std::vector<std::vector<double>> lats(T);
      std::vector<std::thread> th;
      const double t0 = now_s();
      for (int t = 0; t < T; t++)
        th.emplace_back([&, t]() {
          uint64_t keys[cap]; double scores[cap]; uint8_t src[cap], ct[cap], rrf[1], cert[1]; uint32_t counts[1];
          rag_fused_out out = {cap, keys, scores, src, ct, counts, rrf, nullptr, nullptr, nullptr, cert};
          uint32_t i = (uint32_t)t * 7919u;
          while (!stop.load(std::memory_order_relaxed)) {
            const double a = now_s();
            const int rc = rag_batcher_submit(bt, Q.data() + (size_t)(i % NQ) * dim, kw.data() + (size_t)(i % NQ) * kl, kl, &out);
            if (rc != RAG_OK || !cert[0] || counts[0] == 0) bad++;
            lats[t].push_back(now_s() - a);
            i++;
          }
        });
      std::this_thread::sleep_for(std::chrono::duration<double>(seconds));
      stop = true;
      for (auto& x : th) x.join();
      const double dt = now_s() - t0;
      std::vector<double> lat;
      for (auto& v : lats) lat.insert(lat.end(), v.begin(), v.end());
      std::sort(lat.begin(), lat.end());
      uint64_t nb = 0, nq = 0, big = 0;
      CHECK(rag_batcher_stats(bt, &nb, &nq, &big));
      printf("batcher wait %4u us, %4d threads  : %9.0f req/s   p50 %8.3f ms  p99 %8.3f ms   %llu batches, mean %.1f, largest %llu, failed %llu\n", wait_us, T,
             lat.size() / dt, lat[lat.size() / 2] * 1e3, lat[lat.size() * 99 / 100] * 1e3, (unsigned long long)nb, nb ? (double)nq / nb : 0.0,
             (unsigned long long)big, (unsigned long long)bad.load());
      rag_batcher_destroy(bt);
    }
  }
  // ONE thread, W requests in flight through rag_batcher_submit_async (what the N-API addon does on Node's event loop): a
  // completion hands the request's index back through a queue, the thread re-submits it at once (closed loop)
  for (int W : {64, 256, 1024, 4096}) {
    rag_batcher_desc bd;
    memset(&bd, 0, sizeof bd);
    bd.max_batch = 1024; bd.max_wait_us = 1000; bd.opts = o;
    rag_batcher* bt = nullptr;
    CHECK(rag_batcher_create(idx, &bd, &bt));
    struct req { uint64_t keys[cap]; double scores[cap]; uint8_t src[cap], ct[cap], rrf[1], cert[1]; uint32_t counts[1]; rag_fused_out out; double t_submit; int id; void* owner; };
    struct shared { std::mutex mu; std::condition_variable cv; std::vector<int> ready; uint64_t bad = 0; } sh;
    std::vector<req> reqs(W);
    for (int i = 0; i < W; i++) {
      req& r = reqs[i];
      r.out = {cap, r.keys, r.scores, r.src, r.ct, r.counts, r.rrf, nullptr, nullptr, nullptr, r.cert};
      r.id = i; r.owner = &sh;
      sh.ready.push_back(i);
    }
    auto on_done = [](void* user, int rc, const char*) {
      req* r = static_cast<req*>(user);
      shared* s = static_cast<shared*>(r->owner);
      std::lock_guard<std::mutex> lk(s->mu);
      if (rc != RAG_OK || !r->cert[0] || r->counts[0] == 0) s->bad++;
      s->ready.push_back(r->id);
      s->cv.notify_one();
    };
    std::vector<double> lat;
    uint64_t busy = 0, issued = 0;
    int outstanding = 0;
    const double t0 = now_s();
    std::vector<int> mine;
    bool first_round = true;
    while (now_s() - t0 < seconds || outstanding > 0) {
      {
        std::unique_lock<std::mutex> lk(sh.mu);
        if (sh.ready.empty()) sh.cv.wait_for(lk, std::chrono::milliseconds(1));
        mine.swap(sh.ready);
      }
      const double now = now_s();
      for (int id : mine) {
        req& r = reqs[id];
        if (!first_round || r.t_submit != 0.0) { lat.push_back(now - r.t_submit); outstanding--; }
        if (now - t0 >= seconds) continue;   // draining
        const uint32_t q = (uint32_t)((issued * 7919u) % NQ);
        for (;;) {
          r.t_submit = now_s();
          const int rc = rag_batcher_submit_async(bt, Q.data() + (size_t)q * dim, kw.data() + (size_t)q * kl, kl, &r.out, on_done, &r);
          if (rc == RAG_OK) break;
          if (rc != RAG_ERR_BUSY) { fprintf(stderr, "submit_async failed (%d): %s\n", rc, rag_last_error()); exit(2); }
          busy++;
          std::this_thread::sleep_for(std::chrono::microseconds(20));
        }
        issued++;
        outstanding++;
      }
      mine.clear();
      first_round = false;
    }
    const double dt = now_s() - t0;
    std::sort(lat.begin(), lat.end());
    uint64_t nb = 0, nq = 0, big = 0;
    CHECK(rag_batcher_stats(bt, &nb, &nq, &big));
    printf("async, 1 thread, %4d in flight   : %9.0f req/s   p50 %8.3f ms  p99 %8.3f ms   %llu batches, mean %.1f, largest %llu, busy %llu, failed %llu\n", W,
           lat.size() / dt, lat.empty() ? 0.0 : lat[lat.size() / 2] * 1e3, lat.empty() ? 0.0 : lat[lat.size() * 99 / 100] * 1e3, (unsigned long long)nb,
           nb ? (double)nq / nb : 0.0, (unsigned long long)big, (unsigned long long)busy, (unsigned long long)sh.bad);
    rag_batcher_destroy(bt);
  }
  rag_index_destroy(idx);
  return 0;
}
