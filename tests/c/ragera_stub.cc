// ragera_stub.cc — TEST HARNESS: a stand-in for libragera.so's device entry points, so that the N-API addon
// (integration/node/ragera_addon.cc), the mock Node-API host (napi_mock.cc) and the REAL micro-batcher (batcher.cu, host
// code) can run together on a machine without a GPU, under ThreadSanitizer (tests/test_napi_tsan.py). What is under test is
// the addon's threading and handle lifetime — submits answered through the thread-safe function, the RAG_ERR_BUSY
// fallback, destroy()/destroyBatcher() with calls still in flight — not numerics: every "result" here is a fixed function
// of the query's first element and its keyword keys. Every entry point checks that the handle it is given is still alive
// and aborts on a use after destroy.
#include "common.cuh"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <mutex>
#include <set>
#include <thread>

static thread_local char g_err[1024];
int rag_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

namespace {
std::mutex g_mu;
std::set<const rag_index*> g_live;
struct stub_rows { uint64_t rows = 0; };
void must_be_alive(const rag_index* idx, const char* who) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_live.count(idx)) {
    fprintf(stderr, "STUB: %s on an index that was destroyed (or never created)\n", who);
    abort();
  }
}
uint64_t tag_of(const float* q) { return (uint64_t)q[0]; }
}  // namespace

extern "C" {
const char* rag_last_error(void) { return g_err; }
void* rag_host_alloc(uint64_t bytes) { return malloc(bytes); }
void rag_host_free(void* p) { free(p); }
void rag_free(void* p) { free(p); }

int rag_index_create(const rag_index_desc* d, rag_index** out) {
  rag_index* idx = new rag_index();
  idx->dim = d->dim;
  idx->rows = 0;
  std::lock_guard<std::mutex> lk(g_mu);
  g_live.insert(idx);
  *out = idx;
  return RAG_OK;
}
void rag_index_destroy(rag_index* idx) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_live.erase(idx)) {
      fprintf(stderr, "STUB: double destroy\n");
      abort();
    }
  }
  delete idx;
}
int rag_index_upload(rag_index* idx, uint64_t row0, uint64_t nrows, const void*) {
  must_be_alive(idx, "rag_index_upload");
  std::lock_guard<std::mutex> lk(g_mu);
  idx->rows = row0 + nrows;
  return RAG_OK;
}
uint64_t rag_index_rows(const rag_index* idx) {
  must_be_alive(idx, "rag_index_rows");
  std::lock_guard<std::mutex> lk(g_mu);
  return idx->rows;
}
int rag_index_set_row_meta(rag_index* idx, uint64_t, uint64_t, const uint8_t*, const double*, const int32_t*, const int64_t*) {
  must_be_alive(idx, "rag_index_set_row_meta");
  return RAG_OK;
}
int rag_index_set_row_keys(rag_index* idx, uint64_t, uint64_t, const uint64_t*) {
  must_be_alive(idx, "rag_index_set_row_keys");
  return RAG_OK;
}
int rag_index_open_store(rag_index* idx, const char*, const char*, uint64_t*, char**, uint64_t*, int*) {
  must_be_alive(idx, "rag_index_open_store");
  return rag_set_error(RAG_ERR_UNSUPPORTED, "stub");
}

int rag_hybrid_search(rag_index* idx, const float* queries, uint32_t B, const rag_hybrid_opts* o, const uint64_t* kw_keys,
                      const uint32_t* kw_counts, rag_fused_out* out) {
  must_be_alive(idx, "rag_hybrid_search");
  if (o->vector_top_k > RAG_MAX_TOPK) return rag_set_error(RAG_ERR_INVALID, "vector_top_k exceeds %d", RAG_MAX_TOPK);
  std::this_thread::sleep_for(std::chrono::microseconds(200));  // a corpus pass: requests pile up behind it
  must_be_alive(idx, "rag_hybrid_search (after the pass)");
  const uint32_t k = o->vector_top_k, cap = out->capacity;
  for (uint32_t b = 0; b < B; b++) {
    const uint64_t tag = tag_of(queries + (size_t)b * idx->dim);
    const uint32_t n = 1 + (uint32_t)(tag % k);
    for (uint32_t i = 0; i < n; i++) {
      out->keys[(size_t)b * cap + i] = tag * 1000 + i;
      out->scores[(size_t)b * cap + i] = (double)tag + 0.001 * i;
      if (out->source) out->source[(size_t)b * cap + i] = (uint8_t)(i % 3);
      if (out->content_type) out->content_type[(size_t)b * cap + i] = (uint8_t)(tag % 3);
    }
    uint64_t kwsum = 0;
    for (uint32_t i = 0; i < kw_counts[b]; i++) kwsum += kw_keys[(size_t)b * o->keyword_limit + i];
    out->keys[(size_t)b * cap + n] = kwsum;  // one extra entry carries the keyword list's checksum
    out->scores[(size_t)b * cap + n] = 0.0;
    if (out->source) out->source[(size_t)b * cap + n] = 1;
    if (out->content_type) out->content_type[(size_t)b * cap + n] = 0;
    out->counts[b] = n + 1;
    if (out->used_rrf) out->used_rrf[b] = kw_counts[b] ? 1 : 0;
    if (out->certified) out->certified[b] = 1;
    if (out->vec_ids && out->vec_counts) {
      out->vec_counts[b] = 1;
      out->vec_ids[(size_t)b * k] = tag;
      out->vec_scores[(size_t)b * k] = (double)tag;
    }
  }
  return RAG_OK;
}

int rag_search(rag_index* idx, const float* queries, uint32_t B, const rag_search_opts* o, rag_topk_out* out) {
  must_be_alive(idx, "rag_search");
  std::this_thread::sleep_for(std::chrono::microseconds(200));
  must_be_alive(idx, "rag_search (after the pass)");
  for (uint32_t b = 0; b < B; b++) {
    const uint64_t tag = tag_of(queries + (size_t)b * idx->dim);
    for (uint32_t i = 0; i < o->k; i++) {
      out->ids[(size_t)b * o->k + i] = tag * 10 + i;
      out->scores[(size_t)b * o->k + i] = (double)tag - i;
    }
    out->counts[b] = o->k;
    if (out->certified) out->certified[b] = 1;
  }
  return RAG_OK;
}

int rag_memory_retrieve(rag_index* idx, const float* queries, uint32_t B, const rag_memory_opts* o, rag_memory_out* out) {
  must_be_alive(idx, "rag_memory_retrieve");
  for (uint32_t b = 0; b < B; b++) {
    const uint64_t tag = tag_of(queries + (size_t)b * idx->dim);
    out->ids[(size_t)b * o->limit] = tag;
    out->scores[(size_t)b * o->limit] = out->relevance[(size_t)b * o->limit] = out->freshness[(size_t)b * o->limit] = (double)tag;
    out->counts[b] = 1;
  }
  return RAG_OK;
}
}
