// ragera_addon.cc — N-API binding of include/ragera.h for the reference's Node process.
//
// NOT compiled in the build container (no Node toolchain there); the same C ABI is exercised end to end
// from Python (rag_era_b200/_native.py) and from C (tests/c/abi_demo.c). Build in the reference repo with
//     cd integration/node && npx node-gyp rebuild
// Pure C N-API (node_api.h), no C++ wrapper dependency. Every search runs on a libuv worker thread
// (napi_create_async_work) so the event loop never blocks. Several requests may be in flight on ONE handle — parallel
// HTTP requests, the Promise.all of engine.ts:108, the synchronous mutators below on the main thread: libragera
// serialises the calls of a handle itself (a per-handle mutex, include/ragera.h), so they take turns on the GPU
// instead of racing; tests/c/napi_mock.cc queues 2B calls on one handle before draining the loop. For throughput
// under load route the requests through the batcher, which turns concurrent batch-1 calls into one corpus pass:
// `submit` does NOT park a pool thread on the request (libuv's pool has 4 threads by default — blocking submits could never
// form a batch above 4): it hands the request to rag_batcher_submit_async on the main thread and the batcher's worker
// resolves the Promise through a thread-safe function, so the event loop keeps thousands of requests in flight.
// Handle lifetime: every queued call holds a reference on its handle; destroy()/destroyBatcher() with calls still in
// flight marks the handle closed (new calls throw) and the LAST completion destroys it — nothing is freed under a worker.
//
// JS surface (see native-retrieval.ts):
//   createIndex({rows, dim, dtype:'f32'|'bf16', device, f16Shadow | bf16Shadow}) -> handle (External)
//   uploadRows(handle, Float32Array rows, nrows, row0 = append) -> first row
//   loadVectorStore(handle, path) -> string[] node ids          (llamaindex vector_store.json, via its binary sidecar)
//   setRowMeta(handle, row0, Uint8Array contentType, Float64Array confidence, Int32Array accessCount, BigInt64Array lastAccessMs)
//   setRowKeys(handle, row0, BigUint64Array keys)
//   hybridSearch(handle, Float32Array queries, B, opts, BigUint64Array kwKeys, Uint32Array kwCounts) -> Promise<result>
//   search(handle, Float32Array queries, B, k) -> Promise<{ids, scores, counts, certified}>      (retriever.retrieve / VectorStore.query)
//   memoryRetrieve(handle, Float32Array queries, B, {limit, minRelevance, nowMs, similarityTopK}) -> Promise<{ids, scores, relevance, freshness, counts}>
//   createBatcher(handle, opts, maxBatch, maxWaitUs) -> batcher; submit(batcher, opts, Float32Array q, BigUint64Array kwKeys) -> Promise<result>
//   destroy(handle) / destroyBatcher(batcher)        (deferred until the calls in flight on the handle have completed)
#include <node_api.h>

#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "ragera.h"

#define NAPI_OK(env, call)                                                   \
  do {                                                                       \
    if ((call) != napi_ok) {                                                 \
      napi_throw_error((env), nullptr, "N-API call failed: " #call);         \
      return nullptr;                                                        \
    }                                                                        \
  } while (0)

namespace {

napi_value throw_rag(napi_env env, int rc) {
  std::string msg = "libragera error " + std::to_string(rc) + ": " + rag_last_error();
  napi_throw_error(env, nullptr, msg.c_str());
  return nullptr;
}

// ---- handle lifetime (main thread only) ------------------------------------------------------------------------------
// What JS holds is an External around one of these. `pending` = queued calls (and live batchers) that still use the
// native object; destroy() only marks it closing, the last release really destroys it; the External's finalizer frees
// the wrapper (after the last release if calls are still in flight when the JS object is collected).
struct IndexHandle {
  rag_index* idx = nullptr;
  uint32_t pending = 0;
  bool closing = false, finalized = false;
};
void index_settle(IndexHandle* h) {
  if (h->pending || !h->closing) return;
  if (h->idx) { rag_index_destroy(h->idx); h->idx = nullptr; }
  if (h->finalized) delete h;
}
void index_finalize(napi_env, void* data, void*) {
  auto* h = static_cast<IndexHandle*>(data);
  h->closing = h->finalized = true;
  index_settle(h);
}
// the live handle behind argv[0], or nullptr with a pending exception
IndexHandle* index_arg(napi_env env, napi_value v) {
  IndexHandle* h = nullptr;
  if (napi_get_value_external(env, v, (void**)&h) != napi_ok || !h) { napi_throw_type_error(env, nullptr, "expected an index handle"); return nullptr; }
  if (h->closing || !h->idx) { napi_throw_error(env, nullptr, "the index handle was destroyed"); return nullptr; }
  return h;
}

struct BatcherHandle {
  rag_batcher* b = nullptr;
  IndexHandle* owner = nullptr;              // holds one `pending` on it for the batcher's lifetime
  napi_threadsafe_function tsfn = nullptr;   // worker thread -> main thread: resolves the Promises of asynchronous submits
  uint32_t pending = 0;
  bool closing = false, finalized = false;
};
void batcher_settle(napi_env env, BatcherHandle* h) {
  if (h->pending || !h->closing) return;
  if (h->b) {
    rag_batcher_destroy(h->b);   // main thread, nothing in flight: returns at once
    h->b = nullptr;
    if (h->tsfn) napi_release_threadsafe_function(h->tsfn, napi_tsfn_release);
    h->tsfn = nullptr;
    h->owner->pending--;
    index_settle(h->owner);
  }
  (void)env;
  if (h->finalized) delete h;
}
void batcher_finalize(napi_env env, void* data, void*) {
  auto* h = static_cast<BatcherHandle*>(data);
  h->closing = h->finalized = true;
  batcher_settle(env, h);
}
BatcherHandle* batcher_arg(napi_env env, napi_value v) {
  BatcherHandle* h = nullptr;
  if (napi_get_value_external(env, v, (void**)&h) != napi_ok || !h) { napi_throw_type_error(env, nullptr, "expected a batcher handle"); return nullptr; }
  if (h->closing || !h->b) { napi_throw_error(env, nullptr, "the batcher was destroyed"); return nullptr; }
  return h;
}
// a request enters / leaves the batcher: the thread-safe function keeps the event loop alive only while requests are out
void batcher_enter(napi_env env, BatcherHandle* h) {
  if (h->pending++ == 0 && h->tsfn) napi_ref_threadsafe_function(env, h->tsfn);
}
void batcher_leave(napi_env env, BatcherHandle* h) {
  if (--h->pending == 0 && h->tsfn) napi_unref_threadsafe_function(env, h->tsfn);
  batcher_settle(env, h);
}

bool get_u32(napi_env env, napi_value obj, const char* name, uint32_t* out) {
  napi_value v;
  bool has = false;
  if (napi_has_named_property(env, obj, name, &has) != napi_ok || !has) return false;
  return napi_get_named_property(env, obj, name, &v) == napi_ok && napi_get_value_uint32(env, v, out) == napi_ok;
}
bool get_f64(napi_env env, napi_value obj, const char* name, double* out) {
  napi_value v;
  bool has = false;
  if (napi_has_named_property(env, obj, name, &has) != napi_ok || !has) return false;
  return napi_get_named_property(env, obj, name, &v) == napi_ok && napi_get_value_double(env, v, out) == napi_ok;
}

// HybridSearchOptions + RRFConfig (src/lib/hybrid-search.ts:40-58) -> rag_hybrid_opts
bool read_opts(napi_env env, napi_value o, rag_hybrid_opts* out) {
  memset(out, 0, sizeof *out);
  out->rrf = rag_rrf_config{60.0, 1.0, 1.0, 0.1};  // PRESET_CONFIGS.document.rrf (:83-88)
  out->vector_top_k = 8;
  out->keyword_limit = 8;
  out->min_vector_score = 0.3;
  get_u32(env, o, "vectorTopK", &out->vector_top_k);
  get_u32(env, o, "keywordLimit", &out->keyword_limit);
  get_f64(env, o, "minVectorScore", &out->min_vector_score);
  napi_value rrf;
  bool has = false;
  if (napi_has_named_property(env, o, "rrf", &has) == napi_ok && has && napi_get_named_property(env, o, "rrf", &rrf) == napi_ok) {
    get_f64(env, rrf, "k", &out->rrf.k);
    get_f64(env, rrf, "vectorWeight", &out->rrf.vector_weight);
    get_f64(env, rrf, "keywordWeight", &out->rrf.keyword_weight);
    get_f64(env, rrf, "bothBonus", &out->rrf.both_bonus);
  }
  return true;
}

template <typename T>
bool typed(napi_env env, napi_value v, napi_typedarray_type want, const T** data, size_t* len) {
  napi_typedarray_type t;
  void* p = nullptr;
  if (napi_get_typedarray_info(env, v, &t, len, &p, nullptr, nullptr) != napi_ok || t != want) return false;
  *data = static_cast<const T*>(p);
  return true;
}

// ---- one hybrid search (direct or through a batcher) as async work --------------------------------
struct SearchWork {
  napi_async_work work = nullptr;
  napi_deferred deferred = nullptr;
  IndexHandle* h = nullptr;        // direct call: the handle it holds a reference on
  BatcherHandle* bh = nullptr;     // through a batcher
  rag_fused_out out;
  std::vector<float> q;
  uint32_t B = 1;
  rag_hybrid_opts opts;
  std::vector<uint64_t> kw_keys;
  std::vector<uint32_t> kw_counts;
  uint32_t cap = 0;
  std::vector<uint64_t> keys, vec_ids;
  std::vector<double> scores, vec_scores;
  std::vector<uint8_t> source, ctype, used_rrf, certified;
  std::vector<uint32_t> counts, vec_counts;
  int rc = RAG_OK;
  std::string err;
};

// the result arrays of a request, sized on the main thread before anything else can touch them
void prepare_out(SearchWork* w) {
  const uint32_t k = w->opts.vector_top_k;
  w->cap = k + w->opts.keyword_limit + w->opts.fresh_limit;
  w->keys.resize((size_t)w->B * w->cap); w->scores.resize((size_t)w->B * w->cap);
  w->source.resize((size_t)w->B * w->cap); w->ctype.resize((size_t)w->B * w->cap);
  w->counts.resize(w->B); w->used_rrf.resize(w->B); w->certified.resize(w->B);
  w->vec_ids.resize((size_t)w->B * k); w->vec_scores.resize((size_t)w->B * k); w->vec_counts.resize(w->B);
  w->out = {w->cap, w->keys.data(), w->scores.data(), w->source.data(), w->ctype.data(), w->counts.data(),
            w->used_rrf.data(), w->vec_ids.data(), w->vec_scores.data(), w->vec_counts.data(), w->certified.data()};
}

void search_execute(napi_env, void* data) {  // pool thread: the only place that touches CUDA
  auto* w = static_cast<SearchWork*>(data);
  if (w->bh)  // the batcher's buffers were all in flight: this request waits for one on a pool thread (back-pressure)
    w->rc = rag_batcher_submit(w->bh->b, w->q.data(), w->kw_keys.data(), w->kw_counts.empty() ? 0 : w->kw_counts[0], &w->out);
  else
    w->rc = rag_hybrid_search(w->h->idx, w->q.data(), w->B, &w->opts, w->kw_keys.data(), w->kw_counts.data(), &w->out);
  if (w->rc != RAG_OK) w->err = rag_last_error();
}

template <typename T>
napi_value make_typed(napi_env env, napi_typedarray_type t, const std::vector<T>& v) {
  napi_value ab, ta;
  void* p = nullptr;
  napi_create_arraybuffer(env, v.size() * sizeof(T), &p, &ab);
  if (!v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
  napi_create_typedarray(env, t, v.size(), ab, 0, &ta);
  return ta;
}

void search_complete(napi_env env, napi_status, void* data) {  // main thread: build the JS result
  auto* w = static_cast<SearchWork*>(data);
  if (w->rc != RAG_OK) {
    napi_value msg, e;
    std::string m = "libragera error " + std::to_string(w->rc) + ": " + w->err;
    napi_create_string_utf8(env, m.c_str(), NAPI_AUTO_LENGTH, &msg);
    napi_create_error(env, nullptr, msg, &e);
    napi_reject_deferred(env, w->deferred, e);  // hybridSearch does not catch; its callers do (engine.ts:295)
  } else {
    napi_value r, v;
    napi_create_object(env, &r);
    napi_create_uint32(env, w->cap, &v);                                         napi_set_named_property(env, r, "capacity", v);
    napi_set_named_property(env, r, "keys", make_typed(env, napi_biguint64_array, w->keys));
    napi_set_named_property(env, r, "scores", make_typed(env, napi_float64_array, w->scores));
    napi_set_named_property(env, r, "source", make_typed(env, napi_uint8_array, w->source));
    napi_set_named_property(env, r, "contentType", make_typed(env, napi_uint8_array, w->ctype));
    napi_set_named_property(env, r, "counts", make_typed(env, napi_uint32_array, w->counts));
    napi_set_named_property(env, r, "usedRrf", make_typed(env, napi_uint8_array, w->used_rrf));
    napi_set_named_property(env, r, "certified", make_typed(env, napi_uint8_array, w->certified));
    napi_set_named_property(env, r, "vecIds", make_typed(env, napi_biguint64_array, w->vec_ids));
    napi_set_named_property(env, r, "vecScores", make_typed(env, napi_float64_array, w->vec_scores));
    napi_set_named_property(env, r, "vecCounts", make_typed(env, napi_uint32_array, w->vec_counts));
    napi_resolve_deferred(env, w->deferred, r);
  }
  if (w->work) napi_delete_async_work(env, w->work);
  if (w->bh) batcher_leave(env, w->bh);   // the last request of a closing handle destroys it
  if (w->h) { w->h->pending--; index_settle(w->h); }
  delete w;
}

// queue the request as pool work; `promise` = the Promise made by the caller, or null to make one here
napi_value queue_search(napi_env env, SearchWork* w, napi_value promise = nullptr) {
  napi_value name;
  if (!promise) NAPI_OK(env, napi_create_promise(env, &w->deferred, &promise));
  NAPI_OK(env, napi_create_string_utf8(env, "ragera.hybridSearch", NAPI_AUTO_LENGTH, &name));
  NAPI_OK(env, napi_create_async_work(env, nullptr, name, search_execute, search_complete, w, &w->work));
  NAPI_OK(env, napi_queue_async_work(env, w->work));
  if (w->bh) batcher_enter(env, w->bh);
  if (w->h) w->h->pending++;
  return promise;
}

// ---- asynchronous submit: batcher worker thread -> thread-safe function -> main thread -------------------------------
void submit_done(void* user, int rc, const char* err) {   // batcher worker thread; the result is already in w->out's arrays
  auto* w = static_cast<SearchWork*>(user);
  w->rc = rc;
  if (rc != RAG_OK) w->err = err ? err : "";
  // (napi_closing: the environment is being torn down and nobody is left to resolve for — the request is simply dropped)
  if (napi_call_threadsafe_function(w->bh->tsfn, w, napi_tsfn_blocking) != napi_ok) delete w;
}
void submit_call_js(napi_env env, napi_value, void*, void* data) {   // main thread
  auto* w = static_cast<SearchWork*>(data);
  if (!env) { delete w; return; }   // the environment is going down: nobody is left to resolve for
  search_complete(env, napi_ok, w);
}

// ---- retriever.retrieve (SimpleVectorStore.query) and MemoryStore.retrieve as async work -----------------------------
struct TopkWork {
  napi_async_work work = nullptr;
  napi_deferred deferred = nullptr;
  IndexHandle* h = nullptr;
  std::vector<float> q;
  uint32_t B = 1;
  bool memory = false;
  rag_search_opts sopts;
  rag_memory_opts mopts;
  std::vector<uint64_t> ids;
  std::vector<double> scores, relevance, freshness;
  std::vector<uint32_t> counts;
  std::vector<uint8_t> certified;
  int rc = RAG_OK;
  std::string err;
};

void topk_execute(napi_env, void* data) {
  auto* w = static_cast<TopkWork*>(data);
  const uint32_t width = w->memory ? w->mopts.limit : w->sopts.k;
  w->ids.resize((size_t)w->B * width); w->scores.resize((size_t)w->B * width); w->counts.resize(w->B);
  if (w->memory) {
    w->relevance.resize((size_t)w->B * width); w->freshness.resize((size_t)w->B * width);
    rag_memory_out out = {w->ids.data(), w->scores.data(), w->relevance.data(), w->freshness.data(), w->counts.data()};
    w->rc = rag_memory_retrieve(w->h->idx, w->q.data(), w->B, &w->mopts, &out);
  } else {
    w->certified.resize(w->B);
    rag_topk_out out = {w->ids.data(), w->scores.data(), w->counts.data(), w->certified.data()};
    w->rc = rag_search(w->h->idx, w->q.data(), w->B, &w->sopts, &out);
  }
  if (w->rc != RAG_OK) w->err = rag_last_error();
}

void topk_complete(napi_env env, napi_status, void* data) {
  auto* w = static_cast<TopkWork*>(data);
  if (w->rc != RAG_OK) {
    napi_value msg, e;
    std::string m = "libragera error " + std::to_string(w->rc) + ": " + w->err;
    napi_create_string_utf8(env, m.c_str(), NAPI_AUTO_LENGTH, &msg);
    napi_create_error(env, nullptr, msg, &e);
    napi_reject_deferred(env, w->deferred, e);
  } else {
    napi_value r;
    napi_create_object(env, &r);
    napi_set_named_property(env, r, "ids", make_typed(env, napi_biguint64_array, w->ids));        // [B][k] chunk ids, rank order
    napi_set_named_property(env, r, "scores", make_typed(env, napi_float64_array, w->scores));    // similarities (or the 0.7/0.3 blend)
    napi_set_named_property(env, r, "counts", make_typed(env, napi_uint32_array, w->counts));
    if (w->memory) {
      napi_set_named_property(env, r, "relevance", make_typed(env, napi_float64_array, w->relevance));
      napi_set_named_property(env, r, "freshness", make_typed(env, napi_float64_array, w->freshness));
    } else {
      napi_set_named_property(env, r, "certified", make_typed(env, napi_uint8_array, w->certified));
    }
    napi_resolve_deferred(env, w->deferred, r);
  }
  napi_delete_async_work(env, w->work);
  w->h->pending--;
  index_settle(w->h);
  delete w;
}

napi_value queue_topk(napi_env env, TopkWork* w) {
  napi_value promise, name;
  NAPI_OK(env, napi_create_promise(env, &w->deferred, &promise));
  NAPI_OK(env, napi_create_string_utf8(env, "ragera.search", NAPI_AUTO_LENGTH, &name));
  NAPI_OK(env, napi_create_async_work(env, nullptr, name, topk_execute, topk_complete, w, &w->work));
  NAPI_OK(env, napi_queue_async_work(env, w->work));
  w->h->pending++;
  return promise;
}

// ---- exported functions ------------------------------------------------------------------------------
napi_value CreateIndex(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value argv[1];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  rag_index_desc d;
  memset(&d, 0, sizeof d);
  uint32_t rows = 0, dev = 0, shadow = 0, shadow16 = 0;
  get_u32(env, argv[0], "rows", &rows);
  get_u32(env, argv[0], "dim", &d.dim);
  get_u32(env, argv[0], "device", &dev);
  get_u32(env, argv[0], "bf16Shadow", &shadow);
  get_u32(env, argv[0], "f16Shadow", &shadow16);
  d.capacity_rows = rows;
  d.device = (int32_t)dev;
  d.dtype = RAG_F32;
  d.flags = shadow16 ? RAG_INDEX_F16_SHADOW : (shadow ? RAG_INDEX_BF16_SHADOW : 0);
  rag_index* idx = nullptr;
  const int rc = rag_index_create(&d, &idx);
  if (rc != RAG_OK) return throw_rag(env, rc);
  auto* h = new IndexHandle();
  h->idx = idx;
  napi_value ext;
  if (napi_create_external(env, h, index_finalize, nullptr, &ext) != napi_ok) {
    rag_index_destroy(idx);
    delete h;
    napi_throw_error(env, nullptr, "napi_create_external failed");
    return nullptr;
  }
  return ext;
}

napi_value UploadRows(napi_env env, napi_callback_info info) {  // uploadRows(handle, Float32Array rows, nrows, row0?)
  size_t argc = 4;
  napi_value argv[4];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  IndexHandle* h = index_arg(env, argv[0]);
  if (!h) return nullptr;
  rag_index* idx = h->idx;
  const float* rows = nullptr;
  size_t n = 0;
  if (!typed(env, argv[1], napi_float32_array, &rows, &n)) { napi_throw_type_error(env, nullptr, "rows must be a Float32Array"); return nullptr; }
  uint32_t nrows = 0, r0 = 0;
  NAPI_OK(env, napi_get_value_uint32(env, argv[2], &nrows));
  uint64_t row0 = rag_index_rows(idx);  // index.insert appends (src/lib/memory/store.ts:67)
  if (argc > 3 && napi_get_value_uint32(env, argv[3], &r0) == napi_ok) row0 = r0;
  const int rc = rag_index_upload(idx, row0, nrows, rows);
  if (rc != RAG_OK) return throw_rag(env, rc);
  napi_value out;
  NAPI_OK(env, napi_create_double(env, (double)row0, &out));
  return out;
}

napi_value LoadVectorStore(napi_env env, napi_callback_info info) {  // loadVectorStore(handle, path) -> string[]
  size_t argc = 2;
  napi_value argv[2];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  IndexHandle* h = index_arg(env, argv[0]);
  if (!h) return nullptr;
  rag_index* idx = h->idx;
  char path[4096];
  size_t len = 0;
  NAPI_OK(env, napi_get_value_string_utf8(env, argv[1], path, sizeof path, &len));
  uint64_t rows = 0, bytes = 0;
  char* ids = nullptr;
  // through the binary sidecar (<path>.ragera) when it is fresh, else the JSON is parsed and the sidecar rewritten;
  // contentType per row comes from the file's metadataDict either way
  int from_cache = 0;
  const int rc = rag_index_open_store(idx, path, nullptr, &rows, &ids, &bytes, &from_cache);
  if (rc != RAG_OK) return throw_rag(env, rc);
  napi_value arr;
  NAPI_OK(env, napi_create_array_with_length(env, (size_t)rows, &arr));
  const char* p = ids;
  for (uint64_t i = 0; i < rows; i++) {
    napi_value s;
    napi_create_string_utf8(env, p, NAPI_AUTO_LENGTH, &s);
    napi_set_element(env, arr, (uint32_t)i, s);
    p += strlen(p) + 1;
  }
  rag_free(ids);
  return arr;
}

napi_value SetRowMeta(napi_env env, napi_callback_info info) {
  size_t argc = 6;
  napi_value argv[6];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  IndexHandle* h = index_arg(env, argv[0]);
  if (!h) return nullptr;
  rag_index* idx = h->idx;
  uint32_t row0 = 0;
  NAPI_OK(env, napi_get_value_uint32(env, argv[1], &row0));
  const uint8_t* ct = nullptr; const double* conf = nullptr; const int32_t* acc = nullptr; const int64_t* last = nullptr;
  size_t n = 0, m = 0;
  if (!typed(env, argv[2], napi_uint8_array, &ct, &n)) { napi_throw_type_error(env, nullptr, "contentType must be a Uint8Array"); return nullptr; }
  if (argc > 3) typed(env, argv[3], napi_float64_array, &conf, &m);
  if (argc > 4) typed(env, argv[4], napi_int32_array, &acc, &m);
  if (argc > 5) typed(env, argv[5], napi_bigint64_array, &last, &m);
  const int rc = rag_index_set_row_meta(idx, row0, n, ct, conf, acc, last);
  return rc == RAG_OK ? nullptr : throw_rag(env, rc);
}

napi_value SetRowKeys(napi_env env, napi_callback_info info) {
  size_t argc = 3;
  napi_value argv[3];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  IndexHandle* h = index_arg(env, argv[0]);
  if (!h) return nullptr;
  rag_index* idx = h->idx;
  uint32_t row0 = 0;
  NAPI_OK(env, napi_get_value_uint32(env, argv[1], &row0));
  const uint64_t* keys = nullptr;
  size_t n = 0;
  if (!typed(env, argv[2], napi_biguint64_array, &keys, &n)) { napi_throw_type_error(env, nullptr, "keys must be a BigUint64Array"); return nullptr; }
  const int rc = rag_index_set_row_keys(idx, row0, n, keys);
  return rc == RAG_OK ? nullptr : throw_rag(env, rc);
}

napi_value HybridSearch(napi_env env, napi_callback_info info) {
  size_t argc = 6;
  napi_value argv[6];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  IndexHandle* h = index_arg(env, argv[0]);
  if (!h) return nullptr;
  const float* q = nullptr;
  size_t nq = 0;
  if (!typed(env, argv[1], napi_float32_array, &q, &nq)) { napi_throw_type_error(env, nullptr, "queries must be a Float32Array"); return nullptr; }
  uint32_t B = 0;
  NAPI_OK(env, napi_get_value_uint32(env, argv[2], &B));
  auto* w = new SearchWork();  // owned by the async work from here on (search_complete deletes it)
  w->h = h;
  w->B = B;
  read_opts(env, argv[3], &w->opts);
  if ((size_t)B * w->opts.vector_top_k == 0 || nq % B != 0) { delete w; napi_throw_type_error(env, nullptr, "queries must hold B rows of dim values, B and vectorTopK >= 1"); return nullptr; }
  w->q.assign(q, q + nq);  // copied: the JS buffer may be reused before the worker runs
  const uint64_t* kk = nullptr; const uint32_t* kc = nullptr;
  size_t nk = 0, nc = 0;
  if (argc > 4 && typed(env, argv[4], napi_biguint64_array, &kk, &nk)) w->kw_keys.assign(kk, kk + nk);
  if (argc > 5 && typed(env, argv[5], napi_uint32_array, &kc, &nc)) w->kw_counts.assign(kc, kc + nc);
  w->kw_counts.resize(w->B, 0);
  w->kw_keys.resize((size_t)w->B * w->opts.keyword_limit + 1, 0);
  prepare_out(w);
  return queue_search(env, w);
}

// search(handle, Float32Array queries, B, k) -> Promise<{ids, scores, counts, certified}> — what a llamaindex
// BaseVectorStore.query({queryEmbedding, similarityTopK}) → {ids, similarities} binds (src/lib/hybrid-search.ts:223-224)
napi_value Search(napi_env env, napi_callback_info info) {
  size_t argc = 4;
  napi_value argv[4];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  IndexHandle* h = index_arg(env, argv[0]);
  if (!h) return nullptr;
  const float* q = nullptr;
  size_t nq = 0;
  if (!typed(env, argv[1], napi_float32_array, &q, &nq)) { napi_throw_type_error(env, nullptr, "queries must be a Float32Array"); return nullptr; }
  uint32_t B = 0, k = 0;
  NAPI_OK(env, napi_get_value_uint32(env, argv[2], &B));
  NAPI_OK(env, napi_get_value_uint32(env, argv[3], &k));
  if (B == 0 || k == 0 || nq % B != 0) { napi_throw_type_error(env, nullptr, "queries must hold B rows of dim values, B and k >= 1"); return nullptr; }
  auto* w = new TopkWork();
  w->h = h;
  w->B = B;
  w->q.assign(q, q + nq);
  memset(&w->sopts, 0, sizeof w->sopts);
  w->sopts.k = k;
  return queue_topk(env, w);
}

// memoryRetrieve(handle, Float32Array queries, B, {limit, minRelevance, nowMs, similarityTopK}) ->
//   Promise<{ids, scores, relevance, freshness, counts}> — MemoryStore.retrieve's filter/blend/sort (src/lib/memory/store.ts:119-175)
napi_value MemoryRetrieve(napi_env env, napi_callback_info info) {
  size_t argc = 4;
  napi_value argv[4];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  IndexHandle* h = index_arg(env, argv[0]);
  if (!h) return nullptr;
  const float* q = nullptr;
  size_t nq = 0;
  if (!typed(env, argv[1], napi_float32_array, &q, &nq)) { napi_throw_type_error(env, nullptr, "queries must be a Float32Array"); return nullptr; }
  uint32_t B = 0;
  NAPI_OK(env, napi_get_value_uint32(env, argv[2], &B));
  rag_memory_opts mo;
  memset(&mo, 0, sizeof mo);
  mo.limit = 10;            // store.ts:104
  mo.min_relevance = 0.5;   // store.ts:105
  double now_ms = 0.0;
  get_u32(env, argv[3], "limit", &mo.limit);
  get_f64(env, argv[3], "minRelevance", &mo.min_relevance);
  get_f64(env, argv[3], "nowMs", &now_ms);
  get_u32(env, argv[3], "similarityTopK", &mo.similarity_top_k);
  mo.now_ms = (int64_t)now_ms;
  if (B == 0 || mo.limit == 0 || nq % B != 0) { napi_throw_type_error(env, nullptr, "queries must hold B rows of dim values, B and limit >= 1"); return nullptr; }
  auto* w = new TopkWork();
  w->h = h;
  w->B = B;
  w->memory = true;
  w->mopts = mo;
  w->q.assign(q, q + nq);
  return queue_topk(env, w);
}

napi_value CreateBatcher(napi_env env, napi_callback_info info) {  // createBatcher(handle, opts, maxBatch, maxWaitUs)
  size_t argc = 4;
  napi_value argv[4];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  IndexHandle* h = index_arg(env, argv[0]);
  if (!h) return nullptr;
  rag_batcher_desc d;
  read_opts(env, argv[1], &d.opts);
  d.max_batch = 1024;
  d.max_wait_us = 1000;   // a cap: a batch goes out earlier, once the arrivals pause
  if (argc > 2) napi_get_value_uint32(env, argv[2], &d.max_batch);
  if (argc > 3) napi_get_value_uint32(env, argv[3], &d.max_wait_us);
  auto* bh = new BatcherHandle();
  napi_value name, ext;
  // the channel from the batcher's worker threads back to the event loop; unreferenced while no request is out, so an idle
  // batcher does not keep the process alive
  if (napi_create_string_utf8(env, "ragera.submit", NAPI_AUTO_LENGTH, &name) != napi_ok ||
      napi_create_threadsafe_function(env, nullptr, nullptr, name, 0, 1, nullptr, nullptr, nullptr, submit_call_js, &bh->tsfn) != napi_ok) {
    delete bh;
    napi_throw_error(env, nullptr, "napi_create_threadsafe_function failed");
    return nullptr;
  }
  napi_unref_threadsafe_function(env, bh->tsfn);
  const int rc = rag_batcher_create(h->idx, &d, &bh->b);
  if (rc != RAG_OK || napi_create_external(env, bh, batcher_finalize, nullptr, &ext) != napi_ok) {
    if (bh->b) rag_batcher_destroy(bh->b);
    napi_release_threadsafe_function(bh->tsfn, napi_tsfn_release);
    delete bh;
    if (rc != RAG_OK) return throw_rag(env, rc);
    napi_throw_error(env, nullptr, "napi_create_external failed");
    return nullptr;
  }
  bh->owner = h;
  h->pending++;   // the index outlives its batchers
  return ext;
}

// submit(batcher, opts, Float32Array q, BigUint64Array kwKeys) -> Promise<result>. Runs on the main thread and does not
// block: the request takes a slot of the batcher's open batch (rag_batcher_submit_async copies the inputs) and the
// Promise is resolved from the batcher's worker through the thread-safe function. Only when every batch buffer is in
// flight (RAG_ERR_BUSY) does the request fall back to a pool thread that waits for one — that is the back-pressure.
napi_value Submit(napi_env env, napi_callback_info info) {
  size_t argc = 4;
  napi_value argv[4];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  BatcherHandle* bh = batcher_arg(env, argv[0]);
  if (!bh) return nullptr;
  const float* q = nullptr;
  size_t nq = 0;
  if (!typed(env, argv[2], napi_float32_array, &q, &nq)) { napi_throw_type_error(env, nullptr, "query must be a Float32Array"); return nullptr; }
  auto* w = new SearchWork();
  w->bh = bh;
  read_opts(env, argv[1], &w->opts);  // must equal the batcher's options (sizes the result arrays)
  w->q.assign(q, q + nq);
  const uint64_t* kk = nullptr;
  size_t nk = 0;
  if (argc > 3 && typed(env, argv[3], napi_biguint64_array, &kk, &nk)) w->kw_keys.assign(kk, kk + nk);
  w->kw_counts.assign(1, (uint32_t)nk);
  w->kw_keys.resize(nk + 1, 0);
  w->B = 1;
  prepare_out(w);
  napi_value promise;
  if (napi_create_promise(env, &w->deferred, &promise) != napi_ok) { delete w; napi_throw_error(env, nullptr, "napi_create_promise failed"); return nullptr; }
  const int rc = rag_batcher_submit_async(bh->b, w->q.data(), w->kw_keys.data(), (uint32_t)nk, &w->out, submit_done, w);
  // (a completion that is already on its way is delivered by the event loop, i.e. after this function has returned)
  if (rc == RAG_OK) { batcher_enter(env, bh); return promise; }
  if (rc == RAG_ERR_BUSY) return queue_search(env, w, promise);
  w->rc = rc;
  w->err = rag_last_error();
  batcher_enter(env, bh);
  search_complete(env, napi_ok, w);   // rejects the Promise with rag_last_error(), leaves the batcher, deletes w
  return promise;
}

// destroy(handle): closes the handle; the native index goes away with the last call still in flight on it (or now)
napi_value Destroy(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value argv[1];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  IndexHandle* h = nullptr;
  NAPI_OK(env, napi_get_value_external(env, argv[0], (void**)&h));
  if (!h) return nullptr;
  h->closing = true;
  index_settle(h);
  return nullptr;
}

napi_value DestroyBatcher(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value argv[1];
  NAPI_OK(env, napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr));
  BatcherHandle* bh = nullptr;
  NAPI_OK(env, napi_get_value_external(env, argv[0], (void**)&bh));
  if (!bh) return nullptr;
  bh->closing = true;
  batcher_settle(env, bh);
  return nullptr;
}

}  // namespace

NAPI_MODULE_INIT() {
  napi_property_descriptor props[] = {
      {"createIndex", nullptr, CreateIndex, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"uploadRows", nullptr, UploadRows, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"loadVectorStore", nullptr, LoadVectorStore, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"setRowMeta", nullptr, SetRowMeta, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"setRowKeys", nullptr, SetRowKeys, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"hybridSearch", nullptr, HybridSearch, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"search", nullptr, Search, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"memoryRetrieve", nullptr, MemoryRetrieve, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"createBatcher", nullptr, CreateBatcher, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"submit", nullptr, Submit, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"destroy", nullptr, Destroy, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"destroyBatcher", nullptr, DestroyBatcher, nullptr, nullptr, nullptr, napi_default, nullptr},
  };
  napi_define_properties(env, exports, sizeof(props) / sizeof(props[0]), props);
  return exports;
}
