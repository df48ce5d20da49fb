// napi_mock.cc — TEST HARNESS: a minimal in-process Node-API host, so that the N-API addon
// (integration/node/ragera_addon.cc) can be loaded and driven without Node (there is no Node in the
// build container or on the GPU box). It implements the subset of Node-API declared in node_api_stub/node_api.h over
// plain C++ objects — numbers, strings, objects, arrays, externals, ArrayBuffers, typed arrays, promises, async work
// (execute on a worker thread, complete on the "main" thread, like libuv's pool) — and a driver that plays the calls
// native-retrieval.ts makes: createIndex → uploadRows → setRowMeta → setRowKeys → hybridSearch (Promise) → createBatcher /
// submit → search / memoryRetrieve → a rejected Promise → destroy. Results are printed as JSON; tests/test_gpu_napi.py compares them with the oracle.
//
// usage: napi_mock <input.bin>   (layout written by the test: see read_input)
#include <node_api.h>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <memory>
#include <string>
#include <thread>
#include <vector>

struct napi_value__ {
  enum Kind { Undefined, Number, String, Object, Array, External, ArrayBuffer, TypedArray, Promise, Error, Function } kind = Undefined;
  double num = 0;
  std::string str;
  std::map<std::string, napi_value> props;
  std::vector<napi_value> elems;
  void* ext = nullptr;
  std::vector<uint8_t> bytes;  // ArrayBuffer storage
  napi_typedarray_type ta_type = napi_uint8_array;
  size_t ta_len = 0, ta_off = 0;
  napi_value ta_buf = nullptr;
  int state = 0;  // Promise: 0 pending, 1 fulfilled, 2 rejected
  napi_value settled = nullptr;
  napi_callback fn = nullptr;
  napi_finalize finalize = nullptr;   // External: run when the "garbage collector" (end of main) collects it
};
struct napi_callback_info__ { std::vector<napi_value> args; };
struct napi_deferred__ { napi_value promise; };
struct napi_async_work__ { napi_async_execute_callback execute; napi_async_complete_callback complete; void* data; };
// thread-safe function: any thread enqueues, the "main" thread (run_event_loop) calls call_js for every item
struct napi_threadsafe_function__ {
  napi_env env = nullptr;
  napi_threadsafe_function_call_js call_js = nullptr;
  void* context = nullptr;
  std::deque<void*> items;   // guarded by env->mu
  bool referenced = true, released = false;
  uint64_t delivered = 0;
};
struct napi_env__ {
  std::vector<std::unique_ptr<napi_value__>> heap;
  bool pending = false;
  std::string exception;
  std::deque<napi_async_work> queue;
  std::mutex mu;                      // thread-safe function queues
  std::condition_variable cv;
  std::vector<napi_threadsafe_function> tsfns;
  napi_value make(napi_value__::Kind k) {
    heap.emplace_back(new napi_value__());
    heap.back()->kind = k;
    return heap.back().get();
  }
};

static size_t elem_size(napi_typedarray_type t) {
  switch (t) {
    case napi_int8_array: case napi_uint8_array: case napi_uint8_clamped_array: return 1;
    case napi_int16_array: case napi_uint16_array: return 2;
    case napi_int32_array: case napi_uint32_array: case napi_float32_array: return 4;
    default: return 8;
  }
}

extern "C" {
napi_status napi_get_cb_info(napi_env, napi_callback_info info, size_t* argc, napi_value* argv, napi_value* this_arg, void** data) {
  const size_t cap = argc ? *argc : 0;
  for (size_t i = 0; i < cap; i++) argv[i] = i < info->args.size() ? info->args[i] : nullptr;  // missing args read as undefined
  if (argc) *argc = info->args.size();
  if (this_arg) *this_arg = nullptr;
  if (data) *data = nullptr;
  return napi_ok;
}
napi_status napi_throw_error(napi_env env, const char*, const char* msg) { env->pending = true; env->exception = msg; return napi_ok; }
napi_status napi_throw_type_error(napi_env env, const char*, const char* msg) { env->pending = true; env->exception = std::string("TypeError: ") + msg; return napi_ok; }
napi_status napi_create_error(napi_env env, napi_value, napi_value msg, napi_value* result) {
  *result = env->make(napi_value__::Error);
  (*result)->str = msg ? msg->str : "";
  return napi_ok;
}
napi_status napi_create_object(napi_env env, napi_value* result) { *result = env->make(napi_value__::Object); return napi_ok; }
napi_status napi_create_array_with_length(napi_env env, size_t length, napi_value* result) {
  *result = env->make(napi_value__::Array);
  (*result)->elems.resize(length, nullptr);
  return napi_ok;
}
napi_status napi_create_double(napi_env env, double value, napi_value* result) { *result = env->make(napi_value__::Number); (*result)->num = value; return napi_ok; }
napi_status napi_create_uint32(napi_env env, uint32_t value, napi_value* result) { return napi_create_double(env, value, result); }
napi_status napi_create_string_utf8(napi_env env, const char* str, size_t length, napi_value* result) {
  *result = env->make(napi_value__::String);
  (*result)->str = length == NAPI_AUTO_LENGTH ? std::string(str) : std::string(str, length);
  return napi_ok;
}
napi_status napi_create_external(napi_env env, void* data, napi_finalize fin, void*, napi_value* result) { *result = env->make(napi_value__::External); (*result)->ext = data; (*result)->finalize = fin; return napi_ok; }
napi_status napi_create_arraybuffer(napi_env env, size_t byte_length, void** data, napi_value* result) {
  *result = env->make(napi_value__::ArrayBuffer);
  (*result)->bytes.resize(byte_length);
  if (data) *data = (*result)->bytes.data();
  return napi_ok;
}
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value ab, size_t byte_offset, napi_value* result) {
  if (!ab || ab->kind != napi_value__::ArrayBuffer || byte_offset + length * elem_size(type) > ab->bytes.size()) return napi_invalid_arg;
  *result = env->make(napi_value__::TypedArray);
  (*result)->ta_type = type; (*result)->ta_len = length; (*result)->ta_off = byte_offset; (*result)->ta_buf = ab;
  return napi_ok;
}
napi_status napi_get_typedarray_info(napi_env, napi_value v, napi_typedarray_type* type, size_t* length, void** data, napi_value* ab, size_t* byte_offset) {
  if (!v || v->kind != napi_value__::TypedArray) return napi_invalid_arg;
  if (type) *type = v->ta_type;
  if (length) *length = v->ta_len;
  if (data) *data = v->ta_buf->bytes.data() + v->ta_off;
  if (ab) *ab = v->ta_buf;
  if (byte_offset) *byte_offset = v->ta_off;
  return napi_ok;
}
napi_status napi_get_value_double(napi_env, napi_value v, double* result) { if (!v || v->kind != napi_value__::Number) return napi_number_expected; *result = v->num; return napi_ok; }
napi_status napi_get_value_uint32(napi_env, napi_value v, uint32_t* result) { if (!v || v->kind != napi_value__::Number) return napi_number_expected; *result = (uint32_t)v->num; return napi_ok; }
napi_status napi_get_value_external(napi_env, napi_value v, void** result) { if (!v || v->kind != napi_value__::External) return napi_invalid_arg; *result = v->ext; return napi_ok; }
napi_status napi_get_value_string_utf8(napi_env, napi_value v, char* buf, size_t bufsize, size_t* result) {
  if (!v || v->kind != napi_value__::String) return napi_string_expected;
  if (!buf) { if (result) *result = v->str.size(); return napi_ok; }
  const size_t n = bufsize ? std::min(bufsize - 1, v->str.size()) : 0;
  memcpy(buf, v->str.data(), n);
  if (bufsize) buf[n] = 0;
  if (result) *result = n;
  return napi_ok;
}
napi_status napi_set_element(napi_env, napi_value obj, uint32_t index, napi_value value) {
  if (!obj || obj->kind != napi_value__::Array) return napi_array_expected;
  if (index >= obj->elems.size()) obj->elems.resize(index + 1, nullptr);
  obj->elems[index] = value;
  return napi_ok;
}
napi_status napi_set_named_property(napi_env, napi_value obj, const char* name, napi_value value) { if (!obj) return napi_object_expected; obj->props[name] = value; return napi_ok; }
napi_status napi_get_named_property(napi_env, napi_value obj, const char* name, napi_value* result) {
  if (!obj) return napi_object_expected;
  auto it = obj->props.find(name);
  *result = it == obj->props.end() ? nullptr : it->second;
  return napi_ok;
}
napi_status napi_has_named_property(napi_env, napi_value obj, const char* name, bool* result) { if (!obj) return napi_object_expected; *result = obj->props.count(name) != 0; return napi_ok; }
napi_status napi_define_properties(napi_env env, napi_value obj, size_t n, const napi_property_descriptor* p) {
  for (size_t i = 0; i < n; i++) {
    napi_value f = env->make(napi_value__::Function);
    f->fn = p[i].method;
    obj->props[p[i].utf8name] = f;
  }
  return napi_ok;
}
napi_status napi_create_promise(napi_env env, napi_deferred* deferred, napi_value* promise) {
  *promise = env->make(napi_value__::Promise);
  *deferred = new napi_deferred__{*promise};
  return napi_ok;
}
napi_status napi_resolve_deferred(napi_env, napi_deferred d, napi_value v) { d->promise->state = 1; d->promise->settled = v; delete d; return napi_ok; }
napi_status napi_reject_deferred(napi_env, napi_deferred d, napi_value v) { d->promise->state = 2; d->promise->settled = v; delete d; return napi_ok; }
napi_status napi_create_async_work(napi_env, napi_value, napi_value, napi_async_execute_callback ex, napi_async_complete_callback co, void* data, napi_async_work* result) {
  *result = new napi_async_work__{ex, co, data};
  return napi_ok;
}
napi_status napi_queue_async_work(napi_env env, napi_async_work w) { env->queue.push_back(w); return napi_ok; }
napi_status napi_delete_async_work(napi_env, napi_async_work w) { delete w; return napi_ok; }

napi_status napi_create_threadsafe_function(napi_env env, napi_value, napi_value, napi_value, size_t, size_t, void*, napi_finalize, void* context,
                                            napi_threadsafe_function_call_js call_js, napi_threadsafe_function* result) {
  auto* f = new napi_threadsafe_function__();
  f->env = env; f->call_js = call_js; f->context = context;
  std::lock_guard<std::mutex> lk(env->mu);
  env->tsfns.push_back(f);
  *result = f;
  return napi_ok;
}
napi_status napi_call_threadsafe_function(napi_threadsafe_function f, void* data, napi_threadsafe_function_call_mode) {  // any thread
  {
    std::lock_guard<std::mutex> lk(f->env->mu);
    if (f->released) return napi_generic_failure;   // napi_closing in the real thing
    f->items.push_back(data);
  }
  f->env->cv.notify_all();
  return napi_ok;
}
napi_status napi_release_threadsafe_function(napi_threadsafe_function f, napi_threadsafe_function_release_mode) {
  std::lock_guard<std::mutex> lk(f->env->mu);
  f->released = true;   // kept in env->tsfns (freed with the environment): items already queued are still delivered
  return napi_ok;
}
napi_status napi_ref_threadsafe_function(napi_env env, napi_threadsafe_function f) { std::lock_guard<std::mutex> lk(env->mu); f->referenced = true; return napi_ok; }
napi_status napi_unref_threadsafe_function(napi_env env, napi_threadsafe_function f) { std::lock_guard<std::mutex> lk(env->mu); f->referenced = false; return napi_ok; }

napi_value napi_register_module_v1(napi_env env, napi_value exports);  // the addon's NAPI_MODULE_INIT
}

// ---- the "event loop": every queued work item executes on its own worker thread (concurrently, like the libuv pool),
//      then its completion runs here on the main thread; items of thread-safe functions are delivered here too. As in
//      Node, the loop stays alive while work is queued or a REFERENCED thread-safe function exists (the addon references
//      its function while requests are out) ------------------------------------------------------------------------
static uint64_t g_tsfn_delivered = 0;
static bool deliver_tsfn_items(napi_env env) {
  bool any = false;
  for (;;) {
    napi_threadsafe_function f = nullptr;
    void* item = nullptr;
    {
      std::lock_guard<std::mutex> lk(env->mu);
      for (napi_threadsafe_function c : env->tsfns)
        if (!c->items.empty()) { f = c; item = c->items.front(); c->items.pop_front(); break; }
    }
    if (!f) return any;
    f->call_js(env, nullptr, f->context, item);
    f->delivered++;
    g_tsfn_delivered++;
    any = true;
  }
}
static void run_event_loop(napi_env env) {
  const auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    if (!env->queue.empty()) {
      std::vector<napi_async_work> batch(env->queue.begin(), env->queue.end());
      env->queue.clear();
      std::vector<std::thread> th;
      for (napi_async_work w : batch) th.emplace_back([=] { w->execute(env, w->data); });
      for (auto& t : th) t.join();   // (a batcher frees its buffers from its own worker threads: pool threads parked in the
                                     //  blocking submit never wait for this thread)
      deliver_tsfn_items(env);
      for (napi_async_work w : batch) w->complete(env, napi_ok, w->data);
      continue;
    }
    if (deliver_tsfn_items(env)) continue;
    std::unique_lock<std::mutex> lk(env->mu);
    bool alive = false;
    for (napi_threadsafe_function f : env->tsfns) alive = alive || (f->referenced && !f->released) || !f->items.empty();
    if (!alive) return;
    env->cv.wait_for(lk, std::chrono::milliseconds(20));
    if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) { fprintf(stderr, "event loop: a referenced thread-safe function never went idle\n"); exit(5); }
  }
}

// ---- driver helpers -------------------------------------------------------------------------------------------------
static napi_value call(napi_env env, napi_value exports, const char* name, std::vector<napi_value> args) {
  auto it = exports->props.find(name);
  if (it == exports->props.end() || !it->second->fn) { fprintf(stderr, "addon does not export %s\n", name); exit(2); }
  napi_callback_info__ info{std::move(args)};
  env->pending = false;
  return it->second->fn(env, &info);
}
static napi_value num(napi_env env, double v) { napi_value r; napi_create_double(env, v, &r); return r; }
template <typename T>
static napi_value typed_array(napi_env env, napi_typedarray_type t, const T* p, size_t n) {
  napi_value ab, ta;
  void* d;
  napi_create_arraybuffer(env, n * sizeof(T), &d, &ab);
  if (n) memcpy(d, p, n * sizeof(T));
  napi_create_typedarray(env, t, n, ab, 0, &ta);
  return ta;
}
static napi_value object(napi_env env, std::vector<std::pair<const char*, napi_value>> kv) {
  napi_value o;
  napi_create_object(env, &o);
  for (auto& e : kv) o->props[e.first] = e.second;
  return o;
}
template <typename T>
static const T* view(napi_value obj, const char* name, size_t* n) {
  napi_value v = obj->props.at(name);
  *n = v->ta_len;
  return (const T*)(v->ta_buf->bytes.data() + v->ta_off);
}
static void print_result(napi_value r, bool last) {
  size_t n;
  const uint32_t cap = (uint32_t)r->props.at("capacity")->num;
  const uint32_t count = view<uint32_t>(r, "counts", &n)[0], vcount = view<uint32_t>(r, "vecCounts", &n)[0];
  const uint64_t* keys = view<uint64_t>(r, "keys", &n);
  const double* scores = view<double>(r, "scores", &n);
  const uint8_t* src = view<uint8_t>(r, "source", &n);
  const uint8_t* ct = view<uint8_t>(r, "contentType", &n);
  const uint64_t* vids = view<uint64_t>(r, "vecIds", &n);
  const double* vs = view<double>(r, "vecScores", &n);
  printf("{\"capacity\": %u, \"usedRrf\": %u, \"certified\": %u, \"keys\": [", cap, view<uint8_t>(r, "usedRrf", &n)[0], view<uint8_t>(r, "certified", &n)[0]);
  for (uint32_t i = 0; i < count; i++) printf("%s%llu", i ? ", " : "", (unsigned long long)keys[i]);
  printf("], \"scores\": [");
  for (uint32_t i = 0; i < count; i++) printf("%s%.17g", i ? ", " : "", scores[i]);
  printf("], \"source\": [");
  for (uint32_t i = 0; i < count; i++) printf("%s%u", i ? ", " : "", src[i]);
  printf("], \"contentType\": [");
  for (uint32_t i = 0; i < count; i++) printf("%s%u", i ? ", " : "", ct[i]);
  printf("], \"vecIds\": [");
  for (uint32_t i = 0; i < vcount; i++) printf("%s%llu", i ? ", " : "", (unsigned long long)vids[i]);
  printf("], \"vecScores\": [");
  for (uint32_t i = 0; i < vcount; i++) printf("%s%.17g", i ? ", " : "", vs[i]);
  printf("]}%s\n", last ? "" : ",");
}

struct input {
  uint32_t n, dim, B, k, kw_limit;
  double min_score;
  std::vector<float> rows, queries;
  std::vector<uint8_t> ctype;
  std::vector<uint64_t> row_keys, kw_keys;
  std::vector<uint32_t> kw_counts;
  std::vector<double> conf;
  std::vector<int32_t> acc;
  std::vector<int64_t> last;
  int64_t now_ms;
};
template <typename T>
static void read_vec(FILE* f, std::vector<T>& v, size_t n) {
  v.resize(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short input\n"); exit(2); }
}
// u32 n, dim, B, k, kw_limit, pad; f64 min_score; f32 rows[n][dim]; f32 queries[B][dim]; u8 ctype[n]; u64 row_keys[n];
// u64 kw_keys[B][kw_limit]; u32 kw_counts[B]; f64 confidence[n]; i32 access_count[n]; i64 last_access_ms[n]; i64 now_ms
static input read_input(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  input in;
  uint32_t h[6];
  if (fread(h, 4, 6, f) != 6 || fread(&in.min_score, 8, 1, f) != 1) { fprintf(stderr, "short input\n"); exit(2); }
  in.n = h[0]; in.dim = h[1]; in.B = h[2]; in.k = h[3]; in.kw_limit = h[4];
  read_vec(f, in.rows, (size_t)in.n * in.dim);
  read_vec(f, in.queries, (size_t)in.B * in.dim);
  read_vec(f, in.ctype, in.n);
  read_vec(f, in.row_keys, in.n);
  read_vec(f, in.kw_keys, (size_t)in.B * in.kw_limit);
  read_vec(f, in.kw_counts, in.B);
  read_vec(f, in.conf, in.n);
  read_vec(f, in.acc, in.n);
  read_vec(f, in.last, in.n);
  if (fread(&in.now_ms, 8, 1, f) != 1) { fprintf(stderr, "short input\n"); exit(2); }
  fclose(f);
  return in;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s input.bin\n", argv[0]); return 2; }
  const input in = read_input(argv[1]);
  napi_env__ env_storage;
  napi_env env = &env_storage;
  napi_value exports;
  napi_create_object(env, &exports);
  exports = napi_register_module_v1(env, exports);

  napi_value h = call(env, exports, "createIndex", {object(env, {{"rows", num(env, in.n)}, {"dim", num(env, in.dim)}, {"device", num(env, 0)}, {"bf16Shadow", num(env, 1)}})});
  if (env->pending) { fprintf(stderr, "createIndex threw: %s\n", env->exception.c_str()); return 3; }
  napi_value r0 = call(env, exports, "uploadRows", {h, typed_array(env, napi_float32_array, in.rows.data(), in.rows.size()), num(env, in.n)});
  if (env->pending || r0->num != 0) { fprintf(stderr, "uploadRows threw: %s\n", env->exception.c_str()); return 3; }
  call(env, exports, "setRowMeta", {h, num(env, 0), typed_array(env, napi_uint8_array, in.ctype.data(), in.n), typed_array(env, napi_float64_array, in.conf.data(), in.n),
                                    typed_array(env, napi_int32_array, in.acc.data(), in.n), typed_array(env, napi_bigint64_array, in.last.data(), in.n)});
  if (env->pending) { fprintf(stderr, "setRowMeta threw: %s\n", env->exception.c_str()); return 3; }
  call(env, exports, "setRowKeys", {h, num(env, 0), typed_array(env, napi_biguint64_array, in.row_keys.data(), in.n)});
  if (env->pending) { fprintf(stderr, "setRowKeys threw: %s\n", env->exception.c_str()); return 3; }

  auto opts = [&](uint32_t k) {
    return object(env, {{"vectorTopK", num(env, k)}, {"keywordLimit", num(env, in.kw_limit)}, {"minVectorScore", num(env, in.min_score)},
                        {"rrf", object(env, {{"k", num(env, 60)}, {"vectorWeight", num(env, 1)}, {"keywordWeight", num(env, 1)}, {"bothBonus", num(env, 0.1)}})}});
  };
  // every request is its own batch-1 hybridSearch Promise; here the loop is drained after each call — as `await` does in
  // hybridSearchNative (the concurrent case follows below)
  std::vector<napi_value> promises;
  for (uint32_t b = 0; b < in.B; b++) {
    napi_value p = call(env, exports, "hybridSearch", {h, typed_array(env, napi_float32_array, in.queries.data() + (size_t)b * in.dim, in.dim), num(env, 1), opts(in.k),
                                                       typed_array(env, napi_biguint64_array, in.kw_keys.data() + (size_t)b * in.kw_limit, in.kw_limit),
                                                       typed_array(env, napi_uint32_array, in.kw_counts.data() + b, 1)});
    if (env->pending || !p || p->kind != napi_value__::Promise) { fprintf(stderr, "hybridSearch threw: %s\n", env->exception.c_str()); return 3; }
    run_event_loop(env);
    promises.push_back(p);
  }
  // CONCURRENT requests on ONE handle (what a Next.js server does: parallel HTTP requests, the Promise.all of
  // engine.ts:108): all B hybridSearch calls and B retriever calls are queued before the loop is drained, so their
  // execute callbacks run on B+B worker threads at once; a synchronous mutator lands in the middle. libragera serialises
  // them per handle (ragera.h), so every result must equal the one-at-a-time result.
  std::vector<napi_value> inflight, inflight_topk;
  for (uint32_t b = 0; b < in.B; b++) {
    inflight.push_back(call(env, exports, "hybridSearch", {h, typed_array(env, napi_float32_array, in.queries.data() + (size_t)b * in.dim, in.dim), num(env, 1), opts(in.k),
                                                           typed_array(env, napi_biguint64_array, in.kw_keys.data() + (size_t)b * in.kw_limit, in.kw_limit),
                                                           typed_array(env, napi_uint32_array, in.kw_counts.data() + b, 1)}));
    inflight_topk.push_back(call(env, exports, "search", {h, typed_array(env, napi_float32_array, in.queries.data() + (size_t)b * in.dim, in.dim), num(env, 1), num(env, in.k)}));
    if (b == in.B / 2) call(env, exports, "setRowKeys", {h, num(env, 0), typed_array(env, napi_biguint64_array, in.row_keys.data(), in.n)});  // same keys again
  }
  run_event_loop(env);
  // the whole batch in ONE call (B queries)
  napi_value pb = call(env, exports, "hybridSearch", {h, typed_array(env, napi_float32_array, in.queries.data(), in.queries.size()), num(env, in.B), opts(in.k),
                                                      typed_array(env, napi_biguint64_array, in.kw_keys.data(), in.kw_keys.size()),
                                                      typed_array(env, napi_uint32_array, in.kw_counts.data(), in.B)});
  run_event_loop(env);
  // the micro-batcher: all B submits are queued first, so their worker threads block inside rag_batcher_submit TOGETHER
  napi_value bt = call(env, exports, "createBatcher", {h, opts(in.k), num(env, 64), num(env, 20000)});
  if (env->pending) { fprintf(stderr, "createBatcher threw: %s\n", env->exception.c_str()); return 3; }
  std::vector<napi_value> batched;
  for (uint32_t b = 0; b < in.B; b++)
    batched.push_back(call(env, exports, "submit", {bt, opts(in.k), typed_array(env, napi_float32_array, in.queries.data() + (size_t)b * in.dim, in.dim),
                                                    typed_array(env, napi_biguint64_array, in.kw_keys.data() + (size_t)b * in.kw_limit, in.kw_counts[b])}));
  const uint64_t delivered0 = g_tsfn_delivered;
  run_event_loop(env);
  const uint64_t via_tsfn = g_tsfn_delivered - delivered0;   // submits answered without a pool thread (rag_batcher_submit_async)
  // more requests than the batcher has slots (4 buffers x 4): the overflow falls back to pool threads (RAG_ERR_BUSY) and is
  // answered all the same; then destroyBatcher with requests still out: they are answered, the batcher goes away with the last
  napi_value bt_small = call(env, exports, "createBatcher", {h, opts(in.k), num(env, 4), num(env, 20000)});
  std::vector<napi_value> overflow;
  for (uint32_t r = 0; r < 3 * in.B; r++) {
    const uint32_t b = r % in.B;
    overflow.push_back(call(env, exports, "submit", {bt_small, opts(in.k), typed_array(env, napi_float32_array, in.queries.data() + (size_t)b * in.dim, in.dim),
                                                     typed_array(env, napi_biguint64_array, in.kw_keys.data() + (size_t)b * in.kw_limit, in.kw_counts[b])}));
  }
  call(env, exports, "destroyBatcher", {bt_small});
  call(env, exports, "submit", {bt_small, opts(in.k), typed_array(env, napi_float32_array, in.queries.data(), in.dim), typed_array(env, napi_biguint64_array, in.kw_keys.data(), 0)});
  const std::string closed_batcher_error = env->pending ? env->exception : "";
  run_event_loop(env);
  call(env, exports, "destroyBatcher", {bt});
  // destroy(handle) with searches in flight on it: a second index, B searches queued, destroyed before the loop runs them
  napi_value h2 = call(env, exports, "createIndex", {object(env, {{"rows", num(env, in.n)}, {"dim", num(env, in.dim)}, {"device", num(env, 0)}})});
  call(env, exports, "uploadRows", {h2, typed_array(env, napi_float32_array, in.rows.data(), in.rows.size()), num(env, in.n)});
  std::vector<napi_value> doomed;
  for (uint32_t b = 0; b < in.B; b++)
    doomed.push_back(call(env, exports, "search", {h2, typed_array(env, napi_float32_array, in.queries.data() + (size_t)b * in.dim, in.dim), num(env, 1), num(env, in.k)}));
  call(env, exports, "destroy", {h2});
  call(env, exports, "search", {h2, typed_array(env, napi_float32_array, in.queries.data(), in.dim), num(env, 1), num(env, in.k)});
  const std::string closed_index_error = env->pending ? env->exception : "";
  run_event_loop(env);
  // the retriever seam (NativeVectorStore.query) and MemoryStore.retrieve's device half, one Promise per query
  std::vector<napi_value> topk, mem;
  for (uint32_t b = 0; b < in.B; b++) {
    napi_value qv = typed_array(env, napi_float32_array, in.queries.data() + (size_t)b * in.dim, in.dim);
    topk.push_back(call(env, exports, "search", {h, qv, num(env, 1), num(env, in.k)}));
    run_event_loop(env);
    mem.push_back(call(env, exports, "memoryRetrieve", {h, qv, num(env, 1), object(env, {{"limit", num(env, in.k)}, {"similarityTopK", num(env, in.k)},
                                                                                        {"minRelevance", num(env, in.min_score)}, {"nowMs", num(env, (double)in.now_ms)}})}));
    run_event_loop(env);
  }
  // a failing call rejects the Promise with rag_last_error() (k beyond RAG_MAX_TOPK); a bad argument throws synchronously
  napi_value bad = call(env, exports, "hybridSearch", {h, typed_array(env, napi_float32_array, in.queries.data(), in.dim), num(env, 1), opts(1000),
                                                       typed_array(env, napi_biguint64_array, in.kw_keys.data(), 0), typed_array(env, napi_uint32_array, in.kw_counts.data(), 0)});
  run_event_loop(env);
  call(env, exports, "hybridSearch", {h, num(env, 1)});
  const std::string sync_error = env->pending ? env->exception : "";

  printf("{\"single\": [\n");
  for (uint32_t b = 0; b < in.B; b++) {
    if (promises[b]->state != 1) { fprintf(stderr, "promise %u not fulfilled: %s\n", b, promises[b]->settled ? promises[b]->settled->str.c_str() : "pending"); return 4; }
    print_result(promises[b]->settled, b + 1 == in.B);
  }
  printf("], \"in_flight\": [\n");
  for (uint32_t b = 0; b < in.B; b++) {
    if (inflight[b]->state != 1 || inflight_topk[b]->state != 1) { fprintf(stderr, "concurrent promise %u not fulfilled\n", b); return 4; }
    print_result(inflight[b]->settled, b + 1 == in.B);
  }
  printf("], \"batched\": [\n");
  for (uint32_t b = 0; b < in.B; b++) {
    if (batched[b]->state != 1) { fprintf(stderr, "batched promise %u not fulfilled: %s\n", b, batched[b]->settled ? batched[b]->settled->str.c_str() : "pending"); return 4; }
    print_result(batched[b]->settled, b + 1 == in.B);
  }
  if (pb->state != 1) { fprintf(stderr, "batch promise not fulfilled\n"); return 4; }
  size_t n;
  const uint32_t* counts = view<uint32_t>(pb->settled, "counts", &n);
  const uint32_t cap = (uint32_t)pb->settled->props.at("capacity")->num;
  const uint64_t* keys = view<uint64_t>(pb->settled, "keys", &n);
  const double* scores = view<double>(pb->settled, "scores", &n);
  printf("], \"one_call\": [\n");
  for (uint32_t b = 0; b < in.B; b++) {
    printf("{\"keys\": [");
    for (uint32_t i = 0; i < counts[b]; i++) printf("%s%llu", i ? ", " : "", (unsigned long long)keys[(size_t)b * cap + i]);
    printf("], \"scores\": [");
    for (uint32_t i = 0; i < counts[b]; i++) printf("%s%.17g", i ? ", " : "", scores[(size_t)b * cap + i]);
    printf("]}%s\n", b + 1 == in.B ? "" : ",");
  }
  printf("], \"search\": [\n");
  for (uint32_t b = 0; b < in.B; b++) {
    if (!topk[b] || topk[b]->state != 1 || !mem[b] || mem[b]->state != 1) { fprintf(stderr, "search/memoryRetrieve promise %u not fulfilled\n", b); return 4; }
    napi_value t = topk[b]->settled, m = mem[b]->settled;
    const uint32_t tc = view<uint32_t>(t, "counts", &n)[0], mc = view<uint32_t>(m, "counts", &n)[0];
    const uint64_t* ti = view<uint64_t>(t, "ids", &n); const double* ts = view<double>(t, "scores", &n);
    const uint64_t* mi = view<uint64_t>(m, "ids", &n); const double* ms = view<double>(m, "scores", &n);
    const double* mr = view<double>(m, "relevance", &n); const double* mf = view<double>(m, "freshness", &n);
    printf("{\"ids\": [");
    for (uint32_t i = 0; i < tc; i++) printf("%s%llu", i ? ", " : "", (unsigned long long)ti[i]);
    printf("], \"scores\": [");
    for (uint32_t i = 0; i < tc; i++) printf("%s%.17g", i ? ", " : "", ts[i]);
    printf("], \"certified\": %u, \"mem_ids\": [", view<uint8_t>(t, "certified", &n)[0]);
    for (uint32_t i = 0; i < mc; i++) printf("%s%llu", i ? ", " : "", (unsigned long long)mi[i]);
    printf("], \"mem_scores\": [");
    for (uint32_t i = 0; i < mc; i++) printf("%s%.17g", i ? ", " : "", ms[i]);
    printf("], \"mem_relevance\": [");
    for (uint32_t i = 0; i < mc; i++) printf("%s%.17g", i ? ", " : "", mr[i]);
    printf("], \"mem_freshness\": [");
    for (uint32_t i = 0; i < mc; i++) printf("%s%.17g", i ? ", " : "", mf[i]);
    printf("]}%s\n", b + 1 == in.B ? "" : ",");
  }
  std::string rej = bad && bad->kind == napi_value__::Promise && bad->state == 2 && bad->settled ? bad->settled->str : "";
  for (std::string* s : {&rej, const_cast<std::string*>(&sync_error)})
    for (char& c : *s) if (c == '"' || c == '\\' || c == '\n') c = ' ';
  printf("], \"overflow\": [\n");
  for (size_t r = 0; r < overflow.size(); r++) {
    if (!overflow[r] || overflow[r]->state != 1) { fprintf(stderr, "overflow promise %zu not fulfilled: %s\n", r, overflow[r] && overflow[r]->settled ? overflow[r]->settled->str.c_str() : "pending"); return 4; }
    print_result(overflow[r]->settled, r + 1 == overflow.size());
  }
  printf("], \"doomed\": [\n");
  for (uint32_t b = 0; b < in.B; b++) {
    if (!doomed[b] || doomed[b]->state != 1) { fprintf(stderr, "search %u on the destroyed handle was not answered\n", b); return 4; }
    napi_value t = doomed[b]->settled;
    const uint32_t tc = view<uint32_t>(t, "counts", &n)[0];
    const uint64_t* ti = view<uint64_t>(t, "ids", &n);
    printf("[");
    for (uint32_t i = 0; i < tc; i++) printf("%s%llu", i ? ", " : "", (unsigned long long)ti[i]);
    printf("]%s\n", b + 1 == in.B ? "" : ",");
  }
  printf("], \"via_tsfn\": %llu, \"closed_batcher\": \"%s\", \"closed_index\": \"%s\", \"rejected\": \"%s\", \"thrown\": \"%s\"}\n",
         (unsigned long long)via_tsfn, closed_batcher_error.c_str(), closed_index_error.c_str(), rej.c_str(), sync_error.c_str());
  call(env, exports, "destroy", {h});
  // "garbage collection" at shutdown: every External's finalizer runs (the wrappers are freed; anything still open is closed)
  for (auto& v : env->heap)
    if (v->kind == napi_value__::External && v->finalize) v->finalize(env, v->ext, nullptr);
  for (napi_threadsafe_function f : env->tsfns) delete f;
  return 0;
}
