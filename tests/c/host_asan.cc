// host_asan.cc — TEST HARNESS: the host-only loader (loader.cu) and sidecar (store_cache.cu) code compiled as plain C++ with
// -fsanitize=address,undefined and driven over a JSON store given on the command line: parse, metadata pass, sidecar write /
// read, then truncations and single-byte corruptions of the sidecar, every one of which must be rejected cleanly or — when it
// hits alignment padding — leave the data intact. No GPU and no CUDA runtime call is reached (the device entry points are
// stubbed). Run by tests/test_loader.py::test_loader_and_sidecar_under_sanitizers.
#include "common.cuh"
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
static thread_local char g_err[1024];
int rag_set_error(int code, const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); return code; }
extern "C" {
const char* rag_last_error(void) { return g_err; }
int rag_index_upload(rag_index*, uint64_t, uint64_t, const void*) { return RAG_ERR_NO_DEVICE; }
int rag_index_set_row_meta(rag_index*, uint64_t, uint64_t, const uint8_t*, const double*, const int32_t*, const int64_t*) { return RAG_ERR_NO_DEVICE; }
int rag_index_set_row_keys(rag_index*, uint64_t, uint64_t, const uint64_t*) { return RAG_ERR_NO_DEVICE; }
int rag_index_read_rows(rag_index*, uint64_t, uint64_t, void*) { return RAG_ERR_NO_DEVICE; }
}
// the sidecar's header checksum (store_cache.cu::fnv64 restated: the harness forges headers whose checksum is VALID, so
// the size checks — not the checksum — have to stop them)
static uint64_t fnv64_test(uint64_t h, const void* data, size_t n) {
  const unsigned char* p = (const unsigned char*)data;
  const uint64_t prime = 0x100000001b3ull;
  uint64_t l0 = h, l1 = h ^ 0x9e3779b97f4a7c15ull, l2 = h ^ 0xc2b2ae3d27d4eb4full, l3 = h ^ 0x165667b19e3779f9ull;
  for (; n >= 32; n -= 32, p += 32) {
    uint64_t w[4];
    memcpy(w, p, 32);
    l0 = (l0 ^ w[0]) * prime; l1 = (l1 ^ w[1]) * prime; l2 = (l2 ^ w[2]) * prime; l3 = (l3 ^ w[3]) * prime;
  }
  h = l0; h = (h ^ l1) * prime; h = (h ^ l2) * prime; h = (h ^ l3) * prime;
  for (; n >= 8; n -= 8, p += 8) { uint64_t w; memcpy(&w, p, 8); h = (h ^ w) * prime; }
  for (; n; n--, p++) h = (h ^ *p) * prime;
  return h;
}
static int on_rows(void* user, uint64_t, uint64_t n, const float* rows) { *(double*)user += rows[0] * (double)n; return RAG_OK; }
int main(int argc, char** argv) {
  // argv[1]: a JSON store, argv[2]: dim, argv[3]: a cache path to write/read/corrupt
  const char* json = argv[1]; uint32_t dim = (uint32_t)atoi(argv[2]); const char* cache = argv[3];
  (void)argc;
  double acc = 0; uint64_t rows = 0, nb = 0; char* ids = nullptr;
  int rc = rag_parse_vector_store_json(json, dim, 7, on_rows, &acc, &rows, &ids, &nb);
  printf("parse rc=%d rows=%llu err=%s\n", rc, (unsigned long long)rows, rc ? g_err : "");
  if (rc == RAG_OK) {
    std::vector<uint8_t> ct(rows ? rows : 1); char* mem = nullptr; uint64_t mb = 0; int found = 0;
    rc = rag_parse_vector_store_metadata(json, ids, nb, rows, ct.data(), &mem, &mb, &found);
    printf("meta rc=%d found=%d\n", rc, found);
    rag_free(mem);
    std::vector<float> X((size_t)rows * dim, 0.5f);
    rc = rag_cache_write_host(cache, RAG_F32, dim, rows, X.data(), nullptr, nullptr, nullptr, nullptr, nullptr, ids, nb, json);
    printf("write rc=%d fresh=%d\n", rc, rag_cache_is_fresh(cache, json));
    std::vector<float> Y((size_t)rows * dim); char* ids2 = nullptr; uint64_t nb2 = 0;
    rc = rag_cache_read_host(cache, 0, rows, Y.data(), nullptr, nullptr, nullptr, nullptr, nullptr, &ids2, &nb2);
    printf("read rc=%d same=%d\n", rc, (int)(rc == RAG_OK && nb2 == nb && (X.empty() || memcmp(X.data(), Y.data(), X.size() * 4) == 0)));
    rag_free(ids2);
    {  // the refresh routes: rebuild, fresh, and — after the JSON grew by whitespace — the in-place extension (0 new rows)
      const std::string c2 = std::string(cache) + ".refresh";
      int route = -1; uint64_t nr = 0;
      int r0 = rag_cache_refresh_host(c2.c_str(), json, RAG_BF16, dim, &route, &nr);
      printf("refresh rc=%d route=%d rows=%llu\n", r0, route, (unsigned long long)nr);
      r0 = rag_cache_refresh_host(c2.c_str(), json, RAG_BF16, dim, &route, &nr);
      printf("refresh rc=%d route=%d\n", r0, route);
      FILE* a = fopen(json, "ab"); fputc('\n', a); fclose(a);
      r0 = rag_cache_refresh_host(c2.c_str(), json, RAG_BF16, dim, &route, &nr);
      printf("refresh rc=%d route=%d fresh=%d\n", r0, route, rag_cache_is_fresh(c2.c_str(), json));
      uint64_t hsh = 0; int hok = 0;
      rag_file_prefix_hash(json, 1ull << 40, &hsh, &hok);
      printf("prefix past the end ok=%d\n", hok);
    }
    // truncate / corrupt at many offsets: must fail cleanly, never crash
    FILE* f = fopen(cache, "rb"); std::string blob; char buf[65536]; size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) blob.append(buf, n);
    fclose(f);
    int bad = 0, ok = 0;
    for (size_t cut : {(size_t)0, (size_t)1, (size_t)64, (size_t)127, (size_t)128, blob.size() / 2, blob.size() - 1}) {
      if (cut > blob.size() || blob.empty()) continue;
      f = fopen(cache, "wb"); fwrite(blob.data(), 1, cut, f); fclose(f);
      rag_cache_info info;
      (rag_cache_info_read(cache, &info) == RAG_OK ? ok : bad)++;
    }
    for (size_t off = 0; off < blob.size(); off += blob.size() / 97 + 1) {
      std::string b = blob; b[off] ^= 0x5A;
      f = fopen(cache, "wb"); fwrite(b.data(), 1, b.size(), f); fclose(f);
      char* i3 = nullptr; uint64_t n3 = 0;
      int r = rag_cache_read_host(cache, 0, rows, Y.data(), nullptr, nullptr, nullptr, nullptr, nullptr, &i3, &n3);
      (r == RAG_OK ? ok : bad)++;
      if (r == RAG_OK) {
        if (n3 != nb || (nb && memcmp(i3, ids, nb) != 0) || (!X.empty() && memcmp(X.data(), Y.data(), X.size() * 4) != 0)) printf("ACCEPTED CORRUPT DATA at offset %zu\n", off);
        rag_free(i3);
      }
    }
    printf("corruptions: rejected=%d accepted=%d\n", bad, ok);
    // forged headers with a VALID checksum and sizes chosen to wrap the section arithmetic (rows*dim*4, rows*8, ids_bytes):
    // header layout: magic[8] version dtype dim flags (u32 x4) rows ids_bytes (u64 x2) ... head_sum at byte 120
    if (blob.size() >= 128) {
      const uint64_t forged[][2] = {{1ull << 61, 0}, {(1ull << 62) + 3, 1ull << 63}, {0xFFFFFFFFFFFFFFF0ull, 16}, {rows, ~0ull - 100}, {1ull << 33, 1ull << 45}};
      int fr = 0, fa = 0;
      for (const auto& fz : forged) {
        std::string b = blob;
        memcpy(&b[24], &fz[0], 8);
        memcpy(&b[32], &fz[1], 8);
        const uint64_t hs = fnv64_test(0xcbf29ce484222325ull, b.data(), 120);
        memcpy(&b[120], &hs, 8);
        f = fopen(cache, "wb"); fwrite(b.data(), 1, b.size(), f); fclose(f);
        rag_cache_info info;
        char* i4 = nullptr; uint64_t n4 = 0;
        const int r1 = rag_cache_info_read(cache, &info);
        const int r2 = rag_cache_read_host(cache, 0, 1, Y.data(), nullptr, nullptr, nullptr, nullptr, nullptr, &i4, &n4);
        if (r2 == RAG_OK) rag_free(i4);
        ((r1 != RAG_OK && r2 != RAG_OK) ? fr : fa)++;
      }
      printf("forged headers: rejected=%d accepted=%d\n", fr, fa);
    }
  }
  rag_free(ids);
  return 0;
}
