// loader.cu — §8f N1: read the reference's on-disk vector store into a rag_index.
//
// The reference persists its index with llamaindex `storageContextFromDefaults({persistDir})`
// (src/lib/llm/index-manager.ts:218-220,264-270) under ./storage/kb_<id>/. The dense rows live in
// `vector_store.json`, written by SimpleVectorStore.persist as
//     { "embeddingDict": { "<nodeId>": [n, n, ...], ... }, "textIdToRefDocId": {...}, "metadataDict": {...} }
// (upstream-recalled layout of llamaindex@0.12.1; only "embeddingDict" is read here). JSON object
// order is the Map/insertion order the reference scans in, so row r of the device matrix is the
// r-th entry and "lower row wins ties" stays the reference's stable sort.
//
// Host-only streaming parser (no DOM: a 1M x 1536 store is ~20 GB of text). Numbers are parsed with
// strtod and narrowed to the index dtype; values that are not exactly representable in fp32 lose
// their low bits here (the reference keeps the fp64 parse — see DESIGN.md §2 "stored precision").
#include "common.cuh"

#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

namespace {

struct reader {
  FILE* f = nullptr;
  std::vector<char> buf;
  size_t pos = 0, len = 0;
  uint64_t offset = 0;
  explicit reader(FILE* fp) : f(fp), buf(1 << 22) {}
  int peek() {
    if (pos == len) {
      len = fread(buf.data(), 1, buf.size(), f);
      offset += pos;
      pos = 0;
      if (len == 0) return EOF;
    }
    return (unsigned char)buf[pos];
  }
  int get() {
    const int c = peek();
    if (c != EOF) pos++;
    return c;
  }
  void skip_ws() {
    for (int c = peek(); c == ' ' || c == '\n' || c == '\t' || c == '\r'; c = peek()) pos++;
  }
  uint64_t where() const { return offset + pos; }
};

bool parse_string(reader& r, std::string* out) {
  if (r.get() != '"') return false;
  if (out) out->clear();
  for (;;) {
    int c = r.get();
    if (c == EOF) return false;
    if (c == '"') return true;
    if (c == '\\') {
      c = r.get();
      if (c == EOF) return false;
      if (c == 'u') {  // keep \uXXXX escapes verbatim: node ids are UUIDs, this is only for robustness
        if (out) out->append("\\u");
        continue;
      }
      const char* map = "\"\"\\\\//b\bf\fn\nr\rt\t";
      char v = (char)c;
      for (const char* m = map; *m; m += 2)
        if (*m == c) v = m[1];
      if (out) out->push_back(v);
    } else if (out) {
      out->push_back((char)c);
    }
  }
}

// skip any JSON value
bool skip_value(reader& r) {
  r.skip_ws();
  const int c = r.peek();
  if (c == '"') return parse_string(r, nullptr);
  if (c == '{' || c == '[') {
    const int close = c == '{' ? '}' : ']';
    r.get();
    r.skip_ws();
    if (r.peek() == close) { r.get(); return true; }
    for (;;) {
      if (c == '{') {
        r.skip_ws();
        if (!parse_string(r, nullptr)) return false;
        r.skip_ws();
        if (r.get() != ':') return false;
      }
      if (!skip_value(r)) return false;
      r.skip_ws();
      const int d = r.get();
      if (d == close) return true;
      if (d != ',') return false;
    }
  }
  // number / true / false / null
  for (int d = r.peek(); d != EOF && d != ',' && d != '}' && d != ']' && d != ' ' && d != '\n' && d != '\r' && d != '\t'; d = r.peek())
    r.get();
  return true;
}

bool parse_number(reader& r, double* v) {
  char tmp[64];
  int n = 0;
  for (int c = r.peek(); n < 63 && (c == '-' || c == '+' || c == '.' || c == 'e' || c == 'E' || (c >= '0' && c <= '9')); c = r.peek()) {
    tmp[n++] = (char)c;
    r.get();
  }
  if (n == 0) return false;
  tmp[n] = 0;
  char* end = nullptr;
  *v = strtod(tmp, &end);
  return end == tmp + n;
}

}  // namespace

extern "C" {

// Parse `path` and hand out rows in slabs. Host-only (usable without a GPU): the device loader below
// and the CPU tests both go through it.
//   on_rows(user, first_row, nrows, rows_f32[nrows][dim]) is called for every slab of <= slab_rows rows
//   ids: if non-NULL, receives the node ids as a '\0'-separated blob allocated with malloc (rag_free)
int rag_parse_vector_store_json(const char* path, uint32_t dim, uint64_t slab_rows,
                                int (*on_rows)(void* user, uint64_t first_row, uint64_t nrows, const float* rows),
                                void* user, uint64_t* rows_out, char** ids, uint64_t* ids_bytes) {
  if (!path || dim == 0 || slab_rows == 0) return rag_set_error(RAG_ERR_INVALID, "rag_parse_vector_store_json: bad argument");
  FILE* f = fopen(path, "rb");
  if (!f) return rag_set_error(RAG_ERR_INVALID, "cannot open %s: %s", path, strerror(errno));
  reader r(f);
  std::string key, idblob;
  std::vector<float> slab((size_t)slab_rows * dim);
  uint64_t rows = 0, in_slab = 0;
  int rc = RAG_OK;
  auto fail = [&](const char* what) {
    rc = rag_set_error(RAG_ERR_INVALID, "%s: %s near byte %llu", path, what, (unsigned long long)r.where());
  };
  do {
    r.skip_ws();
    if (r.get() != '{') { fail("expected a JSON object"); break; }
    bool found = false;
    for (;;) {
      r.skip_ws();
      if (r.peek() == '}') { r.get(); break; }
      if (!parse_string(r, &key)) { fail("bad key"); break; }
      r.skip_ws();
      if (r.get() != ':') { fail("expected ':'"); break; }
      r.skip_ws();
      if (key != "embeddingDict") {
        if (!skip_value(r)) { fail("bad value"); break; }
      } else {
        found = true;
        if (r.get() != '{') { fail("embeddingDict is not an object"); break; }
        r.skip_ws();
        if (r.peek() == '}') r.get();
        else {
          for (;;) {
            r.skip_ws();
            if (!parse_string(r, &key)) { fail("bad node id"); break; }
            if (ids) { idblob.append(key); idblob.push_back('\0'); }
            r.skip_ws();
            if (r.get() != ':') { fail("expected ':' after node id"); break; }
            r.skip_ws();
            if (r.get() != '[') { fail("embedding is not an array"); break; }
            float* dst = slab.data() + (size_t)in_slab * dim;
            uint32_t n = 0;
            r.skip_ws();
            if (r.peek() == ']') r.get();
            else {
              for (;;) {
                r.skip_ws();
                double v;
                if (!parse_number(r, &v)) { fail("bad number"); break; }
                if (n < dim) dst[n] = (float)v;
                n++;
                r.skip_ws();
                const int c = r.get();
                if (c == ']') break;
                if (c != ',') { fail("expected ',' in embedding"); break; }
              }
              if (rc != RAG_OK) break;
            }
            if (n != dim) {
              rc = rag_set_error(RAG_ERR_INVALID, "%s: embedding %llu has %u values, index dim is %u", path,
                                 (unsigned long long)rows, n, dim);
              break;
            }
            rows++;
            if (++in_slab == slab_rows) {
              if (on_rows && (rc = on_rows(user, rows - in_slab, in_slab, slab.data())) != RAG_OK) break;
              in_slab = 0;
            }
            r.skip_ws();
            const int c = r.get();
            if (c == '}') break;
            if (c != ',') { fail("expected ',' between embeddings"); break; }
          }
          if (rc != RAG_OK) break;
        }
      }
      r.skip_ws();
      const int c = r.peek();
      if (c == ',') { r.get(); continue; }
      if (c == '}') { r.get(); break; }
      fail("expected ',' or '}'");
      break;
    }
    if (rc != RAG_OK) break;
    if (!found) { rc = rag_set_error(RAG_ERR_INVALID, "%s has no \"embeddingDict\"", path); break; }
    if (in_slab && on_rows) rc = on_rows(user, rows - in_slab, in_slab, slab.data());
  } while (0);
  fclose(f);
  if (rc != RAG_OK) return rc;
  if (rows_out) *rows_out = rows;
  if (ids) {
    char* blob = (char*)malloc(idblob.size() ? idblob.size() : 1);
    if (!blob) return rag_set_error(RAG_ERR_NOMEM, "out of host memory");
    memcpy(blob, idblob.data(), idblob.size());
    *ids = blob;
    if (ids_bytes) *ids_bytes = idblob.size();
  }
  return RAG_OK;
}

void rag_free(void* p) { free(p); }

}  // extern "C"

namespace {
struct upload_ctx {
  rag_index* idx;
  uint64_t row0;
  std::vector<uint16_t> bf16;
};
int upload_rows(void* user, uint64_t first, uint64_t n, const float* rows) {
  upload_ctx* c = (upload_ctx*)user;
  if (c->idx->desc.dtype == RAG_F32) return rag_index_upload(c->idx, c->row0 + first, n, rows);
  c->bf16.resize((size_t)n * c->idx->dim);
  for (size_t i = 0; i < c->bf16.size(); i++) c->bf16[i] = rg_f32_to_bf16(rows[i]);
  return rag_index_upload(c->idx, c->row0 + first, n, c->bf16.data());
}
}  // namespace

// Append every embedding of a llamaindex `vector_store.json` to the index (rows keep the file's order).
extern "C" int rag_index_load_vector_store(rag_index* idx, const char* path, uint64_t* rows_loaded, char** ids,
                                           uint64_t* ids_bytes) {
  if (!idx) return rag_set_error(RAG_ERR_INVALID, "null index handle");
  upload_ctx c;
  c.idx = idx;
  c.row0 = idx->rows;
  return rag_parse_vector_store_json(path, idx->dim, 4096, upload_rows, &c, rows_loaded, ids, ids_bytes);
}
