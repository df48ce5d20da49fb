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

#include <charconv>
#include <string>
#include <unordered_map>
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
      if (c == 'u') {  // \uXXXX → UTF-8; a surrogate pair becomes one code point, a lone surrogate U+FFFD
        auto hex4 = [&](uint32_t* v) {
          *v = 0;
          for (int i = 0; i < 4; i++) {
            const int h = r.get();
            const int d = (h >= '0' && h <= '9') ? h - '0' : (h >= 'a' && h <= 'f') ? h - 'a' + 10 : (h >= 'A' && h <= 'F') ? h - 'A' + 10 : -1;
            if (d < 0) return false;
            *v = *v * 16 + (uint32_t)d;
          }
          return true;
        };
        uint32_t cp;
        if (!hex4(&cp)) return false;
        if (cp >= 0xD800 && cp <= 0xDBFF) {
          uint32_t lo = 0;
          if (r.peek() == '\\') {
            r.get();
            if (r.get() != 'u' || !hex4(&lo)) return false;
            if (lo >= 0xDC00 && lo <= 0xDFFF) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            else { if (out) out->append("\xEF\xBF\xBD"); cp = lo; if (cp >= 0xD800 && cp <= 0xDFFF) cp = 0xFFFD; }
          } else cp = 0xFFFD;
        } else if (cp >= 0xDC00 && cp <= 0xDFFF) cp = 0xFFFD;
        if (cp == 0) cp = 0xFFFD;  // ids travel as '\0'-separated blobs
        if (out) {
          if (cp < 0x80) out->push_back((char)cp);
          else if (cp < 0x800) { out->push_back((char)(0xC0 | (cp >> 6))); out->push_back((char)(0x80 | (cp & 0x3F))); }
          else if (cp < 0x10000) { out->push_back((char)(0xE0 | (cp >> 12))); out->push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out->push_back((char)(0x80 | (cp & 0x3F))); }
          else { out->push_back((char)(0xF0 | (cp >> 18))); out->push_back((char)(0x80 | ((cp >> 12) & 0x3F))); out->push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out->push_back((char)(0x80 | (cp & 0x3F))); }
        }
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

// JSON number → binary64, correctly rounded (std::from_chars: same value as strtod / V8's parse, ~6x faster than
// strtod, which dominated the load time)
bool parse_number(reader& r, double* v) {
  char tmp[64];
  int n = 0;
  for (int c = r.peek(); n < 63 && (c == '-' || c == '+' || c == '.' || c == 'e' || c == 'E' || (c >= '0' && c <= '9')); c = r.peek()) {
    tmp[n++] = (char)c;
    r.pos++;
  }
  if (n == 0) return false;
  const std::from_chars_result res = std::from_chars(tmp, tmp + n, *v);
  if (res.ec == std::errc::result_out_of_range) {  // 1e999 / 1e-999: strtod's answer (inf / 0 with the sign)
    tmp[n] = 0;
    *v = strtod(tmp, nullptr);
    return true;
  }
  return res.ec == std::errc() && res.ptr == tmp + n;
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

// Second pass over the same file: "metadataDict" { "<nodeId>": { ...flat node metadata... } } (written by
// SimpleVectorStore.add next to embeddingDict; upstream-recalled). Per row it yields what the hot path reads:
//   content_type[row]  — the rule of vectorSearch (src/lib/hybrid-search.ts:229-234) without the per-call
//                        isCodebase flag: metadata.type === 'memory' → memory, else metadata.language !== undefined
//                        (the key is present, whatever its value) → code, else document. Rows without an entry are documents.
//   memory_ids         — metadata.memoryId (src/lib/memory/store.ts:58-61), one '\0'-terminated string per row
//                        (empty when absent): the host joins it with the Prisma Memory rows to fill
//                        confidence / accessCount / lastAccessedAt (rag_index_set_row_meta).
// ids / ids_bytes: the blob rag_parse_vector_store_json returned for this file (row order).
// Returns RAG_OK also when the file has no metadataDict (*found = 0, every row a document).
int rag_parse_vector_store_metadata(const char* path, const char* ids, uint64_t ids_bytes, uint64_t rows,
                                    uint8_t* content_type, char** memory_ids, uint64_t* memory_ids_bytes, int* found) {
  if (!path || (rows && (!ids || !content_type)))
    return rag_set_error(RAG_ERR_INVALID, "rag_parse_vector_store_metadata: bad argument");
  std::vector<const char*> id_of(rows);
  {
    const char* p = ids;
    const char* end = ids + ids_bytes;
    for (uint64_t r = 0; r < rows; r++) {
      if (p >= end) return rag_set_error(RAG_ERR_INVALID, "ids blob holds fewer than %llu ids", (unsigned long long)rows);
      id_of[r] = p;
      p += strlen(p) + 1;
    }
  }
  memset(content_type, RAG_CT_DOCUMENT, rows);
  std::vector<std::string> mem(memory_ids ? rows : 0);
  std::unordered_map<std::string, uint64_t> by_id;  // built only if the file's order differs from the row order
  uint64_t expect = 0;
  auto row_of = [&](const std::string& id, uint64_t* row) {
    if (expect < rows && id == id_of[expect]) { *row = expect++; return true; }
    if (by_id.empty())
      for (uint64_t r = 0; r < rows; r++) by_id.emplace(id_of[r], r);
    auto it = by_id.find(id);
    if (it == by_id.end()) return false;
    *row = it->second;
    expect = it->second + 1;
    return true;
  };

  FILE* f = fopen(path, "rb");
  if (!f) return rag_set_error(RAG_ERR_INVALID, "cannot open %s: %s", path, strerror(errno));
  reader r(f);
  std::string key, id, sval;
  int rc = RAG_OK;
  bool seen = false;
  auto fail = [&](const char* what) {
    rc = rag_set_error(RAG_ERR_INVALID, "%s: %s near byte %llu", path, what, (unsigned long long)r.where());
  };
  do {
    r.skip_ws();
    if (r.get() != '{') { fail("expected a JSON object"); break; }
    for (;;) {
      r.skip_ws();
      if (r.peek() == '}') { r.get(); break; }
      if (!parse_string(r, &key)) { fail("bad key"); break; }
      r.skip_ws();
      if (r.get() != ':') { fail("expected ':'"); break; }
      r.skip_ws();
      if (key != "metadataDict" || r.peek() != '{') {
        if (!skip_value(r)) { fail("bad value"); break; }
      } else {
        seen = true;
        r.get();
        r.skip_ws();
        if (r.peek() == '}') r.get();
        else {
          for (;;) {  // one node
            r.skip_ws();
            if (!parse_string(r, &id)) { fail("bad node id"); break; }
            uint64_t row = 0;
            const bool known = row_of(id, &row);
            r.skip_ws();
            if (r.get() != ':') { fail("expected ':' after node id"); break; }
            r.skip_ws();
            if (r.peek() != '{') {  // null or a scalar: no metadata
              if (!skip_value(r)) { fail("bad metadata value"); break; }
            } else {
              r.get();
              bool is_memory = false, has_language = false;
              r.skip_ws();
              if (r.peek() == '}') r.get();
              else {
                for (;;) {
                  r.skip_ws();
                  if (!parse_string(r, &key)) { fail("bad metadata key"); break; }
                  r.skip_ws();
                  if (r.get() != ':') { fail("expected ':' in metadata"); break; }
                  r.skip_ws();
                  if (key == "type" && r.peek() == '"') {
                    if (!parse_string(r, &sval)) { fail("bad string"); break; }
                    is_memory = sval == "memory";
                  } else if (key == "language") {  // `!== undefined`: present with any value, null and "" included
                    has_language = true;
                    if (!skip_value(r)) { fail("bad value"); break; }
                  } else if (key == "memoryId" && r.peek() == '"') {
                    if (!parse_string(r, &sval)) { fail("bad string"); break; }
                    if (known && memory_ids) mem[row] = sval;
                  } else if (!skip_value(r)) { fail("bad value"); break; }
                  r.skip_ws();
                  const int c = r.get();
                  if (c == '}') break;
                  if (c != ',') { fail("expected ',' in metadata"); break; }
                }
                if (rc != RAG_OK) break;
              }
              if (known) content_type[row] = is_memory ? RAG_CT_MEMORY : has_language ? RAG_CT_CODE : RAG_CT_DOCUMENT;
            }
            r.skip_ws();
            const int c = r.get();
            if (c == '}') break;
            if (c != ',') { fail("expected ',' between nodes"); break; }
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
  } while (0);
  fclose(f);
  if (rc != RAG_OK) return rc;
  if (found) *found = seen ? 1 : 0;
  if (memory_ids) {
    size_t total = 0;
    for (const std::string& m : mem) total += m.size() + 1;
    char* blob = (char*)malloc(total ? total : 1);
    if (!blob) return rag_set_error(RAG_ERR_NOMEM, "out of host memory");
    char* w = blob;
    for (const std::string& m : mem) { memcpy(w, m.c_str(), m.size() + 1); w += m.size() + 1; }
    *memory_ids = blob;
    if (memory_ids_bytes) *memory_ids_bytes = total;
  }
  return RAG_OK;
}

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
  char* blob = nullptr;
  uint64_t nbytes = 0, n = 0;
  RAG_CHECK(rag_parse_vector_store_json(path, idx->dim, 4096, upload_rows, &c, &n, &blob, &nbytes));
  // metadata.type / metadata.language of the same file decide contentType (hybrid-search.ts:229-234)
  int rc = RAG_OK, found = 0;
  if (n) {
    std::vector<uint8_t> ct(n);
    rc = rag_parse_vector_store_metadata(path, blob, nbytes, n, ct.data(), nullptr, nullptr, &found);
    if (rc == RAG_OK && found) rc = rag_index_set_row_meta(idx, c.row0, n, ct.data(), nullptr, nullptr, nullptr);
  }
  if (rc != RAG_OK || !ids) free(blob);
  if (rc != RAG_OK) return rc;
  if (rows_loaded) *rows_loaded = n;
  if (ids) { *ids = blob; if (ids_bytes) *ids_bytes = nbytes; }
  return RAG_OK;
}
