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
// Host-only parser over a read-only mapping of the file (no DOM: a 1M x 1536 store is ~20 GB of text). One thread
// walks the structure — node ids are parsed, the end of every embedding array is found with memchr — and hands
// slabs of row spans to a pool of threads that convert the decimal text (std::from_chars, correctly rounded: the same
// binary64 value as V8's parse) and narrow it to fp32; values that are not exactly representable in fp32 lose their
// low bits here (the reference keeps the fp64 parse — see DESIGN.md §2 "stored precision"). The number text is
// ~97% of the file, so the load scales with the host cores (one core converts ~200 MB/s of text).
//
// Append: index.insert (src/lib/memory/store.ts:56-67) makes llamaindex rewrite the whole JSON with the new node's
// embedding added at the END of embeddingDict (object order = insertion order); every byte before the old last
// embedding's ']' is unchanged. rag_parse_vector_store_json_ex can therefore RESUME at that byte offset — the
// sidecar (store_cache.cu) records it together with a hash of the prefix — and parse only the new rows.
#include "common.cuh"

#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <charconv>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

// read-only mapping of a whole file
struct mapped_file {
  const char* data = nullptr;
  size_t size = 0;
  int fd = -1;
  uint64_t st_size = 0;
  int64_t st_mtime_ns = 0;
  ~mapped_file() {
    if (data && size) munmap((void*)data, size);
    if (fd >= 0) close(fd);
  }
  int open_ro(const char* path) {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) return rag_set_error(RAG_ERR_INVALID, "cannot open %s: %s", path, strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0) return rag_set_error(RAG_ERR_INVALID, "cannot stat %s: %s", path, strerror(errno));
    // the stamp of the bytes that are actually parsed: taken from the open descriptor BEFORE reading
    st_size = (uint64_t)st.st_size;
    st_mtime_ns = (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
    size = (size_t)st.st_size;
    if (size == 0) return RAG_OK;
    void* p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (p == MAP_FAILED) { size = 0; return rag_set_error(RAG_ERR_INVALID, "cannot map %s: %s", path, strerror(errno)); }
    madvise(p, size, MADV_SEQUENTIAL);
    data = (const char*)p;
    return RAG_OK;
  }
};

// cursor over the mapping (the interface of the old FILE-based reader)
struct reader {
  const char* base;
  const char* p;
  const char* end;
  reader(const char* b, size_t n, size_t at = 0) : base(b), p(b + at), end(b + n) {}
  int peek() const { return p < end ? (unsigned char)*p : EOF; }
  int get() { return p < end ? (unsigned char)*p++ : EOF; }
  void skip_ws() {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
  }
  uint64_t where() const { return (uint64_t)(p - base); }
};

uint32_t loader_threads() {
  uint32_t n = std::thread::hardware_concurrency();
  if (const char* e = getenv("RAGERA_LOADER_THREADS")) n = (uint32_t)atoi(e);
  return std::max(1u, std::min(n, 64u));
}

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

// skip { "<id>": [numbers], ... } (the cursor sits on '{'): ids are walked as strings, arrays are jumped with memchr
bool skip_embedding_dict(reader& r) {
  if (r.get() != '{') return false;
  r.skip_ws();
  if (r.peek() == '}') { r.get(); return true; }
  for (;;) {
    r.skip_ws();
    if (!parse_string(r, nullptr)) return false;
    r.skip_ws();
    if (r.get() != ':') return false;
    r.skip_ws();
    if (r.peek() == '[') {
      const char* close = (const char*)memchr(r.p, ']', (size_t)(r.end - r.p));
      const char* nested = (const char*)memchr(r.p + 1, '[', (size_t)((close ? close : r.end) - r.p - 1));
      if (!close) return false;
      if (nested) { if (!skip_value(r)) return false; }   // not a flat array of numbers: the careful way
      else r.p = close + 1;
    } else if (!skip_value(r)) return false;
    r.skip_ws();
    const int c = r.get();
    if (c == '}') return true;
    if (c != ',') return false;
  }
}

// JSON number at [p, end) → binary64, correctly rounded (std::from_chars: same value as strtod / V8's parse, ~6x
// faster than strtod, which dominated the load time). Returns the first byte after the number, nullptr if none.
const char* parse_number(const char* p, const char* end, double* v) {
  const std::from_chars_result res = std::from_chars(p, end, *v);
  if (res.ec == std::errc::result_out_of_range) {  // 1e999 / 1e-999: strtod's answer (inf / 0 with the sign)
    char tmp[64];
    const size_t n = std::min<size_t>((size_t)(res.ptr - p), 63);
    memcpy(tmp, p, n);
    tmp[n] = 0;
    *v = strtod(tmp, nullptr);
    return res.ptr;
  }
  return res.ec == std::errc() ? res.ptr : nullptr;
}

// one embedding: the text between '[' and ']' (exclusive) → dim floats. 0 = ok, 1 = bad number, 2 = bad separator,
// 3 = wrong count (*count = what was found)
int parse_row(const char* p, const char* end, uint32_t dim, float* dst, uint32_t* count) {
  uint32_t n = 0;
  auto ws = [&]() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; };
  ws();
  if (p < end) {
    for (;;) {
      double v;
      const char* q = parse_number(p, end, &v);
      if (!q || q == p) return 1;
      if (n < dim) dst[n] = (float)v;
      n++;
      p = q;
      ws();
      if (p == end) break;
      if (*p != ',') return 2;
      p++;
      ws();
    }
  }
  *count = n;
  return n == dim ? 0 : 3;
}

struct row_span { const char* b; const char* e; };

// convert a slab of row spans with `threads` workers; returns the index of the first bad row (or n) and its error
uint64_t parse_slab(const std::vector<row_span>& spans, uint32_t dim, float* slab, uint32_t threads, int* err, uint32_t* err_count) {
  const uint64_t n = spans.size();
  std::atomic<uint64_t> next(0), first_bad(n);
  std::vector<int> errs(n, 0);
  std::vector<uint32_t> counts(n, 0);
  auto work = [&]() {
    for (;;) {
      const uint64_t i0 = next.fetch_add(16);
      if (i0 >= n) return;
      for (uint64_t i = i0; i < std::min(n, i0 + 16); i++) {
        errs[i] = parse_row(spans[i].b, spans[i].e, dim, slab + (size_t)i * dim, &counts[i]);
        if (errs[i]) {
          uint64_t cur = first_bad.load();
          while (i < cur && !first_bad.compare_exchange_weak(cur, i)) {}
        }
      }
    }
  };
  const uint32_t t = (uint32_t)std::min<uint64_t>(threads, (n + 63) / 64);
  if (t <= 1) work();
  else {
    std::vector<std::thread> pool;
    for (uint32_t i = 1; i < t; i++) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
  }
  const uint64_t bad = first_bad.load();
  if (bad < n) { *err = errs[bad]; *err_count = counts[bad]; }
  return bad;
}

}  // namespace

extern "C" {

// Parse `path` and hand out rows in slabs. Host-only (usable without a GPU): the device loader below
// and the CPU tests both go through it.
//   on_rows(user, first_row, nrows, rows_f32[nrows][dim]) is called for every slab of <= slab_rows rows
//   ids: if non-NULL, receives the node ids as a '\0'-separated blob allocated with malloc (rag_free)
//   resume_offset: 0 = parse the whole file. Otherwise the byte offset just after the ']' of the last embedding a
//     previous parse consumed (its *end_offset): parsing resumes INSIDE embeddingDict there and yields only the
//     entries that follow (first_row of the callback counts from 0 again).
//   end_offset (out): the offset just after the last embedding's ']' (== resume_offset when nothing followed; 0 when
//     embeddingDict is empty) — what a later resume needs.
//   stamp (out, optional): size and mtime of the file taken from the descriptor that was parsed, before reading.
//   on_rows == NULL: structural scan only — ids, row count and resume point at memchr speed, no number is converted.
int rag_parse_vector_store_json_ex(const char* path, uint32_t dim, uint64_t slab_rows,
                                   int (*on_rows)(void* user, uint64_t first_row, uint64_t nrows, const float* rows),
                                   void* user, uint64_t resume_offset, uint64_t* rows_out, char** ids, uint64_t* ids_bytes,
                                   uint64_t* end_offset, uint64_t* stamp_size, int64_t* stamp_mtime_ns) {
  if (!path || dim == 0 || slab_rows == 0) return rag_set_error(RAG_ERR_INVALID, "rag_parse_vector_store_json: bad argument");
  mapped_file mf;
  RAG_CHECK(mf.open_ro(path));
  if (stamp_size) *stamp_size = mf.st_size;
  if (stamp_mtime_ns) *stamp_mtime_ns = mf.st_mtime_ns;
  if (resume_offset > mf.size) return rag_set_error(RAG_ERR_INVALID, "%s: resume offset %llu is past the end of the file", path, (unsigned long long)resume_offset);
  reader r(mf.data, mf.size, (size_t)resume_offset);
  const uint32_t threads = loader_threads();
  std::string key, idblob;
  std::vector<float> slab;
  std::vector<row_span> spans;
  spans.reserve((size_t)slab_rows);
  uint64_t rows = 0, last_end = resume_offset;
  int rc = RAG_OK;
  auto fail = [&](const char* what) {
    rc = rag_set_error(RAG_ERR_INVALID, "%s: %s near byte %llu", path, what, (unsigned long long)r.where());
  };
  auto flush = [&]() {  // convert the collected spans (in parallel) and hand the slab out
    if (spans.empty() || rc != RAG_OK) return;
    if (!on_rows) { spans.clear(); return; }  // structural scan only (ids, row count, resume point): no number is converted
    if (slab.size() < spans.size() * (size_t)dim) slab.resize(spans.size() * (size_t)dim);
    int err = 0;
    uint32_t cnt = 0;
    const uint64_t bad = parse_slab(spans, dim, slab.data(), threads, &err, &cnt);
    if (bad < spans.size()) {
      const uint64_t row = rows - spans.size() + bad;
      if (err == 3) rc = rag_set_error(RAG_ERR_INVALID, "%s: embedding %llu has %u values, index dim is %u", path, (unsigned long long)row, cnt, dim);
      else rc = rag_set_error(RAG_ERR_INVALID, "%s: %s in embedding %llu near byte %llu", path, err == 1 ? "bad number" : "expected ',' in embedding",
                              (unsigned long long)row, (unsigned long long)(spans[bad].b - mf.data));
    } else {
      rc = on_rows(user, rows - spans.size(), spans.size(), slab.data());
    }
    spans.clear();
  };
  // the entries of embeddingDict from the cursor on; `first` = the cursor sits on the first entry (no leading ',')
  auto entries = [&](bool first) {
    for (;;) {
      r.skip_ws();
      if (!first) {
        const int c = r.get();
        if (c == '}') return;
        if (c != ',') { fail("expected ',' between embeddings"); return; }
        r.skip_ws();
      }
      first = false;
      if (!parse_string(r, &key)) { fail("bad node id"); return; }
      if (ids) { idblob.append(key); idblob.push_back('\0'); }
      r.skip_ws();
      if (r.get() != ':') { fail("expected ':' after node id"); return; }
      r.skip_ws();
      if (r.get() != '[') { fail("embedding is not an array"); return; }
      const char* close = (const char*)memchr(r.p, ']', (size_t)(r.end - r.p));  // numbers cannot contain ']'
      if (!close) { r.p = r.end; fail("unterminated embedding"); return; }
      spans.push_back(row_span{r.p, close});
      r.p = close + 1;
      last_end = r.where();
      rows++;
      if (spans.size() == slab_rows) { flush(); if (rc != RAG_OK) return; }
    }
  };
  do {
    if (resume_offset) {  // inside embeddingDict, right after an embedding
      entries(false);
      break;
    }
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
        if (r.peek() == '}') { r.get(); last_end = 0; }
        else entries(true);
        if (rc != RAG_OK) break;
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
  } while (0);
  flush();
  if (rc != RAG_OK) return rc;
  if (rows_out) *rows_out = rows;
  if (end_offset) *end_offset = last_end;
  if (ids) {
    char* blob = (char*)malloc(idblob.size() ? idblob.size() : 1);
    if (!blob) return rag_set_error(RAG_ERR_NOMEM, "out of host memory");
    memcpy(blob, idblob.data(), idblob.size());
    *ids = blob;
    if (ids_bytes) *ids_bytes = idblob.size();
  }
  return RAG_OK;
}

int rag_parse_vector_store_json(const char* path, uint32_t dim, uint64_t slab_rows,
                                int (*on_rows)(void* user, uint64_t first_row, uint64_t nrows, const float* rows),
                                void* user, uint64_t* rows_out, char** ids, uint64_t* ids_bytes) {
  return rag_parse_vector_store_json_ex(path, dim, slab_rows, on_rows, user, 0, rows_out, ids, ids_bytes, nullptr, nullptr, nullptr);
}

// Hash of the first `nbytes` bytes of a file (the prefix stamp of a sidecar): FNV-style over 8 MB chunks hashed in
// parallel, then over the chunk hashes. *ok = 0 when the file is shorter than nbytes.
int rag_file_prefix_hash(const char* path, uint64_t nbytes, uint64_t* hash, int* ok) {
  if (!path || !hash || !ok) return rag_set_error(RAG_ERR_INVALID, "rag_file_prefix_hash: null argument");
  *ok = 0;
  *hash = 0;
  mapped_file mf;
  RAG_CHECK(mf.open_ro(path));
  if (nbytes > mf.size) return RAG_OK;
  const uint64_t chunk = 8ull << 20;
  const uint64_t nchunks = (nbytes + chunk - 1) / chunk;
  std::vector<uint64_t> hs((size_t)nchunks);
  std::atomic<uint64_t> next(0);
  auto fnv = [](const char* p, size_t n) {   // four interleaved FNV-1a lanes over 8-byte words
    uint64_t l[4] = {0xcbf29ce484222325ull, 0x84222325cbf29ce4ull, 0x9e3779b97f4a7c15ull, 0xc2b2ae3d27d4eb4full};
    const uint64_t prime = 0x100000001b3ull;
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
      uint64_t w[4];
      memcpy(w, p + i, 32);
      for (int k = 0; k < 4; k++) l[k] = (l[k] ^ w[k]) * prime;
    }
    uint64_t h = l[0];
    for (int k = 1; k < 4; k++) h = (h ^ l[k]) * prime;
    for (; i < n; i++) h = (h ^ (unsigned char)p[i]) * prime;
    return h;
  };
  auto work = [&]() {
    for (;;) {
      const uint64_t c = next.fetch_add(1);
      if (c >= nchunks) return;
      const uint64_t o = c * chunk;
      hs[(size_t)c] = fnv(mf.data + o, (size_t)std::min(chunk, nbytes - o));
    }
  };
  const uint32_t t = (uint32_t)std::min<uint64_t>(loader_threads(), nchunks);
  std::vector<std::thread> pool;
  for (uint32_t i = 1; i < t; i++) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  *hash = fnv((const char*)hs.data(), hs.size() * 8) ^ nbytes;
  *ok = 1;
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

  mapped_file mf;
  RAG_CHECK(mf.open_ro(path));
  reader r(mf.data, mf.size);
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
      if (key == "embeddingDict" && r.peek() == '{') {
        // ~97% of the file: jump from array to array with memchr instead of walking every digit
        if (!skip_embedding_dict(r)) { fail("bad embeddingDict"); break; }
      } else if (key != "metadataDict" || r.peek() != '{') {
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
  RAG_LOCK(idx);
  upload_ctx c;
  c.idx = idx;
  c.row0 = idx->rows;
  char* blob = nullptr;
  uint64_t nbytes = 0, n = 0;
  RAG_CHECK(rag_parse_vector_store_json(path, idx->dim, 16384, upload_rows, &c, &n, &blob, &nbytes));
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
