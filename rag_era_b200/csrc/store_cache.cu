// store_cache.cu — §8f N1: binary sidecar of a persisted index.
//
// The reference reloads ./storage/kb_<id>/vector_store.json (decimal text, ~20 bytes per value) on every cold
// start (src/lib/llm/index-manager.ts:246-275). Parsing that text is the slow part of bringing a shard up, so
// the first load writes the parsed rows next to it in the index dtype, and later loads stream that file straight
// into HBM. The sidecar is only a cache: it records size + mtime of the JSON it was made from — taken from the
// descriptor that was parsed, before reading — and is not trusted when they no longer match.
//
// index.insert appends to the JSON (src/lib/memory/store.ts:56-67: memories arrive one at a time), which would make
// the sidecar stale after every insert. It therefore also records WHERE its last embedding ended in the JSON and a
// hash of the bytes up to there: when the JSON has changed but still starts with those bytes, only the embeddings
// that follow are parsed, and the sidecar is extended in place (new rows after the old ones, the small tail
// sections rewritten). Anything else — a rebuilt index, a deleted node — falls back to the full parse.
//
// File layout (little endian, every section 16-byte aligned):
//   [0,128)   rag_cache_header (magic, shape, section sizes, source stamp, checksums; written LAST, so an
//             interrupted writer leaves a file without a valid magic; the file is renamed into place when done)
//   rows      rows x dim elements of the index dtype, row-major, unpadded
//   blocks    one 64-bit FNV-style checksum per block of `block_rows` rows — a shard can load and verify only its own row range
//   meta      (flag META) content_type u8[rows] · confidence f64[rows] · access_count i32[rows] · last_access_ms i64[rows]
//   keys      (flag KEYS) fusion key u64[rows]
//   ids       node ids, '\0'-separated, row order
// tail_sum covers blocks..ids; head_sum covers the header itself.
//
// Everything except rag_index_save_cache / rag_index_load_cache is host-only and usable without a GPU.
#include "common.cuh"

#include <errno.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <memory>
#include <new>
#include <string>
#include <vector>

// (loader.cu) the parser with a resume point and the stamp of the descriptor it read; the prefix hash
extern "C" int rag_parse_vector_store_json_ex(const char* path, uint32_t dim, uint64_t slab_rows,
                                              int (*on_rows)(void* user, uint64_t first_row, uint64_t nrows, const float* rows),
                                              void* user, uint64_t resume_offset, uint64_t* rows_out, char** ids,
                                              uint64_t* ids_bytes, uint64_t* end_offset, uint64_t* stamp_size,
                                              int64_t* stamp_mtime_ns);
extern "C" int rag_file_prefix_hash(const char* path, uint64_t nbytes, uint64_t* hash, int* ok);

namespace {

constexpr char kMagic[8] = {'R', 'A', 'G', 'E', 'R', 'A', 'C', '1'};
constexpr uint32_t kVersion = 1;
constexpr uint32_t kBlockRows = 4096;
constexpr uint64_t kFnvBasis = 0xcbf29ce484222325ull, kFnvPrime = 0x100000001b3ull;

struct rag_cache_header {
  char magic[8];
  uint32_t version, dtype, dim, flags;
  uint64_t rows, ids_bytes;
  uint64_t source_size;
  int64_t source_mtime_ns;
  uint32_t block_rows, reserved0;
  uint64_t tail_sum;
  uint64_t src_prefix_bytes;  // byte offset in the JSON just after the last embedding this sidecar holds (0 = unknown)
  uint64_t src_prefix_hash;   // rag_file_prefix_hash of those bytes: an appended-to JSON still starts with them
  uint64_t reserved[4];
  uint64_t head_sum;  // FNV-1a-64 of the 120 bytes before it
};
static_assert(sizeof(rag_cache_header) == 128, "cache header layout");

// FNV-1a-style checksum over 8-byte words in four interleaved lanes (word i goes to lane i % 4; the lanes are folded
// into the running value at the end of every call, the tail byte by byte): four independent multiply chains keep up
// with the page cache, one chain does not. Call boundaries are part of the definition — writer and reader both feed
// whole sections / whole row blocks.
uint64_t fnv64(uint64_t h, const void* data, size_t n) {
  const unsigned char* p = (const unsigned char*)data;
  uint64_t l0 = h, l1 = h ^ 0x9e3779b97f4a7c15ull, l2 = h ^ 0xc2b2ae3d27d4eb4full, l3 = h ^ 0x165667b19e3779f9ull;
  for (; n >= 32; n -= 32, p += 32) {
    uint64_t w[4];
    memcpy(w, p, 32);
    l0 = (l0 ^ w[0]) * kFnvPrime;
    l1 = (l1 ^ w[1]) * kFnvPrime;
    l2 = (l2 ^ w[2]) * kFnvPrime;
    l3 = (l3 ^ w[3]) * kFnvPrime;
  }
  h = l0;
  h = (h ^ l1) * kFnvPrime;
  h = (h ^ l2) * kFnvPrime;
  h = (h ^ l3) * kFnvPrime;
  for (; n >= 8; n -= 8, p += 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    h = (h ^ w) * kFnvPrime;
  }
  for (; n; n--, p++) h = (h ^ *p) * kFnvPrime;
  return h;
}

size_t pad16(size_t n) { return (n + 15) & ~(size_t)15; }
size_t esize(uint32_t dtype) { return dtype == RAG_BF16 ? 2 : 4; }

struct sections {
  uint64_t rows_off, blocks_off, nblocks, ct_off, conf_off, acc_off, last_off, keys_off, ids_off, end;
};
// false when a size wraps or exceeds anything a real file could hold (a corrupt or crafted header)
bool layout(const rag_cache_header& h, sections* out) {
  sections s;
  if (h.rows >= 0xFFFFFFFFull || h.dim == 0 || h.dim > 8192 || h.block_rows == 0 || h.ids_bytes > (1ull << 40)) return false;
  uint64_t o = sizeof(rag_cache_header);
  bool ok = true;
  auto take = [&](uint64_t count, uint64_t each) {
    uint64_t bytes = 0;
    if (__builtin_mul_overflow(count, each, &bytes) || bytes > (1ull << 46)) ok = false;
    const uint64_t at = o;
    if (ok) o += pad16((size_t)bytes);
    return at;
  };
  s.rows_off = take(h.rows, (uint64_t)h.dim * esize(h.dtype));
  s.nblocks = (h.rows + h.block_rows - 1) / h.block_rows;
  s.blocks_off = take(s.nblocks, 8);
  s.ct_off = s.conf_off = s.acc_off = s.last_off = s.keys_off = 0;
  if (h.flags & RAG_CACHE_META) {
    s.ct_off = take(h.rows, 1);
    s.conf_off = take(h.rows, 8);
    s.acc_off = take(h.rows, 4);
    s.last_off = take(h.rows, 8);
  }
  if (h.flags & RAG_CACHE_KEYS) s.keys_off = take(h.rows, 8);
  s.ids_off = take(h.ids_bytes, 1);
  s.end = o;
  *out = s;
  return ok;
}

// where a sidecar's rows came from: size + mtime of the JSON (from the parsed descriptor) and the resume point
struct src_stamp {
  uint64_t size = 0;
  int64_t mtime_ns = 0;
  uint64_t prefix_bytes = 0, prefix_hash = 0;
};

bool stat_source(const char* path, uint64_t* size, int64_t* mtime_ns) {
  struct stat st;
  if (!path || stat(path, &st) != 0) return false;
  *size = (uint64_t)st.st_size;
  *mtime_ns = (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
  return true;
}

// ---- streaming writer: begin → append_rows* → finish ------------------------------------------------------------
struct cache_writer {
  FILE* f = nullptr;
  std::string path, tmp;
  rag_cache_header h;
  uint64_t written = 0;        // rows so far
  uint64_t in_block = 0;
  std::vector<char> pending;   // rows of the block being filled
  std::vector<uint64_t> blocks;

  ~cache_writer() {
    if (f) { fclose(f); remove(tmp.c_str()); }
  }
  int begin(const char* p, uint32_t dtype, uint32_t dim, uint64_t rows) {
    path = p;
    tmp = path + ".tmp";
    f = fopen(tmp.c_str(), "wb");
    if (!f) return rag_set_error(RAG_ERR_INVALID, "cannot create %s: %s", tmp.c_str(), strerror(errno));
    memset(&h, 0, sizeof(h));
    h.version = kVersion;
    h.dtype = dtype;
    h.dim = dim;
    h.rows = rows;
    h.block_rows = kBlockRows;
    char zero[sizeof(rag_cache_header)] = {0};  // the real header goes in last
    return put(zero, sizeof(zero));
  }
  int put(const void* p, size_t n) {
    if (n && fwrite(p, 1, n, f) != n) return rag_set_error(RAG_ERR_INVALID, "write to %s failed: %s", tmp.c_str(), strerror(errno));
    return RAG_OK;
  }
  int put_padded(const void* p, size_t n, uint64_t* sum) {
    static const char zeros[16] = {0};
    RAG_CHECK(put(p, n));
    if (sum) *sum = fnv64(*sum, p, n);
    return put(zeros, pad16(n) - n);
  }
  int append_rows(const void* rows, uint64_t n) {
    if (h.rows != ~0ull && written + n > h.rows) return rag_set_error(RAG_ERR_INVALID, "cache writer: more rows than announced");
    const size_t rb = (size_t)h.dim * esize(h.dtype);
    const char* p = (const char*)rows;
    RAG_CHECK(put(p, (size_t)n * rb));
    while (n) {  // a block is always hashed in ONE call (the reader does the same): partial blocks wait in `pending`
      const uint64_t take = std::min<uint64_t>(n, h.block_rows - in_block);
      if (in_block == 0 && take == h.block_rows) {
        blocks.push_back(fnv64(kFnvBasis, p, (size_t)take * rb));
      } else {
        pending.insert(pending.end(), p, p + (size_t)take * rb);
        if ((in_block += take) == h.block_rows) {
          blocks.push_back(fnv64(kFnvBasis, pending.data(), pending.size()));
          pending.clear();
          in_block = 0;
        }
      }
      p += (size_t)take * rb;
      n -= take;
      written += take;
    }
    return RAG_OK;
  }
  int finish(const uint8_t* ct, const double* conf, const int32_t* acc, const int64_t* last, const uint64_t* keys,
             const char* ids, uint64_t ids_bytes, const src_stamp* stamp) {
    if (h.rows == ~0ull) h.rows = written;  // streamed: the count is known only now
    if (written != h.rows) return rag_set_error(RAG_ERR_INVALID, "cache writer: %llu of %llu rows written",
                                                (unsigned long long)written, (unsigned long long)h.rows);
    if (in_block) blocks.push_back(fnv64(kFnvBasis, pending.data(), pending.size()));
    static const char zeros[16] = {0};
    const size_t rows_bytes = (size_t)h.rows * h.dim * esize(h.dtype);
    RAG_CHECK(put(zeros, pad16(rows_bytes) - rows_bytes));
    const bool meta = ct && conf && acc && last;
    h.flags = (meta ? RAG_CACHE_META : 0u) | (keys ? RAG_CACHE_KEYS : 0u);
    h.ids_bytes = ids ? ids_bytes : 0;
    uint64_t sum = kFnvBasis;
    RAG_CHECK(put_padded(blocks.data(), blocks.size() * 8, &sum));
    if (meta) {
      RAG_CHECK(put_padded(ct, h.rows, &sum));
      RAG_CHECK(put_padded(conf, h.rows * 8, &sum));
      RAG_CHECK(put_padded(acc, h.rows * 4, &sum));
      RAG_CHECK(put_padded(last, h.rows * 8, &sum));
    }
    if (keys) RAG_CHECK(put_padded(keys, h.rows * 8, &sum));
    RAG_CHECK(put_padded(ids, h.ids_bytes, &sum));
    h.tail_sum = sum;
    if (stamp) {
      h.source_size = stamp->size;
      h.source_mtime_ns = stamp->mtime_ns;
      h.src_prefix_bytes = stamp->prefix_bytes;
      h.src_prefix_hash = stamp->prefix_hash;
    }
    memcpy(h.magic, kMagic, 8);
    h.head_sum = fnv64(kFnvBasis, &h, offsetof(rag_cache_header, head_sum));
    if (fseek(f, 0, SEEK_SET) != 0) return rag_set_error(RAG_ERR_INVALID, "seek in %s failed", tmp.c_str());
    RAG_CHECK(put(&h, sizeof(h)));
    if (fflush(f) != 0 || fclose(f) != 0) { f = nullptr; remove(tmp.c_str()); return rag_set_error(RAG_ERR_INVALID, "closing %s failed: %s", tmp.c_str(), strerror(errno)); }
    f = nullptr;
    if (rename(tmp.c_str(), path.c_str()) != 0) {
      remove(tmp.c_str());
      return rag_set_error(RAG_ERR_INVALID, "rename %s -> %s failed: %s", tmp.c_str(), path.c_str(), strerror(errno));
    }
    return RAG_OK;
  }
};

// ---- reader ------------------------------------------------------------------------------------------------------
struct cache_reader {
  FILE* f = nullptr;
  std::string path;
  rag_cache_header h;
  sections s;
  std::vector<uint64_t> blocks;
  std::vector<char> tail;  // blocks..ids, verified against tail_sum

  ~cache_reader() { if (f) fclose(f); }
  int bad(const char* what) { return rag_set_error(RAG_ERR_INVALID, "%s: %s", path.c_str(), what); }
  int read_at(uint64_t off, void* dst, size_t n) {
    if (n == 0) return RAG_OK;
    if (fseeko(f, (off_t)off, SEEK_SET) != 0 || fread(dst, 1, n, f) != n) return bad("short read (truncated cache)");
    return RAG_OK;
  }
  int open(const char* p, bool want_tail) {
    path = p ? p : "";
    if (!p) return rag_set_error(RAG_ERR_INVALID, "null cache path");
    f = fopen(p, "rb");
    if (!f) return rag_set_error(RAG_ERR_INVALID, "cannot open %s: %s", p, strerror(errno));
    if (fread(&h, 1, sizeof(h), f) != sizeof(h)) return bad("not a ragera cache (too short)");
    if (memcmp(h.magic, kMagic, 8) != 0) return bad("not a ragera cache (bad magic)");
    if (h.head_sum != fnv64(kFnvBasis, &h, offsetof(rag_cache_header, head_sum))) return bad("header checksum mismatch");
    if (h.version != kVersion) return bad("unsupported cache version");
    if ((h.dtype != RAG_F32 && h.dtype != RAG_BF16) || h.dim == 0 || h.dim > 8192 || h.block_rows == 0) return bad("bad header fields");
    if (!layout(h, &s)) return bad("bad header fields (section sizes out of range)");
    struct stat st;
    if (fstat(fileno(f), &st) != 0 || (uint64_t)st.st_size != s.end) return bad("file size does not match the header (truncated cache)");
    if (!want_tail) return RAG_OK;
    tail.resize((size_t)(s.end - s.blocks_off));
    RAG_CHECK(read_at(s.blocks_off, tail.data(), tail.size()));
    // the checksum runs over the unpadded sections, in file order
    uint64_t sum = kFnvBasis;
    auto sec = [&](uint64_t off, uint64_t n) { sum = fnv64(sum, tail.data() + (off - s.blocks_off), (size_t)n); };
    sec(s.blocks_off, s.nblocks * 8);
    if (h.flags & RAG_CACHE_META) { sec(s.ct_off, h.rows); sec(s.conf_off, h.rows * 8); sec(s.acc_off, h.rows * 4); sec(s.last_off, h.rows * 8); }
    if (h.flags & RAG_CACHE_KEYS) sec(s.keys_off, h.rows * 8);
    sec(s.ids_off, h.ids_bytes);
    if (sum != h.tail_sum) return bad("metadata checksum mismatch");
    blocks.resize((size_t)s.nblocks);
    if (!blocks.empty()) memcpy(blocks.data(), tail.data(), blocks.size() * 8);
    return RAG_OK;
  }
  const char* at(uint64_t off) const { return tail.data() + (off - s.blocks_off); }
  // rows [first, first+n) through on_rows in slabs of whole blocks; every touched block is verified
  template <typename F>
  int rows(uint64_t first, uint64_t n, void* slab, uint64_t slab_rows, F&& on_rows) {
    if (first > h.rows || n > h.rows - first) return bad("row range exceeds the cache");
    if (n == 0) return RAG_OK;
    const size_t rb = (size_t)h.dim * esize(h.dtype);
    const uint64_t b0 = first / h.block_rows, b1 = (first + n - 1) / h.block_rows;
    const uint64_t per_slab = std::max<uint64_t>(1, slab_rows / h.block_rows);
    for (uint64_t b = b0; b <= b1; b += per_slab) {
      const uint64_t be = std::min(b1 + 1, b + per_slab);
      const uint64_t r0 = b * h.block_rows, r1 = std::min<uint64_t>(h.rows, be * h.block_rows);
      RAG_CHECK(read_at(s.rows_off + r0 * rb, slab, (size_t)(r1 - r0) * rb));
      for (uint64_t k = b; k < be; k++) {
        const uint64_t k0 = k * h.block_rows, k1 = std::min<uint64_t>(h.rows, k0 + h.block_rows);
        if (fnv64(kFnvBasis, (const char*)slab + (size_t)(k0 - r0) * rb, (size_t)(k1 - k0) * rb) != blocks[(size_t)k])
          return rag_set_error(RAG_ERR_INVALID, "%s: rows %llu..%llu are corrupt (block checksum mismatch)", path.c_str(),
                               (unsigned long long)k0, (unsigned long long)k1);
      }
      const uint64_t u0 = std::max(first, r0), u1 = std::min(first + n, r1);
      RAG_CHECK(on_rows(u0, u1 - u0, (const char*)slab + (size_t)(u0 - r0) * rb));
    }
    return RAG_OK;
  }
};

int copy_ids(const cache_reader& r, char** ids, uint64_t* ids_bytes) {
  if (!ids) return RAG_OK;
  char* blob = (char*)malloc(r.h.ids_bytes ? (size_t)r.h.ids_bytes : 1);
  if (!blob) return rag_set_error(RAG_ERR_NOMEM, "out of host memory");
  if (r.h.ids_bytes) memcpy(blob, r.at(r.s.ids_off), (size_t)r.h.ids_bytes);
  *ids = blob;
  if (ids_bytes) *ids_bytes = r.h.ids_bytes;
  return RAG_OK;
}

void fill_info(const rag_cache_header& h, rag_cache_info* out) {
  out->version = h.version;
  out->dtype = h.dtype;
  out->dim = h.dim;
  out->flags = h.flags;
  out->rows = h.rows;
  out->ids_bytes = h.ids_bytes;
  out->source_size = h.source_size;
  out->source_mtime_ns = h.source_mtime_ns;
  out->source_prefix_bytes = h.src_prefix_bytes;
  out->source_prefix_hash = h.src_prefix_hash;
}

// stamp from a path (public entry points that take `source_json`): size + mtime as they are NOW — only as good as the
// caller's promise that the JSON has not changed since the rows were read from it — and, when the JSON holds exactly
// `rows` embeddings, the resume point behind the last one (a structural scan at memchr speed), so that a sidecar saved
// by the host (with its Memory columns and fusion keys) can later be extended in place like one written by the loader
bool stamp_of_path(const char* source_json, uint64_t rows, src_stamp* st, bool* have) {
  *have = false;
  if (!source_json) return true;
  uint64_t n = 0, end = 0;
  if (rag_parse_vector_store_json_ex(source_json, 1, 65536, nullptr, nullptr, 0, &n, nullptr, nullptr, &end, &st->size, &st->mtime_ns) != RAG_OK) {
    if (!stat_source(source_json, &st->size, &st->mtime_ns)) return false;  // not a store we can scan: size + mtime only
    *have = true;
    return true;
  }
  int ok = 0;
  if (n == rows && end > 0 && rag_file_prefix_hash(source_json, end, &st->prefix_hash, &ok) == RAG_OK && ok) st->prefix_bytes = end;
  *have = true;
  return true;
}

int save_cache_stamped(rag_index* idx, const char* cache_path, const char* ids, uint64_t ids_bytes, const src_stamp* stamp);

// Extend a sidecar in place with `n_new` rows parsed from the tail of the JSON: new rows go after the old ones (over
// the old tail sections, which `r` holds in memory), then the rebuilt tail, then the header — which is invalidated
// first, so a crash in between leaves a file that is simply not a sidecar (the next start parses the JSON in full).
int cache_extend(cache_reader& r, const void* new_rows, uint64_t n_new, const uint8_t* new_ct, const char* new_ids,
                 uint64_t new_ids_bytes, const src_stamp& stamp) {
  const rag_cache_header old = r.h;
  const sections os = r.s;
  const size_t rb = (size_t)old.dim * esize(old.dtype);
  rag_cache_header h = old;
  h.rows = old.rows + n_new;
  h.ids_bytes = old.ids_bytes + new_ids_bytes;
  bool any_special = false;
  for (uint64_t i = 0; i < n_new && new_ct; i++) any_special |= new_ct[i] != RAG_CT_DOCUMENT;
  if (any_special) h.flags |= RAG_CACHE_META;
  sections ns;
  if (!layout(h, &ns)) return r.bad("extended cache would be too large");
  // block checksums: the whole old blocks stay; the last partial block is re-hashed together with the rows that fill it
  std::vector<uint64_t> blocks(r.blocks);
  const uint64_t partial = old.rows % old.block_rows;
  std::vector<char> pend;
  if (partial) {
    blocks.pop_back();
    pend.resize((size_t)partial * rb);
    RAG_CHECK(r.read_at(os.rows_off + (old.rows - partial) * rb, pend.data(), pend.size()));
  }
  const char* p = (const char*)new_rows;
  uint64_t left = n_new, in_block = partial;
  while (left) {
    const uint64_t take = std::min<uint64_t>(left, old.block_rows - in_block);
    pend.insert(pend.end(), p, p + (size_t)take * rb);
    p += (size_t)take * rb;
    left -= take;
    if ((in_block += take) == old.block_rows) {
      blocks.push_back(fnv64(kFnvBasis, pend.data(), pend.size()));
      pend.clear();
      in_block = 0;
    }
  }
  if (in_block) blocks.push_back(fnv64(kFnvBasis, pend.data(), pend.size()));
  // the tail sections, old values followed by the new rows' (Memory columns and fusion keys of new rows are the host's to set)
  std::vector<uint8_t> ct;
  std::vector<double> conf;
  std::vector<int32_t> acc;
  std::vector<int64_t> last;
  std::vector<uint64_t> keys;
  if (h.flags & RAG_CACHE_META) {
    ct.assign((size_t)h.rows, RAG_CT_DOCUMENT); conf.assign((size_t)h.rows, 0.0); acc.assign((size_t)h.rows, 0); last.assign((size_t)h.rows, 0);
    if (old.flags & RAG_CACHE_META) {
      memcpy(ct.data(), r.at(os.ct_off), (size_t)old.rows);
      memcpy(conf.data(), r.at(os.conf_off), (size_t)old.rows * 8);
      memcpy(acc.data(), r.at(os.acc_off), (size_t)old.rows * 4);
      memcpy(last.data(), r.at(os.last_off), (size_t)old.rows * 8);
    }
    if (new_ct) memcpy(ct.data() + old.rows, new_ct, (size_t)n_new);
  }
  if (h.flags & RAG_CACHE_KEYS) {
    keys.resize((size_t)h.rows);
    memcpy(keys.data(), r.at(os.keys_off), (size_t)old.rows * 8);
    for (uint64_t i = old.rows; i < h.rows; i++) keys[(size_t)i] = i;  // default key = chunk id, until the host sets one
  }
  std::string ids(r.at(os.ids_off), (size_t)old.ids_bytes);
  ids.append(new_ids, (size_t)new_ids_bytes);

  FILE* f = fopen(r.path.c_str(), "r+b");
  if (!f) return rag_set_error(RAG_ERR_INVALID, "cannot reopen %s for writing: %s", r.path.c_str(), strerror(errno));
  int rc = RAG_OK;
  auto put_at = [&](uint64_t off, const void* src, size_t n) {
    if (rc != RAG_OK || n == 0) return;
    if (fseeko(f, (off_t)off, SEEK_SET) != 0 || fwrite(src, 1, n, f) != n)
      rc = rag_set_error(RAG_ERR_INVALID, "write to %s failed: %s", r.path.c_str(), strerror(errno));
  };
  static const char zeros[sizeof(rag_cache_header)] = {0};
  put_at(0, zeros, sizeof(zeros));  // not a sidecar while it is being rewritten
  if (rc == RAG_OK && fflush(f) != 0) rc = rag_set_error(RAG_ERR_INVALID, "flush of %s failed", r.path.c_str());
  put_at(os.rows_off + old.rows * rb, new_rows, (size_t)n_new * rb);
  uint64_t sum = kFnvBasis;
  auto section = [&](uint64_t off, const void* src, size_t n) {
    put_at(off, src, n);
    sum = fnv64(sum, src, n);
    const size_t pad = pad16(n) - n;
    put_at(off + n, zeros, pad);
  };
  {  // padding after the rows section
    const size_t rows_bytes = (size_t)h.rows * rb;
    put_at(ns.rows_off + rows_bytes, zeros, pad16(rows_bytes) - rows_bytes);
  }
  section(ns.blocks_off, blocks.data(), blocks.size() * 8);
  if (h.flags & RAG_CACHE_META) {
    section(ns.ct_off, ct.data(), ct.size());
    section(ns.conf_off, conf.data(), conf.size() * 8);
    section(ns.acc_off, acc.data(), acc.size() * 4);
    section(ns.last_off, last.data(), last.size() * 8);
  }
  if (h.flags & RAG_CACHE_KEYS) section(ns.keys_off, keys.data(), keys.size() * 8);
  section(ns.ids_off, ids.data(), ids.size());
  h.tail_sum = sum;
  h.source_size = stamp.size;
  h.source_mtime_ns = stamp.mtime_ns;
  h.src_prefix_bytes = stamp.prefix_bytes;
  h.src_prefix_hash = stamp.prefix_hash;
  h.head_sum = fnv64(kFnvBasis, &h, offsetof(rag_cache_header, head_sum));
  if (rc == RAG_OK && (fflush(f) != 0 || ftruncate(fileno(f), (off_t)ns.end) != 0))
    rc = rag_set_error(RAG_ERR_INVALID, "resizing %s failed: %s", r.path.c_str(), strerror(errno));
  put_at(0, &h, sizeof(h));  // valid again
  if (fclose(f) != 0 && rc == RAG_OK) rc = rag_set_error(RAG_ERR_INVALID, "closing %s failed: %s", r.path.c_str(), strerror(errno));
  if (rc != RAG_OK) remove(r.path.c_str());
  return rc;
}

}  // namespace

extern "C" {

// header of a sidecar (validated: magic, header checksum, version, file size)
int rag_cache_info_read(const char* cache_path, rag_cache_info* out) {
  if (!out) return rag_set_error(RAG_ERR_INVALID, "rag_cache_info_read: null out");
  cache_reader r;
  RAG_CHECK(r.open(cache_path, false));
  fill_info(r.h, out);
  return RAG_OK;
}

// 1 = the sidecar exists, is well formed and was made from `source_json` as it is now (same size and mtime);
// 0 = missing, unreadable or stale — re-parse the JSON. Never an error.
int rag_cache_is_fresh(const char* cache_path, const char* source_json) {
  cache_reader r;
  uint64_t size = 0;
  int64_t mtime = 0;
  if (r.open(cache_path, false) != RAG_OK) return 0;
  if (!stat_source(source_json, &size, &mtime)) return 0;
  return (r.h.source_size == size && r.h.source_mtime_ns == mtime) ? 1 : 0;
}

// write a sidecar from host arrays (rows in `dtype`, [rows][dim]); meta needs all four arrays or none
int rag_cache_write_host(const char* cache_path, uint32_t dtype, uint32_t dim, uint64_t rows, const void* host_rows,
                         const uint8_t* content_type, const double* confidence, const int32_t* access_count,
                         const int64_t* last_access_ms, const uint64_t* keys, const char* ids, uint64_t ids_bytes,
                         const char* source_json) {
  if (!cache_path || (dtype != RAG_F32 && dtype != RAG_BF16) || dim == 0 || dim > 8192 || (rows && !host_rows))
    return rag_set_error(RAG_ERR_INVALID, "rag_cache_write_host: bad argument");
  const int nmeta = (content_type != nullptr) + (confidence != nullptr) + (access_count != nullptr) + (last_access_ms != nullptr);
  if (nmeta != 0 && nmeta != 4) return rag_set_error(RAG_ERR_INVALID, "rag_cache_write_host: give all four metadata arrays or none");
  src_stamp st;
  bool have = false;
  if (!stamp_of_path(source_json, rows, &st, &have)) return rag_set_error(RAG_ERR_INVALID, "cannot stat %s: %s", source_json, strerror(errno));
  cache_writer w;
  RAG_CHECK(w.begin(cache_path, dtype, dim, rows));
  RAG_CHECK(w.append_rows(host_rows, rows));
  return w.finish(content_type, confidence, access_count, last_access_ms, keys, ids, ids_bytes, have ? &st : nullptr);
}

// read a sidecar into host arrays (any pointer may be NULL; sizes from rag_cache_info_read). Verifies every checksum.
int rag_cache_read_host(const char* cache_path, uint64_t first_row, uint64_t nrows, void* host_rows, uint8_t* content_type,
                        double* confidence, int32_t* access_count, int64_t* last_access_ms, uint64_t* keys, char** ids,
                        uint64_t* ids_bytes) {
  cache_reader r;
  RAG_CHECK(r.open(cache_path, true));
  if (first_row > r.h.rows || nrows > r.h.rows - first_row) return r.bad("row range exceeds the cache");
  const size_t rb = (size_t)r.h.dim * esize(r.h.dtype);
  if (host_rows && nrows) {
    const uint64_t slab_rows = std::min<uint64_t>((uint64_t)r.h.block_rows * 4, ((nrows + 2 * r.h.block_rows - 1) / r.h.block_rows) * r.h.block_rows);
    std::unique_ptr<char[]> slab(new (std::nothrow) char[(size_t)slab_rows * rb]);  // uninitialised on purpose
    if (!slab) return rag_set_error(RAG_ERR_NOMEM, "out of host memory");
    RAG_CHECK(r.rows(first_row, nrows, slab.get(), slab_rows, [&](uint64_t r0, uint64_t n, const char* p) {
      memcpy((char*)host_rows + (size_t)(r0 - first_row) * rb, p, (size_t)n * rb);
      return (int)RAG_OK;
    }));
  }
  if ((content_type || confidence || access_count || last_access_ms) && !(r.h.flags & RAG_CACHE_META))
    return r.bad("the cache holds no row metadata");
  if (keys && !(r.h.flags & RAG_CACHE_KEYS)) return r.bad("the cache holds no fusion keys");
  if (content_type) memcpy(content_type, r.at(r.s.ct_off) + first_row, (size_t)nrows);
  if (confidence) memcpy(confidence, r.at(r.s.conf_off) + first_row * 8, (size_t)nrows * 8);
  if (access_count) memcpy(access_count, r.at(r.s.acc_off) + first_row * 4, (size_t)nrows * 4);
  if (last_access_ms) memcpy(last_access_ms, r.at(r.s.last_off) + first_row * 8, (size_t)nrows * 8);
  if (keys) memcpy(keys, r.at(r.s.keys_off) + first_row * 8, (size_t)nrows * 8);
  return copy_ids(r, ids, ids_bytes);
}

// ---- device side -------------------------------------------------------------------------------------------------

// write rows [0, rows) of this handle (its shard), their metadata / fusion keys if set, and the caller's node ids.
// source_json (may be NULL): the JSON these rows were read from — it is stat'ed NOW, so only pass it when the file has
// not changed since it was read (rag_index_open_store stamps from the descriptor it parsed instead).
int rag_index_save_cache(rag_index* idx, const char* cache_path, const char* ids, uint64_t ids_bytes, const char* source_json) {
  if (!idx || !cache_path) return rag_set_error(RAG_ERR_INVALID, "rag_index_save_cache: null argument");
  RAG_LOCK(idx);
  src_stamp st;
  bool have = false;
  if (!stamp_of_path(source_json, idx->rows, &st, &have)) return rag_set_error(RAG_ERR_INVALID, "cannot stat %s: %s", source_json, strerror(errno));
  return save_cache_stamped(idx, cache_path, ids, ids_bytes, have ? &st : nullptr);
}
}  // extern "C"

namespace {
int save_cache_stamped(rag_index* idx, const char* cache_path, const char* ids, uint64_t ids_bytes, const src_stamp* stamp) {
  RAG_CUDA(cudaSetDevice(idx->device));
  const uint64_t rows = idx->rows;
  const size_t rb = (size_t)idx->dim * (idx->desc.dtype == RAG_BF16 ? 2 : 4);
  cache_writer w;
  RAG_CHECK(w.begin(cache_path, idx->desc.dtype, idx->dim, rows));
  const uint64_t slab_rows = 4 * kBlockRows;
  void* slab = nullptr;
  RAG_CUDA(cudaHostAlloc(&slab, (size_t)slab_rows * rb, cudaHostAllocDefault));
  int rc = RAG_OK;
  for (uint64_t r = 0; r < rows && rc == RAG_OK; r += slab_rows) {
    const uint64_t n = std::min(slab_rows, rows - r);
    if ((rc = rag_index_read_rows(idx, r, n, slab)) == RAG_OK) rc = w.append_rows(slab, n);
  }
  cudaFreeHost(slab);
  RAG_CHECK(rc);
  std::vector<uint8_t> ct;
  std::vector<double> conf;
  std::vector<int32_t> acc;
  std::vector<int64_t> last;
  std::vector<uint64_t> keys;
  if (idx->ctype && rows) {
    ct.resize(rows); conf.resize(rows); acc.resize(rows); last.resize(rows);
    RAG_CUDA(cudaMemcpyAsync(ct.data(), idx->ctype, rows, cudaMemcpyDeviceToHost, idx->stream));
    RAG_CUDA(cudaMemcpyAsync(conf.data(), idx->conf, rows * 8, cudaMemcpyDeviceToHost, idx->stream));
    RAG_CUDA(cudaMemcpyAsync(acc.data(), idx->access, rows * 4, cudaMemcpyDeviceToHost, idx->stream));
    RAG_CUDA(cudaMemcpyAsync(last.data(), idx->last_ms, rows * 8, cudaMemcpyDeviceToHost, idx->stream));
  }
  if (idx->row_keys && rows) {
    keys.resize(rows);
    RAG_CUDA(cudaMemcpyAsync(keys.data(), idx->row_keys, rows * 8, cudaMemcpyDeviceToHost, idx->stream));
  }
  RAG_CUDA(cudaStreamSynchronize(idx->stream));
  const bool meta = !ct.empty();
  return w.finish(meta ? ct.data() : nullptr, meta ? conf.data() : nullptr, meta ? acc.data() : nullptr,
                  meta ? last.data() : nullptr, keys.empty() ? nullptr : keys.data(), ids, ids_bytes, stamp);
}
}  // namespace

extern "C" {

// append rows [first_row, first_row + nrows) of the sidecar (nrows = 0: to its end) after the handle's rows, with
// their metadata and keys. A shard passes its own range (rag_era_b200.sharded.shard_range); ids come back whole.
int rag_index_load_cache(rag_index* idx, const char* cache_path, uint64_t first_row, uint64_t nrows, uint64_t* rows_loaded,
                         char** ids, uint64_t* ids_bytes) {
  if (!idx) return rag_set_error(RAG_ERR_INVALID, "null index handle");
  RAG_LOCK(idx);
  cache_reader r;
  RAG_CHECK(r.open(cache_path, true));
  if (r.h.dtype != idx->desc.dtype || r.h.dim != idx->dim)
    return rag_set_error(RAG_ERR_INVALID, "%s holds %s rows of dim %u; the index is %s, dim %u", cache_path,
                         r.h.dtype == RAG_BF16 ? "bf16" : "f32", r.h.dim, idx->desc.dtype == RAG_BF16 ? "bf16" : "f32", idx->dim);
  if (first_row > r.h.rows) return r.bad("first_row is past the end of the cache");
  if (nrows == 0) nrows = r.h.rows - first_row;
  if (nrows > r.h.rows - first_row) return r.bad("row range exceeds the cache");
  const uint64_t row0 = idx->rows;
  if (row0 + nrows > idx->desc.capacity_rows)
    return rag_set_error(RAG_ERR_INVALID, "rag_index_load_cache: %llu rows exceed capacity %llu",
                         (unsigned long long)(row0 + nrows), (unsigned long long)idx->desc.capacity_rows);
  RAG_CUDA(cudaSetDevice(idx->device));
  const size_t rb = (size_t)r.h.dim * esize(r.h.dtype);
  const uint64_t slab_rows = 4 * kBlockRows;
  void* slab = nullptr;
  RAG_CUDA(cudaHostAlloc(&slab, (size_t)slab_rows * rb, cudaHostAllocDefault));
  const int rc = r.rows(first_row, nrows, slab, slab_rows, [&](uint64_t r0, uint64_t n, const char* p) {
    return rag_index_upload(idx, row0 + (r0 - first_row), n, p);  // synchronous: the slab is free again on return
  });
  cudaFreeHost(slab);
  RAG_CHECK(rc);
  if ((r.h.flags & RAG_CACHE_META) && nrows)
    RAG_CHECK(rag_index_set_row_meta(idx, row0, nrows, (const uint8_t*)r.at(r.s.ct_off) + first_row,
                                     (const double*)r.at(r.s.conf_off) + first_row, (const int32_t*)r.at(r.s.acc_off) + first_row,
                                     (const int64_t*)r.at(r.s.last_off) + first_row));
  if ((r.h.flags & RAG_CACHE_KEYS) && nrows)
    RAG_CHECK(rag_index_set_row_keys(idx, row0, nrows, (const uint64_t*)r.at(r.s.keys_off) + first_row));
  if (rows_loaded) *rows_loaded = nrows;
  return copy_ids(r, ids, ids_bytes);
}

// Bring the sidecar of a vector_store.json up to date WITHOUT a GPU (also what a deployment runs ahead of a cold start).
// *route reports what it took:
//   1  nothing: the sidecar is fresh (the JSON has the size and mtime the sidecar was made from)
//   2  the JSON has CHANGED but still starts with the bytes the sidecar was made from (index.insert appended —
//      src/lib/memory/store.ts:56-67): only the appended embeddings are parsed, the sidecar is extended in place
//   0  anything else (no sidecar, another dtype / dim, a rebuilt or edited JSON): full parse on all host threads, rewrite
// The stamp is taken from the descriptor that is parsed, before reading; if the JSON changes while it is being read,
// the call fails with RAG_ERR_STATE and leaves no sidecar that would claim to be fresh.
int rag_cache_refresh_host(const char* cache_path, const char* source_json, uint32_t dtype, uint32_t dim, int* route,
                           uint64_t* rows_out) {
  if (!source_json || (dtype != RAG_F32 && dtype != RAG_BF16) || dim == 0 || dim > 8192)
    return rag_set_error(RAG_ERR_INVALID, "rag_cache_refresh_host: bad argument");
  const std::string cp = cache_path ? std::string(cache_path) : std::string(source_json) + ".ragera";
  if (route) *route = 0;
  rag_cache_info info;
  const bool usable = rag_cache_info_read(cp.c_str(), &info) == RAG_OK && info.dtype == dtype && info.dim == dim;
  if (usable && rag_cache_is_fresh(cp.c_str(), source_json)) {
    if (route) *route = 1;
    if (rows_out) *rows_out = info.rows;
    return RAG_OK;
  }
  struct sink {
    uint32_t dtype, dim;
    cache_writer* w;
    std::vector<char>* rows;   // extension: collected in memory (appends are small)
    std::vector<uint16_t> tmp;
    static int put(void* user, uint64_t, uint64_t n, const float* rows_f32) {
      sink* s = (sink*)user;
      const void* src = rows_f32;
      size_t bytes = (size_t)n * s->dim * 4;
      if (s->dtype == RAG_BF16) {
        s->tmp.resize((size_t)n * s->dim);
        for (size_t i = 0; i < s->tmp.size(); i++) s->tmp[i] = rg_f32_to_bf16(rows_f32[i]);
        src = s->tmp.data();
        bytes /= 2;
      }
      if (s->w) return s->w->append_rows(src, n);
      s->rows->insert(s->rows->end(), (const char*)src, (const char*)src + bytes);
      return (int)RAG_OK;
    }
  };
  auto unchanged = [&](const src_stamp& st) {
    uint64_t size = 0;
    int64_t mtime = 0;
    return stat_source(source_json, &size, &mtime) && size == st.size && mtime == st.mtime_ns;
  };
  if (usable) {
    // appended-to JSON? the sidecar knows where its last embedding ended and what the bytes before that hash to
    cache_reader r;
    uint64_t hash = 0;
    int ok = 0;
    if (r.open(cp.c_str(), true) == RAG_OK && r.h.src_prefix_bytes > 0 &&
        rag_file_prefix_hash(source_json, r.h.src_prefix_bytes, &hash, &ok) == RAG_OK && ok && hash == r.h.src_prefix_hash) {
      std::vector<char> new_rows;
      sink sk{dtype, dim, nullptr, &new_rows, {}};
      src_stamp st;
      char* new_ids = nullptr;
      uint64_t n_new = 0, new_bytes = 0;
      int rc = rag_parse_vector_store_json_ex(source_json, dim, 16384, sink::put, &sk, r.h.src_prefix_bytes, &n_new, &new_ids, &new_bytes,
                                              &st.prefix_bytes, &st.size, &st.mtime_ns);
      if (rc == RAG_OK) {
        std::vector<uint8_t> ct;
        if (n_new) {  // contentType of the new rows: the metadata pass matches entries by id, so it gets all ids
          std::string all(r.at(r.s.ids_off), (size_t)r.h.ids_bytes);
          all.append(new_ids, (size_t)new_bytes);
          std::vector<uint8_t> all_ct((size_t)(r.h.rows + n_new));
          int found = 0;
          rc = rag_parse_vector_store_metadata(source_json, all.data(), all.size(), r.h.rows + n_new, all_ct.data(), nullptr, nullptr, &found);
          if (rc == RAG_OK && found) ct.assign(all_ct.begin() + (ptrdiff_t)r.h.rows, all_ct.end());
        }
        int hok = 0;
        if (rc == RAG_OK) rc = rag_file_prefix_hash(source_json, st.prefix_bytes, &st.prefix_hash, &hok);
        if (rc == RAG_OK && (!hok || !unchanged(st))) rc = rag_set_error(RAG_ERR_STATE, "%s changed while it was being read", source_json);
        const uint64_t n_old = r.h.rows;
        if (rc == RAG_OK) rc = cache_extend(r, new_rows.data(), n_new, ct.empty() ? nullptr : ct.data(), new_ids, new_bytes, st);
        if (rc == RAG_OK) {
          free(new_ids);
          if (route) *route = 2;
          if (rows_out) *rows_out = n_old + n_new;
          return RAG_OK;
        }
      }
      free(new_ids);
      // a malformed tail, a JSON that moved under us, a write error: fall through to the full rebuild
    }
  }
  cache_writer w;
  RAG_CHECK(w.begin(cp.c_str(), dtype, dim, ~0ull));
  sink sk{dtype, dim, &w, nullptr, {}};
  src_stamp st;
  char* ids = nullptr;
  uint64_t n = 0, nbytes = 0;
  RAG_CHECK(rag_parse_vector_store_json_ex(source_json, dim, 16384, sink::put, &sk, 0, &n, &ids, &nbytes, &st.prefix_bytes, &st.size, &st.mtime_ns));
  std::vector<uint8_t> ct((size_t)n);
  std::vector<double> conf;
  std::vector<int32_t> acc;
  std::vector<int64_t> last;
  int found = 0, rc = RAG_OK, hok = 0;
  if (n) rc = rag_parse_vector_store_metadata(source_json, ids, nbytes, n, ct.data(), nullptr, nullptr, &found);
  bool special = false;
  for (uint64_t i = 0; i < n && found; i++) special |= ct[(size_t)i] != RAG_CT_DOCUMENT;
  if (special) { conf.assign((size_t)n, 0.0); acc.assign((size_t)n, 0); last.assign((size_t)n, 0); }
  if (rc == RAG_OK && st.prefix_bytes && (rag_file_prefix_hash(source_json, st.prefix_bytes, &st.prefix_hash, &hok) != RAG_OK || !hok)) st.prefix_bytes = 0;
  if (rc == RAG_OK && !unchanged(st)) rc = rag_set_error(RAG_ERR_STATE, "%s changed while it was being read", source_json);
  if (rc == RAG_OK)
    rc = w.finish(special ? ct.data() : nullptr, special ? conf.data() : nullptr, special ? acc.data() : nullptr,
                  special ? last.data() : nullptr, nullptr, ids, nbytes, &st);
  free(ids);
  if (rc == RAG_OK && rows_out) *rows_out = n;
  return rc;
}

// loadIndex (src/lib/llm/index-manager.ts:246-275) for an EMPTY handle: bring the sidecar up to date
// (rag_cache_refresh_host: nothing to do / extend with the appended embeddings / full parse) and stream it into HBM.
// *from_cache = the refresh route (1 fresh, 2 extended in place, 0 the JSON was parsed in full). cache_path NULL →
// "<vector_store_json>.ragera". When the sidecar cannot be written or read (read-only storage, a corrupt file) the JSON
// is parsed straight into the index instead: a missing cache is never an error.
int rag_index_open_store(rag_index* idx, const char* vector_store_json, const char* cache_path, uint64_t* rows_loaded,
                         char** ids, uint64_t* ids_bytes, int* from_cache) {
  if (!idx || !vector_store_json) return rag_set_error(RAG_ERR_INVALID, "rag_index_open_store: null argument");
  RAG_LOCK(idx);
  if (idx->rows != 0) return rag_set_error(RAG_ERR_STATE, "rag_index_open_store needs an empty index (rows=%llu)", (unsigned long long)idx->rows);
  const std::string cp = cache_path ? std::string(cache_path) : std::string(vector_store_json) + ".ragera";
  if (from_cache) *from_cache = 0;
  int route = 0;
  if (rag_cache_refresh_host(cp.c_str(), vector_store_json, idx->desc.dtype, idx->dim, &route, nullptr) == RAG_OK) {
    const int rc = rag_index_load_cache(idx, cp.c_str(), 0, 0, rows_loaded, ids, ids_bytes);
    if (rc == RAG_OK) { if (from_cache) *from_cache = route; return RAG_OK; }
    if (idx->rows != 0) return rc;  // corrupt half-way: the handle already holds rows, the caller must start over
  }
  return rag_index_load_vector_store(idx, vector_store_json, rows_loaded, ids, ids_bytes);
}

}  // extern "C"
