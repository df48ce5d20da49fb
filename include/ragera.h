/*
 * ragera.h — C ABI of libragera.so, the B200-native (sm_100a) retrieval hot path
 * of gong9/rag-era: dense cosine scoring → per-query top-k → min-cosine filter →
 * Reciprocal Rank Fusion (+ memory freshness), behind hybridSearch().
 *
 * This is the drop-in boundary. The reference is TypeScript; the calls below are
 * what an N-API addon / FFI shim binds (INTEGRATION.md shows the stub). Each
 * entry point cites the reference interface it replaces (paths relative to the
 * reference repository root).
 *
 * Conventions
 *  - plain pointers and sizes only; no C++ / torch types
 *  - return 0 (RAG_OK) or a negative rag_status; rag_last_error() returns a
 *    thread-local message for the last failing call on this thread
 *  - the caller owns every host buffer (outputs are caller-allocated); the
 *    library owns all device memory behind the opaque rag_index handle
 *  - calls on ONE handle are serialised by the library (a per-handle mutex held for the whole call): several threads
 *    may share a handle — concurrent requests of a web server, the micro-batcher's worker next to direct calls — and
 *    simply take turns; distinct handles run concurrently. The staged form (stage → run → fetch) is a multi-call
 *    sequence and belongs to one owner at a time
 *  - calls are synchronous w.r.t. the host unless named *_async / *_staged
 *  - there is NO CPU fallback: without a usable sm_100 device calls fail
 *  - ids are uint64 chunk ids = id_base + local row (row order = insertion order
 *    of the reference's embeddingDict, so "lower id wins ties" == its stable sort)
 */
#ifndef RAGERA_H
#define RAGERA_H

#include <stdint.h>
#include "ragera_gen.h"

#ifdef __cplusplus
extern "C" {
#endif

#define RAGERA_VERSION 0x00010000 /* major<<16 | minor<<8 | patch */

#define RAG_MAX_TOPK 64        /* similarityTopK; reference uses 2..29 */
#define RAG_MAX_CANDIDATES 128 /* K' = k + slack candidates rescored exactly     */
#define RAG_MAX_KEYWORDS 64    /* keywordLimit; reference uses 0..19             */
#define RAG_MAX_FRESH 64

typedef struct rag_index rag_index;

typedef enum rag_status {
  RAG_OK = 0,
  RAG_ERR_INVALID = -1,     /* bad argument                                      */
  RAG_ERR_CUDA = -2,        /* CUDA runtime / driver error (message has detail)   */
  RAG_ERR_NOMEM = -3,
  RAG_ERR_STATE = -4,       /* call order (e.g. search before any rows exist)     */
  RAG_ERR_UNSUPPORTED = -5, /* e.g. tensor path requested without a bf16 shadow   */
  RAG_ERR_NCCL = -6,
  RAG_ERR_NO_DEVICE = -7,   /* no sm_100 GPU: there is deliberately no fallback   */
  RAG_ERR_TIMEOUT = -8,     /* sharded search: a peer rank never arrived at the exchange */
  RAG_ERR_BUSY = -9         /* rag_batcher_submit_async: every batch buffer is in flight; nothing was queued */
} rag_status;

typedef enum rag_dtype { RAG_F32 = 0, RAG_BF16 = 1 } rag_dtype;

/* HybridSearchResult.source — src/lib/hybrid-search.ts:24 (+ the N-c4 extension) */
typedef enum rag_source { RAG_SRC_VECTOR = 0, RAG_SRC_KEYWORD = 1, RAG_SRC_BOTH = 2, RAG_SRC_FRESHNESS = 3 } rag_source;
/* HybridSearchResult.contentType — src/lib/hybrid-search.ts:25,229-234 */
typedef enum rag_content_type { RAG_CT_DOCUMENT = 0, RAG_CT_MEMORY = 1, RAG_CT_CODE = 2 } rag_content_type;

/* which scoring kernel produces the candidates (results are identical) */
typedef enum rag_path {
  RAG_PATH_AUTO = 0,  /* one query on an fp32 index with an fp16 shadow → shadow stream; B < gemm threshold →
                         stream; else tensor (if an operand exists)                 */
  RAG_PATH_STREAM = 1,/* K1: HBM-bound fused cosine GEMV + top-k, fp32 accumulate   */
  RAG_PATH_TENSOR = 2,/* K2: tcgen05 bf16 GEMM + fused top-k epilogue               */
  RAG_PATH_EXACT = 3, /* K1x: fp64 reference-order scan of every row (slow, exact)   */
  RAG_PATH_SHADOW_STREAM = 4 /* K1 over the fp16 shadow of the normalised rows (RAG_INDEX_F16_SHADOW): the same HBM-bound
                         scan at 2 bytes per element; certified by the rows' measured rounding residual, escalates to
                         RAG_PATH_STREAM                                            */
} rag_path;

/* 16-bit operand of the tensor path for an fp32 corpus (selection only: reported scores are always recomputed from the
 * fp32 rows in fp64). F16: fp16 of the NORMALISED row — 11 significant bits, so the rigorous certification bound is
 * ~8x tighter than with bf16 and nearly every query certifies in the first pass (preferred). BF16: bf16 of the row. */
#define RAG_INDEX_BF16_SHADOW 1u
#define RAG_INDEX_F16_SHADOW 2u

typedef struct rag_index_desc {
  uint64_t capacity_rows; /* rows this handle (shard) can hold                      */
  uint32_t dim;           /* embedding dimension D                                  */
  uint32_t dtype;         /* rag_dtype of the stored corpus                         */
  int32_t  device;        /* CUDA device ordinal                                    */
  uint32_t flags;         /* RAG_INDEX_*                                            */
  uint64_t id_base;       /* chunk id of local row 0 (row-sharded multi-GPU)        */
} rag_index_desc;

/* RRFConfig — src/lib/hybrid-search.ts:40-45; presets :77-105 */
typedef struct rag_rrf_config {
  double k;
  double vector_weight;
  double keyword_weight;
  double both_bonus;
} rag_rrf_config;

typedef struct rag_search_opts {
  uint32_t k;       /* similarityTopK (src/lib/hybrid-search.ts:223)                */
  uint32_t path;    /* rag_path                                                     */
  uint32_t slack;   /* K' = k + slack candidates; 0 → library default               */
  uint32_t flags;   /* RAG_SEARCH_*                                                 */
  double   epsilon; /* selection-error bound used for certification; 0 → default    */
} rag_search_opts;

#define RAG_SEARCH_NO_ESCALATE 1u /* do not re-run uncertified queries on a stronger path */
/* Certification bound of the tensor path (K2). Default: RIGOROUS — a deterministic bound on
 * |selected score - exact cosine| from the measured rounding residuals of the operands
 * (||q - round(q)|| / ||q|| per query, max over rows of ||x - round(x)|| / ||x||, Cauchy-Schwarz), plus the
 * fp32 accumulation and normalisation terms: certified = 1 is then a proof that the ids equal an exact scan.
 * RAG_SEARCH_STAT_EPS selects the round-1 statistical bound instead (0.024/sqrt(ld) for bf16 operands,
 * 0.006/sqrt(ld) for tf32: ~11 sigma on near-Gaussian rows — tighter, certifies more queries in the first
 * pass, but NOT a proof: rows whose rounding errors align can exceed it). */
#define RAG_SEARCH_STAT_EPS 2u

/* result of SimpleVectorStore.query → {ids, similarities} (llamaindex, via
 * src/lib/hybrid-search.ts:223-224); rank-ordered, scores are exact fp64 cosines */
typedef struct rag_topk_out {
  uint64_t* ids;       /* [B][k]                                                    */
  double*   scores;    /* [B][k]                                                    */
  uint32_t* counts;    /* [B]   results per query (<= k)                            */
  uint8_t*  certified; /* [B]   optional (may be NULL): 1 = proven equal to an exact scan */
} rag_topk_out;

/* HybridSearchOptions + the engine's call (src/lib/hybrid-search.ts:51-58,275-295) */
typedef struct rag_hybrid_opts {
  uint32_t vector_top_k;     /* vectorTopK                                          */
  uint32_t keyword_limit;    /* row stride of kw_keys (>= every kw_counts[b])       */
  double   min_vector_score; /* minVectorScore, applied to raw cosine before fusion */
  rag_rrf_config rrf;
  uint32_t path;             /* rag_path                                            */
  uint32_t slack;
  uint32_t flags;
  /* north-star extension (SURVEY N-c4 ii); fresh_limit = 0 → reference-faithful    */
  uint32_t fresh_limit;      /* length of the freshness-ranked third list           */
  double   fresh_weight;
  int64_t  now_ms;
  double   time_decay_factor; /* freshness.ts:20-23 defaults 0.05 / 0.1 when 0      */
  double   frequency_bonus;
  double   epsilon;
} rag_hybrid_opts;

/* HybridSearchResult[] per query (src/lib/hybrid-search.ts:18-27) as columns */
typedef struct rag_fused_out {
  uint32_t  capacity;     /* entries per query in the arrays below;
                             must be >= vector_top_k + keyword_limit + fresh_limit */
  uint64_t* keys;         /* [B][capacity] fusion key (RRF branch) or chunk id (vector-only branch) */
  double*   scores;       /* [B][capacity] RRF score, or raw cosine in the vector-only branch (:346-354) */
  uint8_t*  source;       /* [B][capacity] rag_source                               */
  uint8_t*  content_type; /* [B][capacity] rag_content_type                         */
  uint32_t* counts;       /* [B]                                                    */
  uint8_t*  used_rrf;     /* [B] 1 = fused (:333), 0 = vector-only branch (:346)    */
  /* the filtered vector stage (optional, may be NULL) */
  uint64_t* vec_ids;      /* [B][vector_top_k]                                      */
  double*   vec_scores;   /* [B][vector_top_k]                                      */
  uint32_t* vec_counts;   /* [B]                                                    */
  uint8_t*  certified;    /* [B] optional: 1 = proven equal to the reference's result. The score filter is part of the
                             proof: rows that provably lie below min_vector_score (below min_relevance for
                             rag_memory_retrieve) can never reach the result, so a query whose non-candidates all
                             lie there is certified whatever its k-th score is, and the tensor path never makes such
                             rows candidates (filtering commutes with taking the best k) */
} rag_fused_out;

/* MemoryStore.retrieve(query, limit, minRelevance) — src/lib/memory/store.ts:102-180 */
typedef struct rag_memory_opts {
  uint32_t limit;
  uint32_t path;
  double   min_relevance;     /* default 0.5 at the call site                       */
  int64_t  now_ms;
  double   time_decay_factor; /* 0 → 0.05                                           */
  double   frequency_bonus;   /* 0 → 0.1                                            */
  uint32_t similarity_top_k;  /* retriever similarityTopK; 0 → limit * 2 (store.ts:112). A host that must drop
                                 memories whose DB record is gone (`if (dbMemory)`, store.ts:153) asks for
                                 limit = similarity_top_k = 2L, filters, and keeps the first L        */
  uint32_t flags;             /* RAG_SEARCH_* (escalation is always on here)                       */
} rag_memory_opts;

typedef struct rag_memory_out {
  uint64_t* ids;        /* [B][limit]                                               */
  double*   scores;     /* [B][limit] 0.7·relevance + 0.3·freshness (store.ts:160)  */
  double*   relevance;  /* [B][limit] cosine                                        */
  double*   freshness;  /* [B][limit]                                               */
  uint32_t* counts;     /* [B]                                                      */
} rag_memory_out;

/* ---- library ------------------------------------------------------------- */
int rag_version(void);
const char* rag_last_error(void);
int rag_device_count(void);

/* ---- index lifetime: replaces VectorStoreIndex.init / SimpleVectorStore
 *      (src/lib/llm/index-manager.ts:218-227,264-270) -------------------------- */
int rag_index_create(const rag_index_desc* desc, rag_index** out);
void rag_index_destroy(rag_index* idx);
/* copy host rows (dtype of the index, row-major [nrows][dim]) into rows
 * [row0,row0+nrows); extends the row count; index.insert (memory/store.ts:67) appends */
int rag_index_upload(rag_index* idx, uint64_t row0, uint64_t nrows, const void* host_rows);
/* fill rows [0,nrows) with the synthetic corpus of include/ragera_gen.h on the device
 * (global row = id_base + local row); also fills row meta from the generator */
int rag_index_generate(rag_index* idx, const rag_gen_desc* gen, uint64_t nrows);
/* per-row metadata: metadata.type (hybrid-search.ts:229-234) and the Memory columns
 * confidence/accessCount/lastAccessedAt (prisma/schema.prisma:94-99). Any pointer may be NULL. */
int rag_index_set_row_meta(rag_index* idx, uint64_t row0, uint64_t nrows, const uint8_t* content_type,
                           const double* confidence, const int32_t* access_count,
                           const int64_t* last_access_ms);
/* optional row → fusion key map (key of content.substring(0,100), hybrid-search.ts:149);
 * default key = chunk id */
int rag_index_set_row_keys(rag_index* idx, uint64_t row0, uint64_t nrows, const uint64_t* keys);
int rag_index_read_rows(rag_index* idx, uint64_t row0, uint64_t nrows, void* host_rows);
/* read back what rag_index_set_row_meta / rag_index_set_row_keys hold (any pointer may be NULL); rows whose metadata
 * was never set read as documents with zeroed Memory columns, keys default to the chunk id */
int rag_index_read_row_meta(rag_index* idx, uint64_t row0, uint64_t nrows, uint8_t* content_type, double* confidence,
                            int32_t* access_count, int64_t* last_access_ms, uint64_t* keys);
uint64_t rag_index_rows(const rag_index* idx);
/* synthetic queries of ragera_gen.h, generated on the device, copied to host_out [B][dim] */
int rag_generate_queries(rag_index* idx, const rag_gen_desc* gen, uint64_t b0, uint32_t B, float* host_out);

/* ---- the reference's persisted index: llamaindex `vector_store.json` under
 *      ./storage/kb_<id>/ (src/lib/llm/index-manager.ts:218-220,264-270). Rows are appended in file
 *      order (= the insertion order the reference scans). `ids`, if non-NULL, receives the node ids
 *      as a '\0'-separated blob allocated by the library — release it with rag_free(). ------------- */
int rag_index_load_vector_store(rag_index* idx, const char* vector_store_json, uint64_t* rows_loaded,
                                char** ids, uint64_t* ids_bytes);
/* the host-side parser on its own (no GPU needed): calls on_rows for every slab of <= slab_rows rows */
int rag_parse_vector_store_json(const char* path, uint32_t dim, uint64_t slab_rows,
                                int (*on_rows)(void* user, uint64_t first_row, uint64_t nrows, const float* rows),
                                void* user, uint64_t* rows_out, char** ids, uint64_t* ids_bytes);
/* the same parser with a RESUME point: resume_offset = 0 parses the whole file; otherwise it is the *end_offset of an
 * earlier parse — the byte just after the ']' of the last embedding that parse consumed — and only the embeddings that
 * follow are parsed (index.insert appends to embeddingDict; src/lib/memory/store.ts:56-67). *end_offset: the resume
 * point after this parse. stamp_*: size and mtime of the file, taken from the descriptor that was parsed before reading
 * it. The number text is converted on all host threads (RAGERA_LOADER_THREADS to override). */
int rag_parse_vector_store_json_ex(const char* path, uint32_t dim, uint64_t slab_rows,
                                   int (*on_rows)(void* user, uint64_t first_row, uint64_t nrows, const float* rows),
                                   void* user, uint64_t resume_offset, uint64_t* rows_out, char** ids, uint64_t* ids_bytes,
                                   uint64_t* end_offset, uint64_t* stamp_size, int64_t* stamp_mtime_ns);
/* hash of the first nbytes bytes of a file (threaded); *ok = 0 when the file is shorter */
int rag_file_prefix_hash(const char* path, uint64_t nbytes, uint64_t* hash, int* ok);
void rag_free(void* p);
/* second pass over the same file: "metadataDict" (node metadata next to the embeddings). content_type[row] follows
 * vectorSearch's rule (src/lib/hybrid-search.ts:229-234, without the per-call isCodebase flag): metadata.type ===
 * 'memory' → RAG_CT_MEMORY, else metadata.language !== undefined (key present) → RAG_CT_CODE, else RAG_CT_DOCUMENT. memory_ids (optional):
 * metadata.memoryId per row (src/lib/memory/store.ts:58-61), '\0'-terminated, empty when absent, rag_free() it — the
 * host joins it with the Prisma Memory rows for rag_index_set_row_meta. ids/ids_bytes: the blob of the first pass.
 * *found = 0 when the file has no metadataDict. Host-only. rag_index_load_vector_store applies content_type itself. */
int rag_parse_vector_store_metadata(const char* path, const char* ids, uint64_t ids_bytes, uint64_t rows,
                                    uint8_t* content_type, char** memory_ids, uint64_t* memory_ids_bytes, int* found);

/* ---- binary sidecar of a persisted index: the parsed rows in the index dtype + row metadata +
 *      fusion keys + node ids, checksummed per 4096-row block, stamped with size/mtime of the JSON it was made from
 *      (index.insert appends to the JSON — src/lib/memory/store.ts:67 — which makes the sidecar stale). Cold start
 *      (loadIndex, src/lib/llm/index-manager.ts:246-275) then streams binary rows into HBM instead of parsing text. */
#define RAG_CACHE_META 1u /* content_type / confidence / access_count / last_access_ms present */
#define RAG_CACHE_KEYS 2u /* fusion keys present                                                */
typedef struct rag_cache_info {
  uint32_t version, dtype, dim, flags;
  uint64_t rows, ids_bytes;
  uint64_t source_size;     /* of the vector_store.json the sidecar was written from (0 = unknown) */
  int64_t  source_mtime_ns;
  uint64_t source_prefix_bytes; /* where the sidecar's last embedding ended in that JSON (0 = unknown): the resume point */
  uint64_t source_prefix_hash;  /* rag_file_prefix_hash of the JSON up to there                                          */
} rag_cache_info;
/* host-only: header (validated), freshness against the JSON (1 fresh / 0 stale, missing or malformed), whole-file
 * write and read with every checksum verified (any output pointer may be NULL; metadata arrays: all four or none) */
int rag_cache_info_read(const char* cache_path, rag_cache_info* out);
int rag_cache_is_fresh(const char* cache_path, const char* source_json);
int rag_cache_write_host(const char* cache_path, uint32_t dtype, uint32_t dim, uint64_t rows, const void* host_rows,
                         const uint8_t* content_type, const double* confidence, const int32_t* access_count,
                         const int64_t* last_access_ms, const uint64_t* keys, const char* ids, uint64_t ids_bytes,
                         const char* source_json /* may be NULL */);
int rag_cache_read_host(const char* cache_path, uint64_t first_row, uint64_t nrows, void* host_rows, uint8_t* content_type,
                        double* confidence, int32_t* access_count, int64_t* last_access_ms, uint64_t* keys, char** ids,
                        uint64_t* ids_bytes);
/* device side: save this handle's rows (+ metadata / keys if set) with the caller's node ids; append a row range of
 * a sidecar to the handle (nrows = 0: to the end; a shard passes its own range); open a store through its sidecar:
 * rag_cache_refresh_host, then the binary rows into HBM (*from_cache = the refresh route; cache_path NULL →
 * "<json>.ragera"; if the sidecar cannot be written the JSON is parsed straight into the index) */
int rag_index_save_cache(rag_index* idx, const char* cache_path, const char* ids, uint64_t ids_bytes, const char* source_json);
/* host-only: bring the sidecar of `source_json` up to date. *route: 1 = it was fresh, 2 = the JSON had grown by appended
 * embeddings (only those were parsed; the sidecar was extended in place), 0 = full parse + rewrite. cache_path NULL →
 * "<source_json>.ragera". RAG_ERR_STATE if the JSON changes while it is being read (no sidecar is left behind). */
int rag_cache_refresh_host(const char* cache_path, const char* source_json, uint32_t dtype, uint32_t dim, int* route,
                           uint64_t* rows_out);
int rag_index_load_cache(rag_index* idx, const char* cache_path, uint64_t first_row, uint64_t nrows, uint64_t* rows_loaded,
                         char** ids, uint64_t* ids_bytes);
int rag_index_open_store(rag_index* idx, const char* vector_store_json, const char* cache_path, uint64_t* rows_loaded,
                         char** ids, uint64_t* ids_bytes, int* from_cache);

/* ---- search: replaces retriever.retrieve → SimpleVectorStore.query →
 *      getTopKEmbeddings (src/lib/hybrid-search.ts:223-224) ---------------------- */
int rag_search(rag_index* idx, const float* queries /*[B][dim] host*/, uint32_t B,
               const rag_search_opts* opts, rag_topk_out* out);

/* ---- hybrid search: replaces the body of hybridSearch() after the embedding and
 *      Meilisearch round-trips (src/lib/hybrid-search.ts:303-354). kw_keys holds each
 *      query's keyword hits as fusion keys in Meilisearch rank order (meilisearch.ts:
 *      226-236); kw_counts[b] = 0 selects the vector-only branch for that query. */
int rag_hybrid_search(rag_index* idx, const float* queries, uint32_t B, const rag_hybrid_opts* opts,
                      const uint64_t* kw_keys /*[B][keyword_limit]*/, const uint32_t* kw_counts /*[B]*/,
                      rag_fused_out* out);

/* ---- fusion only: replaces reciprocalRankFusion() (src/lib/hybrid-search.ts:129-208)
 *      for B independent (vector list, keyword list) pairs given as keys. */
int rag_rrf_fuse(rag_index* idx, uint32_t B, const rag_rrf_config* cfg,
                 const uint64_t* vec_keys /*[B][vec_stride]*/, const uint8_t* vec_ctype /*may be NULL*/,
                 const uint32_t* vec_counts, uint32_t vec_stride,
                 const uint64_t* kw_keys /*[B][kw_stride]*/, const uint32_t* kw_counts, uint32_t kw_stride,
                 rag_fused_out* out);

/* ---- memory: replaces MemoryStore.retrieve post-processing and
 *      calculateFreshnessScore (src/lib/memory/store.ts:102-180, freshness.ts:37-56) */
int rag_memory_retrieve(rag_index* idx, const float* queries, uint32_t B, const rag_memory_opts* opts,
                        rag_memory_out* out);
int rag_freshness_scores(rag_index* idx, uint64_t n, const double* confidence, const int32_t* access_count,
                         const int64_t* last_access_ms, int64_t now_ms, double time_decay_factor,
                         double frequency_bonus, double* out_scores);

/* ---- staged (device-resident) form of rag_hybrid_search, for callers that keep
 *      a batch in flight and for measurement: stage → run (async) → fetch. ------- */
int rag_stage_batch(rag_index* idx, const float* queries, uint32_t B, const uint64_t* kw_keys,
                    const uint32_t* kw_counts, uint32_t keyword_limit);
/* select the window [first, first+count) of the staged pool as the batch of the following
 * staged runs (default after rag_stage_batch: the whole pool) */
int rag_stage_window(rag_index* idx, uint32_t first, uint32_t count);
int rag_hybrid_search_staged(rag_index* idx, uint32_t B, const rag_hybrid_opts* opts); /* async */
int rag_fetch_fused(rag_index* idx, uint32_t B, const rag_hybrid_opts* opts, rag_fused_out* out);
int rag_sync(rag_index* idx);

/* ---- host-side post-filter: processResults of src/lib/context/rag/dedup-filter.ts:193-247,
 *      the step ContextEngine.buildContext applies to the fused list (engine.ts:289). Strings are UTF-16
 *      code units (what a JS string is); no GPU involved. opts == NULL → the reference's defaults. -------- */
typedef struct rag_text { const uint16_t* units; uint32_t len; } rag_text;
typedef struct rag_process_opts {
  double   similarity_threshold; /* 0.85  DEFAULT_CONFIG, dedup-filter.ts:16-20 */
  uint32_t min_content_length;   /* 20                                          */
  uint32_t max_results;          /* 10                                          */
  uint32_t enable_noise_filter;  /* 1     (enableNoiseFiltler)                  */
  uint32_t enable_rerank;        /* 1                                           */
} rag_process_opts;
typedef struct rag_processed_out {
  uint32_t  capacity;      /* entries in the arrays below (>= min(n, max_results))      */
  uint32_t* index;         /* [capacity] index of the surviving input result            */
  double*   fusion_score;  /* [capacity] FusedResult.fusionScore                        */
  uint8_t*  deduplicated;  /* [capacity] optional                                       */
  uint32_t* source_mask;   /* [capacity] optional: bit s set if a merged result had source s */
  uint32_t* n_sources;     /* [capacity] optional: FusedResult.sources.length           */
  uint32_t  count;         /* out: number of results                                    */
} rag_processed_out;
int rag_process_results(const rag_text* contents, const double* scores, const uint8_t* sources, uint32_t n,
                        rag_text query, const rag_process_opts* opts, rag_processed_out* out);

/* ---- micro-batching front end: concurrent batch-1 callers (one per request thread,
 *      like the reference's per-request hybridSearch) share one corpus pass. One batcher per call-site
 *      class (fixed options). submit() blocks until the caller's own result is ready; results are
 *      identical to a direct batch-1 call. Callers stage their own inputs into the open batch and copy their own results
 *      out; two worker threads take turns running batches (the only users of the index handle). */
typedef struct rag_batcher rag_batcher;
typedef struct rag_batcher_desc {
  uint32_t max_batch;    /* 1..4096 queries per pass                                    */
  uint32_t max_wait_us;  /* CAP on how long the first request of a batch waits for company: the batch goes out earlier,
                            as soon as the arrivals pause (20-50 us without a new request) and no finished batch's callers
                            are still being woken; ~1000 keeps closed-loop callers in one pass (profiles/r02_batcher_load.md) */
  rag_hybrid_opts opts;  /* HybridSearchOptions of this call site (keyword_limit = row stride) */
} rag_batcher_desc;
/* RAG_ERR_UNSUPPORTED on a row-sharded index (nranks > 1): the exchange is collective, every rank must run the same batches —
 * batch on one front end and hand each batch to all ranks. */
int rag_batcher_create(rag_index* idx, const rag_batcher_desc* desc, rag_batcher** out);
int rag_batcher_submit(rag_batcher* b, const float* query /*[dim]*/, const uint64_t* kw_keys, uint32_t kw_count,
                       rag_fused_out* out /* shaped for one query */);
/* The same request without parking a thread on it — what a single-threaded host (Node's event loop; libuv's pool has 4
 * threads by default, so blocking submits could never form a batch above 4) uses to keep thousands of requests in
 * flight. Takes a slot of the open batch, copies `query` and `kw_keys` into it (the caller may reuse them on return) and
 * returns. `out` (shaped for one query, as above) must stay valid until `done(user, rc, err)` runs — on a batcher worker
 * thread, after the result has been copied into `out` (rc == RAG_OK) or with the batch's error text (`err` is valid
 * during the callback only; it must not call rag_batcher_destroy). RAG_ERR_BUSY: every batch buffer is in flight,
 * nothing was queued — retry, or fall back to the blocking call. A non-zero return never calls `done`. */
typedef void (*rag_batcher_done_fn)(void* user, int rc, const char* err);
int rag_batcher_submit_async(rag_batcher* b, const float* query /*[dim]*/, const uint64_t* kw_keys, uint32_t kw_count,
                             rag_fused_out* out /* shaped for one query */, rag_batcher_done_fn done, void* user);
int rag_batcher_stats(rag_batcher* b, uint64_t* batches, uint64_t* queries, uint64_t* largest_batch);
/* destroy may be called while submit calls are still inside the batcher: requests that hold a slot are run and answered,
 * (asynchronous ones get their `done`), callers still waiting for a batch buffer fail with RAG_ERR_STATE, and nothing is
 * freed before the last of them has returned. No submit may START on the handle once destroy has been called. */
void rag_batcher_destroy(rag_batcher* b);

/* ---- diagnostics: the raw scaled scores (dot_bf16 * 1/||x||, no 1/||q||) the tensor path (K2)
 *      computes, written as out_scores[B][rows]; for validating the tcgen05 pipeline on small inputs */
int rag_debug_tensor_scores(rag_index* idx, const float* queries, uint32_t B, float* out_scores);
/*      the same launch followed by the K3 merge: out_keys[B][kp] = the K' packed candidate keys per query
 *      (ordered score bits << 32 | ~row, best first, 0 = empty) next to the scores they were selected from,
 *      so the fused tcgen05 selection can be checked exactly (ties included) on small inputs */
int rag_debug_tensor_candidates(rag_index* idx, const float* queries, uint32_t B, uint32_t kp, float* out_scores,
                                uint64_t* out_keys);

/*      rho_x of the rigorous certification bound: max over the loaded rows of ||x - operand(x)|| / ||x||, where
 *      operand(x) is what the tensor path multiplies (the bf16 shadow row, or the row read as tf32; 0 for a bf16
 *      corpus, whose rows are the operand). Returns a negative rag_status on error. */
double rag_index_row_residual(rag_index* idx);

/* ---- measurement helpers (CUDA events on the library's own stream) ------------ */
int rag_timer_start(rag_index* idx);
int rag_timer_stop(rag_index* idx, float* elapsed_ms);           /* synchronises */
uint64_t rag_launch_count(const rag_index* idx);                 /* kernels launched so far */
/* the successor of a sharded-exchange counter value (20 bits, never 0, consecutive values differ in parity across the wrap) */
uint32_t rag_debug_p2p_next_step(uint32_t step);
/* queries finished by the fusion kernel since the last call, and how many of them the scoring pass certified
 * (counted on the device: covers the asynchronous *_staged runs, which fetch nothing) */
int rag_certified_totals(rag_index* idx, uint64_t* certified, uint64_t* queries);
/* per-kernel device time: when enabled, every pipeline kernel launch is bracketed by a
 * pair of events on the library stream; rag_profile_read synchronises and returns the
 * summed milliseconds and launch counts per kernel class since the last read. */
enum { RAG_PROF_STREAM = 0, RAG_PROF_TENSOR = 1, RAG_PROF_MERGE = 2, RAG_PROF_RESCORE = 3,
       RAG_PROF_FUSE = 4, RAG_PROF_COMM = 5, RAG_PROF_CLASSES = 6 };
int rag_profile_enable(rag_index* idx, int on);
int rag_profile_read(rag_index* idx, float ms[RAG_PROF_CLASSES], uint32_t counts[RAG_PROF_CLASSES]);
/* pinned host memory for query / result buffers (what the N-API shim backs its
 * Float32Array with, so H2D/D2H are true async DMA) */
void* rag_host_alloc(uint64_t bytes);
void rag_host_free(void* p);

/* ---- row-sharded multi-GPU (one process per GPU of one node — or several processes on one GPU; every rank's
 *      exact local top-k is exchanged and merged on every rank, rank 0 is the consumer).
 *
 *      Default exchange: peer-to-peer MAILBOXES (device memory shared through CUDA IPC, over NVLink between GPUs),
 *      written and awaited inside the final fusion kernel — no collective call on the hot path. The bootstrap is
 *      HOST-DRIVEN and needs no NCCL: every rank exports a 64-byte handle of its own mailbox, the host exchanges
 *      the handles by whatever channel it has (torch.distributed, a pipe, Node's cluster messaging), every rank
 *      imports all of them:
 *          rag_comm_p2p_export(idx, nranks, rank, max_batch, max_k, mine);
 *          ... host: all-gather of the RAG_COMM_HANDLE_BYTES-byte handles ...
 *          rag_comm_p2p_import(idx, all_handles);
 *      A search whose batch or k exceeds (max_batch, max_k) fails with RAG_ERR_STATE: export again, larger.
 *      rag_comm_p2p_import returns RAG_ERR_UNSUPPORTED when a peer's memory cannot be mapped (no P2P access):
 *      the host then falls back to the NCCL path below on ALL ranks.
 *
 *      NCCL path: rag_comm_unique_id / rag_comm_init — one ncclAllGather of the [B][k] records before the fusion
 *      kernel (libnccl is dlopen'ed; environment RAGERA_COMM=nccl selects this exchange, otherwise rag_comm_init
 *      bootstraps the same mailboxes by carrying the handles over NCCL).
 *
 *      All ranks must issue the same sequence of search calls (same batch size and k): the exchange, like a
 *      collective, pairs the i-th call of every rank. A rank that waits longer than RAGERA_P2P_TIMEOUT_MS
 *      (default 20000) for a peer gives up: the call returns RAG_ERR_TIMEOUT (no device trap, the context stays
 *      usable), the communicator is marked broken and every later sharded call fails with RAG_ERR_STATE until the
 *      ranks bootstrap again. ---------------------------------------------------------------------------- */
#define RAG_COMM_ID_BYTES 128
#define RAG_COMM_HANDLE_BYTES 64
int rag_comm_p2p_export(rag_index* idx, int nranks, int rank, uint32_t max_batch, uint32_t max_k,
                        uint8_t handle[RAG_COMM_HANDLE_BYTES]);
int rag_comm_p2p_import(rag_index* idx, const uint8_t* handles /* [nranks][RAG_COMM_HANDLE_BYTES], by rank */);
/* Teardown in two phases (freeing a mailbox a peer still maps is undefined in CUDA): every rank calls
 * rag_comm_detach (closes its mappings of the peers' mailboxes), the host runs a barrier, then rag_comm_destroy /
 * rag_index_destroy / a new rag_comm_p2p_export may free. */
int rag_comm_detach(rag_index* idx);
int rag_comm_unique_id(uint8_t id[RAG_COMM_ID_BYTES]);            /* rank 0 creates, host broadcasts */
int rag_comm_init(rag_index* idx, int nranks, int rank, const uint8_t id[RAG_COMM_ID_BYTES]);
int rag_comm_destroy(rag_index* idx);

#ifdef __cplusplus
}
#endif
#endif /* RAGERA_H */
