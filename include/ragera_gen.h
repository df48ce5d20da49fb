/*
 * ragera_gen.h — counter-based synthetic embedding generator (data definition).
 *
 * This is NOT part of the reference's algorithm: it defines the synthetic corpus,
 * queries and memory metadata of SURVEY.md §8(d) so that a 61 GB corpus can be
 * produced on the device while any row can be regenerated bit-for-bit on the
 * host (oracle, tests, CPU baseline). Every value is produced from integer
 * hashing followed by a fixed sequence of IEEE-754 binary32 operations with
 * round-to-nearest and NO fused multiply-add, so host and device agree exactly.
 *
 * Build notes: device code uses __fmul_rn/__fadd_rn (never contracted); host
 * code must be compiled with -ffp-contract=off.
 *
 *   x[row][col] = scale(row) * ( centre[cluster(row)][col] + noise * g(row,col) )
 *   q[b][col]   = x[planted(b)][col] + qnoise * g'(b,col)
 *
 * g is a 4-fold Irwin–Hall sum of 16-bit uniforms (an approximately normal
 * variate with exact integer arithmetic), scaled to unit variance.
 */
#ifndef RAGERA_GEN_H
#define RAGERA_GEN_H

#include <stdint.h>

#if defined(__CUDACC__)
#define RG_HD __host__ __device__ __forceinline__
#else
#define RG_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define RG_MUL(a, b) __fmul_rn((a), (b))
#define RG_ADD(a, b) __fadd_rn((a), (b))
#else
#define RG_MUL(a, b) ((a) * (b))
#define RG_ADD(a, b) ((a) + (b))
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Synthetic-corpus description. All fields are plain data; zero-initialise and
 * fill. The same struct drives rag_index_generate() and the host generator. */
typedef struct rag_gen_desc {
  uint64_t seed;        /* corpus seed (0xC0FFEE)                 */
  uint64_t query_seed;  /* query seed  (0xBEEF)                               */
  uint64_t meta_seed;   /* memory-metadata seed (0xF00D)                      */
  uint64_t total_rows;  /* rows of the WHOLE corpus (all shards) — planted    */
                        /* query rows are drawn from [0,total_rows)           */
  uint32_t n_clusters;  /* number of cluster centres (4096)                   */
  float    noise;       /* per-row noise amplitude (0.6)                      */
  float    query_noise; /* query noise amplitude (0.5)                        */
  uint32_t dup_period;  /* if >0: row r with r%dup_period==dup_period-1 is an */
                        /* exact copy of row r-1 (exercises score ties)       */
  uint64_t memory_rows; /* rows [0,memory_rows) are content_type=memory       */
  int64_t  now_ms;      /* "now" used to place lastAccessedAt                 */
} rag_gen_desc;

#ifdef __cplusplus
}
#endif

RG_HD uint64_t rg_mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

RG_HD uint64_t rg_hash3(uint64_t seed, uint64_t a, uint64_t b) {
  return rg_mix64(rg_mix64(seed ^ rg_mix64(a)) + b);
}

/* approximately N(0,1): sum of four 16-bit uniforms, centred and scaled */
RG_HD float rg_gauss(uint64_t h) {
  int32_t s = (int32_t)(h & 0xFFFFu) + (int32_t)((h >> 16) & 0xFFFFu) +
              (int32_t)((h >> 32) & 0xFFFFu) + (int32_t)(h >> 48);
  float f = (float)(s - 131070);          /* exact: |s-131070| < 2^18 */
  return RG_MUL(f, 2.6428993e-05f);       /* 1/sqrt(4*(65536^2-1)/12) */
}

RG_HD uint64_t rg_src_row(const rag_gen_desc* g, uint64_t row) {
  if (g->dup_period > 1 && row % g->dup_period == g->dup_period - 1) return row - 1;
  return row;
}

RG_HD uint32_t rg_cluster(const rag_gen_desc* g, uint64_t row) {
  return (uint32_t)(rg_mix64(g->seed ^ 0xC1u ^ rg_mix64(row)) % g->n_clusters);
}

RG_HD float rg_row_scale(const rag_gen_desc* g, uint64_t row) {
  uint32_t u = (uint32_t)(rg_mix64(g->seed ^ 0x5CA1Eu ^ rg_mix64(row)) & 0xFFu);
  return RG_ADD(0.5f, RG_MUL((float)u, 0.00390625f)); /* [0.5,1.5), exact */
}

/* corpus element (fp32 definition; bf16 corpora round this with RNE) */
RG_HD float rg_corpus_elem(const rag_gen_desc* g, uint64_t row, uint32_t col) {
  uint64_t r = rg_src_row(g, row);
  uint32_t c = rg_cluster(g, r);
  float centre = rg_gauss(rg_hash3(g->seed ^ 0xCE27E5u, c, col));
  float nz = rg_gauss(rg_hash3(g->seed, r, col));
  float v = RG_ADD(centre, RG_MUL(g->noise, nz));
  return RG_MUL(rg_row_scale(g, r), v);
}

RG_HD uint64_t rg_planted_row(const rag_gen_desc* g, uint64_t b) {
  return rg_mix64(g->query_seed ^ 0x9A17u ^ rg_mix64(b)) % g->total_rows;
}

RG_HD float rg_query_elem(const rag_gen_desc* g, uint64_t b, uint32_t col) {
  float x = rg_corpus_elem(g, rg_planted_row(g, b), col);
  float nz = rg_gauss(rg_hash3(g->query_seed, b, col));
  return RG_ADD(x, RG_MUL(g->query_noise, nz));
}

/* round-to-nearest-even fp32 -> bf16 bit pattern (NaN not produced by the generator) */
RG_HD uint16_t rg_f32_to_bf16(float f) {
  union { float f; uint32_t u; } v;
  v.f = f;
  uint32_t u = v.u;
  if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((u >> 16) | 0x40u);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

RG_HD float rg_bf16_to_f32(uint16_t b) {
  union { float f; uint32_t u; } v;
  v.u = ((uint32_t)b) << 16;
  return v.f;
}

/* memory metadata: confidence~U(0.5,1), accessCount~Geom(1/2),
 * lastAccessed = now - hours*3.6e6 with hours in [0,96) (integer ms, exact) */
RG_HD double rg_meta_confidence(const rag_gen_desc* g, uint64_t row) {
  uint32_t u = (uint32_t)(rg_mix64(g->meta_seed ^ 0xC0Fu ^ rg_mix64(row)) & 0xFFFFu);
  return 0.5 + (double)u * (0.5 / 65536.0);
}
RG_HD int32_t rg_meta_access(const rag_gen_desc* g, uint64_t row) {
  uint64_t h = rg_mix64(g->meta_seed ^ 0xACCu ^ rg_mix64(row));
  int32_t n = 0;
  while ((h & 1u) && n < 40) { n++; h >>= 1; }
  return n;
}
RG_HD int64_t rg_meta_last_access_ms(const rag_gen_desc* g, uint64_t row) {
  uint64_t h = rg_mix64(g->meta_seed ^ 0x1A57u ^ rg_mix64(row));
  int64_t ms = (int64_t)(h % (96ull * 3600000ull));
  return g->now_ms - ms;
}

#endif /* RAGERA_GEN_H */
