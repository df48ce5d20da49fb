/* abi_demo.c — the C ABI from plain C (no Python, no C++): what an FFI / N-API shim does.
 * Build: gcc -std=c11 -I include tests/c/abi_demo.c -o abi_demo -L rag_era_b200 -lragera -Wl,-rpath,$PWD/rag_era_b200
 * Generates a small synthetic index on the device, runs deep_search-shaped hybrid searches through
 * rag_hybrid_search and prints one line per query:  <b> <used_rrf> <count> <key0> <score0 as hex float> */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ragera.h"

#define CHECK(x) do { int rc_ = (x); if (rc_ != RAG_OK) { fprintf(stderr, "%s failed (%d): %s\n", #x, rc_, rag_last_error()); return 2; } } while (0)

int main(int argc, char** argv) {
  const uint64_t rows = argc > 1 ? strtoull(argv[1], NULL, 10) : 20000;
  const uint32_t dim = 256, B = 3, k = 10, kl = 4;
  rag_index_desc d = { rows, dim, RAG_F32, 0, RAG_INDEX_BF16_SHADOW, 0 };
  rag_index* idx = NULL;
  CHECK(rag_index_create(&d, &idx));
  rag_gen_desc g;
  memset(&g, 0, sizeof g);
  g.seed = 0xC0FFEE; g.query_seed = 0xBEEF; g.meta_seed = 0xF00D; g.total_rows = rows;
  g.n_clusters = 32; g.noise = 0.6f; g.query_noise = 0.5f;
  CHECK(rag_index_generate(idx, &g, rows));
  float* q = (float*)rag_host_alloc((uint64_t)B * dim * sizeof(float));
  CHECK(rag_generate_queries(idx, &g, 0, B, q));

  rag_hybrid_opts o;
  memset(&o, 0, sizeof o);
  o.vector_top_k = k; o.keyword_limit = kl; o.min_vector_score = 0.3;
  o.rrf.k = 60; o.rrf.vector_weight = 1; o.rrf.keyword_weight = 1; o.rrf.both_bonus = 0.1;   /* document preset */
  uint64_t kw[3 * 4] = { 1, 2, 3, 4,   0, 0, 0, 0,   rows - 1, 7, 0, 0 };
  uint32_t kwc[3] = { 4, 0, 2 };                      /* query 1: Meilisearch returned nothing → vector-only branch */
  const uint32_t cap = k + kl;
  uint64_t keys[3 * 14]; double scores[3 * 14]; uint8_t src[3 * 14], ct[3 * 14], rrf[3], cert[3];
  uint32_t counts[3];
  rag_fused_out out = { cap, keys, scores, src, ct, counts, rrf, NULL, NULL, NULL, cert };
  CHECK(rag_hybrid_search(idx, q, B, &o, kw, kwc, &out));
  for (uint32_t b = 0; b < B; b++)
    printf("%u %u %u %llu %a %u\n", b, rrf[b], counts[b], (unsigned long long)keys[b * cap], scores[b * cap], cert[b]);
  rag_host_free(q);
  rag_index_destroy(idx);
  return 0;
}
