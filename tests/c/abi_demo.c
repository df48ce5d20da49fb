/* abi_demo.c — the C ABI from plain C (no Python, no C++): what an FFI / N-API shim does.
 * Build: gcc -std=c11 -I include tests/c/abi_demo.c -o abi_demo -L rag_era_b200 -lragera -Wl,-rpath,$PWD/rag_era_b200
 * Generates a small synthetic index on the device, runs deep_search-shaped hybrid searches through
 * rag_hybrid_search and prints one line per query:  <b> <used_rrf> <count> <key0> <score0 as hex float> <certified>;
 * then sends the same queries through rag_batcher_submit_async (completion callback on a worker thread) and prints them again */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "ragera.h"

/* completion of an asynchronous batcher request: called on a batcher worker thread */
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_cv = PTHREAD_COND_INITIALIZER;
static int g_done = 0, g_failed = 0;
static void on_done(void* user, int rc, const char* err) {
  (void)user; (void)err;
  pthread_mutex_lock(&g_mu);
  g_done++;
  if (rc != RAG_OK) g_failed++;
  pthread_cond_signal(&g_cv);
  pthread_mutex_unlock(&g_mu);
}

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
  /* the same three requests through the micro-batcher's non-blocking entry point: submit all, then wait for the callbacks;
   * one more line per query, which must repeat the line above */
  rag_batcher_desc bd;
  memset(&bd, 0, sizeof bd);
  bd.max_batch = 8; bd.max_wait_us = 500; bd.opts = o;
  rag_batcher* bt = NULL;
  CHECK(rag_batcher_create(idx, &bd, &bt));
  uint64_t akeys[3 * 14]; double ascores[3 * 14]; uint8_t arrf[3], acert[3]; uint32_t acounts[3];
  rag_fused_out aout[3];
  for (uint32_t b = 0; b < B; b++) {
    rag_fused_out one = { cap, akeys + b * cap, ascores + b * cap, NULL, NULL, acounts + b, arrf + b, NULL, NULL, NULL, acert + b };
    aout[b] = one;
    CHECK(rag_batcher_submit_async(bt, q + (size_t)b * dim, kw + b * kl, kwc[b], &aout[b], on_done, NULL));
  }
  pthread_mutex_lock(&g_mu);
  while (g_done < (int)B) pthread_cond_wait(&g_cv, &g_mu);
  pthread_mutex_unlock(&g_mu);
  if (g_failed) { fprintf(stderr, "an asynchronous request failed\n"); return 3; }
  for (uint32_t b = 0; b < B; b++)
    printf("%u %u %u %llu %a %u\n", b, arrf[b], acounts[b], (unsigned long long)akeys[b * cap], ascores[b * cap], acert[b]);
  rag_batcher_destroy(bt);
  rag_host_free(q);
  rag_index_destroy(idx);
  return 0;
}
