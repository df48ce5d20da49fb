/*
 * oracle.c — CPU restatement of rag-era's retrieval hot path. See oracle.h:
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (no reference tests/golden vectors exist).
 *
 * Compile with:  gcc -O2 -ffp-contract=off -fno-fast-math -pthread -fPIC -shared
 * (V8 never fuses a*b+c and never reassociates; neither may this file.)
 *
 * All arithmetic is IEEE-754 binary64 like JavaScript numbers. Stored fp32 / bf16
 * embedding values are widened exactly.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* minimal row-parallel helper (pthreads): rows are independent, so threading
 * never changes a bit of any score */
typedef void (*par_fn)(int64_t lo, int64_t hi, void* ctx);
typedef struct { par_fn fn; void* ctx; int64_t lo, hi; } par_job;
static void* par_tramp(void* p) { par_job* j = (par_job*)p; j->fn(j->lo, j->hi, j->ctx); return NULL; }
static void par_for(int64_t n, int threads, par_fn fn, void* ctx) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  if (threads == 1 || n < 2 * threads) { fn(0, n, ctx); return; }
  pthread_t th[256]; par_job jobs[256];
  int64_t chunk = (n + threads - 1) / threads;
  int started = 0;
  for (int t = 0; t < threads; t++) {
    int64_t lo = t * chunk, hi = lo + chunk < n ? lo + chunk : n;
    if (lo >= hi) break;
    jobs[t].fn = fn; jobs[t].ctx = ctx; jobs[t].lo = lo; jobs[t].hi = hi;
    pthread_create(&th[t], NULL, par_tramp, &jobs[t]);
    started++;
  }
  for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

/* ------------------------------------------------------------------------- */
/* similarity(): @llamaindex/core@0.6.22 embeddings utils (upstream-recalled). */
/*   DOT_PRODUCT: result = 0; for i: result += e1[i]*e2[i]                    */
/*   norm(x):     result = 0; for i: result += x[i]*x[i]; sqrt(result)        */
/*   DEFAULT:     dot / (norm(e1) * norm(e2))   — nothing cached, no epsilon  */
/* ------------------------------------------------------------------------- */

static double norm_f32(const float* x, uint32_t d) {
  double r = 0.0;
  for (uint32_t i = 0; i < d; i++) r += (double)x[i] * (double)x[i];
  return sqrt(r);
}

double oracle_cosine_f32(const float* q, const float* x, uint32_t d) {
  double dot = 0.0;
  for (uint32_t i = 0; i < d; i++) dot += (double)q[i] * (double)x[i];
  return dot / (norm_f32(q, d) * norm_f32(x, d));
}

double oracle_cosine_bf16(const float* q, const uint16_t* x, uint32_t d) {
  double dot = 0.0, nx = 0.0;
  for (uint32_t i = 0; i < d; i++) dot += (double)q[i] * (double)rg_bf16_to_f32(x[i]);
  for (uint32_t i = 0; i < d; i++) {
    double v = (double)rg_bf16_to_f32(x[i]);
    nx += v * v;
  }
  return dot / (norm_f32(q, d) * sqrt(nx));
}

/* ------------------------------------------------------------------------- */
/* getTopKEmbeddings (upstream-recalled): push {similarity,id} for every row  */
/* in insertion order, similarities.sort((a,b)=>b.similarity-a.similarity)    */
/* (stable), take the first k. Zero-norm rows give NaN upstream; this build   */
/* defines them as never selected (SURVEY N-nan) — keep them out of datasets. */
/* ------------------------------------------------------------------------- */

typedef struct { double s; uint64_t id; } scored;

static void stable_sort_desc(scored* a, uint64_t n) {
  if (n < 2) return;
  scored* tmp = (scored*)malloc(n * sizeof(scored));
  scored *src = a, *dst = tmp;
  for (uint64_t w = 1; w < n; w *= 2) {
    for (uint64_t lo = 0; lo < n; lo += 2 * w) {
      uint64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
      uint64_t i = lo, j = mid, o = lo;
      while (i < mid && j < hi) {
        /* comparator b.s - a.s: right element moves first only if strictly greater */
        if (src[j].s > src[i].s) dst[o++] = src[j++]; else dst[o++] = src[i++];
      }
      while (i < mid) dst[o++] = src[i++];
      while (j < hi) dst[o++] = src[j++];
    }
    scored* t = src; src = dst; dst = t;
  }
  if (src != a) memcpy(a, src, n * sizeof(scored));
  free(tmp);
}

/* stable partial selection: same output as stable_sort_desc + take k */
static uint32_t select_insert(scored* best, uint32_t cnt, uint32_t k, double s, uint64_t id) {
  if (cnt == k && !(s > best[k - 1].s)) return cnt;
  uint32_t pos = cnt < k ? cnt : k - 1;
  while (pos > 0 && s > best[pos - 1].s) { best[pos] = best[pos - 1]; pos--; }
  best[pos].s = s; best[pos].id = id;
  return cnt < k ? cnt + 1 : cnt;
}

static double row_cosine(const void* X, int dtype, uint64_t row, uint32_t d, const float* q) {
  if (dtype == ORACLE_BF16) return oracle_cosine_bf16(q, (const uint16_t*)X + row * d, d);
  return oracle_cosine_f32(q, (const float*)X + row * d, d);
}

static int64_t finish_topk(double* sims, uint64_t n, uint32_t k, uint64_t id_base, uint64_t* ids,
                           double* scores, int faithful_sort) {
  int64_t out = 0;
  if (faithful_sort) {
    scored* all = (scored*)malloc((n ? n : 1) * sizeof(scored));
    uint64_t m = 0;
    for (uint64_t i = 0; i < n; i++)
      if (!isnan(sims[i])) { all[m].s = sims[i]; all[m].id = id_base + i; m++; }
    stable_sort_desc(all, m);
    for (uint64_t i = 0; i < k && i < m; i++) { ids[i] = all[i].id; scores[i] = all[i].s; out++; }
    free(all);
  } else {
    scored* best = (scored*)malloc((k ? k : 1) * sizeof(scored));
    uint32_t cnt = 0;
    if (k)
      for (uint64_t i = 0; i < n; i++)
        if (!isnan(sims[i])) cnt = select_insert(best, cnt, k, sims[i], id_base + i);
    for (uint32_t i = 0; i < cnt; i++) { ids[i] = best[i].id; scores[i] = best[i].s; }
    out = cnt;
    free(best);
  }
  return out;
}

typedef struct {
  const void* X; const rag_gen_desc* g; int dtype; uint64_t row0; uint32_t d; const float* q;
  double* sims;
} score_ctx;

static void score_rows(int64_t lo, int64_t hi, void* p) {
  score_ctx* c = (score_ctx*)p;
  if (c->X) {
    for (int64_t i = lo; i < hi; i++) c->sims[i] = row_cosine(c->X, c->dtype, (uint64_t)i, c->d, c->q);
  } else {
    void* row = malloc((size_t)c->d * 4);
    for (int64_t i = lo; i < hi; i++) {
      oracle_gen_rows(c->g, c->row0 + (uint64_t)i, 1, c->d, c->dtype, row, 1);
      c->sims[i] = row_cosine(row, c->dtype, 0, c->d, c->q);
    }
    free(row);
  }
}

int64_t oracle_topk(const void* X, int dtype, uint64_t n, uint32_t d, const float* q, uint32_t k,
                    uint64_t id_base, uint64_t* ids, double* scores, int faithful_sort,
                    int threads) {
  double* sims = (double*)malloc((n ? n : 1) * sizeof(double));
  score_ctx c = { X, NULL, dtype, 0, d, q, sims };
  par_for((int64_t)n, threads, score_rows, &c);
  int64_t r = finish_topk(sims, n, k, id_base, ids, scores, faithful_sort);
  free(sims);
  return r;
}

int64_t oracle_topk_generated(const rag_gen_desc* g, int dtype, uint64_t row0, uint64_t n,
                              uint32_t d, const float* q, uint32_t k, uint64_t* ids,
                              double* scores, int faithful_sort, int threads) {
  double* sims = (double*)malloc((n ? n : 1) * sizeof(double));
  score_ctx c = { NULL, g, dtype, row0, d, q, sims };
  par_for((int64_t)n, threads, score_rows, &c);
  int64_t r = finish_topk(sims, n, k, row0, ids, scores, faithful_sort);
  free(sims);
  return r;
}

/* ------------------------------------------------------------------------- */
/* src/lib/hybrid-search.ts:308-314 — drop r.score < minVectorScore, keep order */
/* ------------------------------------------------------------------------- */
uint32_t oracle_filter_min_score(uint64_t* ids, double* scores, uint32_t n, double min_score) {
  uint32_t m = 0;
  for (uint32_t i = 0; i < n; i++) {
    if (scores[i] < min_score) continue; /* :310 */
    ids[m] = ids[i]; scores[m] = scores[i]; m++;
  }
  return m;
}

/* ------------------------------------------------------------------------- */
/* reciprocalRankFusion — src/lib/hybrid-search.ts:129-208                    */
/* ------------------------------------------------------------------------- */
typedef struct {
  uint64_t key; double score; uint8_t source; uint8_t ctype; uint32_t order;
} rrf_entry;

static int32_t rrf_find(const rrf_entry* e, uint32_t n, uint64_t key) {
  for (uint32_t i = 0; i < n; i++) if (e[i].key == key) return (int32_t)i;
  return -1;
}

static uint32_t rrf_core(const uint64_t* vec_keys, const uint8_t* vec_ctype, uint32_t nv,
                         const uint64_t* kw_keys, uint32_t nk, const uint64_t* fr_keys,
                         uint32_t nf, double fresh_weight, const oracle_rrf_config* cfg,
                         uint64_t* out_keys, double* out_scores, uint8_t* out_source,
                         uint8_t* out_ctype) {
  const double k = cfg->k, vw = cfg->vector_weight, kw = cfg->keyword_weight,
               bonus = cfg->both_bonus;                                   /* :134 */
  uint32_t cap = nv + nk + nf;
  rrf_entry* map = (rrf_entry*)malloc((cap ? cap : 1) * sizeof(rrf_entry)); /* :136 (Map keeps insertion order) */
  uint32_t n = 0;

  for (uint32_t rank = 0; rank < nv; rank++) {                            /* :147 */
    double rrf = vw / (k + (double)rank + 1.0);                           /* :148 */
    int32_t at = rrf_find(map, n, vec_keys[rank]);                        /* :151 */
    if (at >= 0) {
      map[at].score = map[at].score + rrf;                                /* :154 */
      map[at].source = ORACLE_SRC_BOTH;                                   /* :155 */
    } else {
      map[n].key = vec_keys[rank]; map[n].score = rrf;                    /* :157-158 */
      map[n].source = ORACLE_SRC_VECTOR;                                  /* :161 */
      map[n].ctype = vec_ctype ? vec_ctype[rank] : ORACLE_CT_DOCUMENT;    /* :162 */
      map[n].order = n; n++;
    }
  }
  for (uint32_t rank = 0; rank < nk; rank++) {                            /* :169 */
    double rrf = kw / (k + (double)rank + 1.0);                           /* :170 */
    int32_t at = rrf_find(map, n, kw_keys[rank]);                         /* :173 */
    if (at >= 0) {
      map[at].score = map[at].score + (rrf + (bonus * map[at].score));    /* :176 */
      map[at].source = ORACLE_SRC_BOTH;                                   /* :177 */
    } else {
      map[n].key = kw_keys[rank]; map[n].score = rrf;                     /* :179-180 */
      map[n].source = ORACLE_SRC_KEYWORD;                                 /* :184 */
      map[n].ctype = ORACLE_CT_DOCUMENT;                                  /* :185 */
      map[n].order = n; n++;
    }
  }
  /* extension pass (not in the reference; SURVEY N-c4 ii), same form as :169-188 */
  for (uint32_t rank = 0; rank < nf; rank++) {
    double rrf = fresh_weight / (k + (double)rank + 1.0);
    int32_t at = rrf_find(map, n, fr_keys[rank]);
    if (at >= 0) {
      map[at].score = map[at].score + (rrf + (bonus * map[at].score));
      map[at].source = ORACLE_SRC_BOTH;
    } else {
      map[n].key = fr_keys[rank]; map[n].score = rrf;
      map[n].source = ORACLE_SRC_FRESHNESS;
      map[n].ctype = ORACLE_CT_MEMORY;
      map[n].order = n; n++;
    }
  }
  /* :191-202 entries in insertion order, then stable sort by b.score - a.score */
  scored* s = (scored*)malloc((n ? n : 1) * sizeof(scored));
  for (uint32_t i = 0; i < n; i++) { s[i].s = map[i].score; s[i].id = i; }
  stable_sort_desc(s, n);
  for (uint32_t i = 0; i < n; i++) {
    const rrf_entry* e = &map[s[i].id];
    out_keys[i] = e->key; out_scores[i] = e->score;
    out_source[i] = e->source; out_ctype[i] = e->ctype;
  }
  free(s); free(map);
  return n;
}

uint32_t oracle_rrf(const uint64_t* vec_keys, const uint8_t* vec_ctype, uint32_t nv,
                    const uint64_t* kw_keys, uint32_t nk, const oracle_rrf_config* cfg,
                    uint64_t* out_keys, double* out_scores, uint8_t* out_source,
                    uint8_t* out_ctype) {
  return rrf_core(vec_keys, vec_ctype, nv, kw_keys, nk, NULL, 0, 0.0, cfg, out_keys, out_scores,
                  out_source, out_ctype);
}

uint32_t oracle_rrf3(const uint64_t* vec_keys, const uint8_t* vec_ctype, uint32_t nv,
                     const uint64_t* kw_keys, uint32_t nk, const uint64_t* fr_keys, uint32_t nf,
                     double fresh_weight, const oracle_rrf_config* cfg, uint64_t* out_keys,
                     double* out_scores, uint8_t* out_source, uint8_t* out_ctype) {
  return rrf_core(vec_keys, vec_ctype, nv, kw_keys, nk, fr_keys, nf, fresh_weight, cfg, out_keys,
                  out_scores, out_source, out_ctype);
}

/* ------------------------------------------------------------------------- */
/* calculateFreshnessScore — src/lib/memory/freshness.ts:37-56                */
/* ------------------------------------------------------------------------- */
double oracle_freshness(double confidence, int32_t access_count, int64_t last_access_ms,
                        int64_t now_ms, double time_decay_factor, double frequency_bonus) {
  double hours = (double)(now_ms - last_access_ms) / 3600000.0;           /* :43 */
  double decay = exp(-time_decay_factor * hours);                         /* :46 */
  double fb = log((double)access_count + 1.0) * frequency_bonus;          /* :49 */
  double score = confidence * decay * (1.0 + fb);                         /* :52 */
  double m = score < 1.0 ? score : 1.0;                                   /* :55 Math.min(1,score) */
  return m > 0.0 ? m : 0.0;                                               /* :55 Math.max(0,·)     */
}

/* ------------------------------------------------------------------------- */
/* MemoryStore.retrieve — src/lib/memory/store.ts:119-175                     */
/* ------------------------------------------------------------------------- */
uint32_t oracle_memory_rank(const double* cos, const uint8_t* is_memory, const double* confidence,
                            const int32_t* access_count, const int64_t* last_access_ms, uint32_t n,
                            int64_t now_ms, uint32_t limit, double min_relevance,
                            uint32_t* out_index, double* out_score, double* out_fresh) {
  scored* sc = (scored*)malloc((n ? n : 1) * sizeof(scored));
  double* fr = (double*)malloc((n ? n : 1) * sizeof(double));
  uint32_t m = 0;
  for (uint32_t i = 0; i < n; i++) {
    if (!is_memory[i]) continue;                                          /* :119-122 */
    double rel = cos[i];                                                  /* :148 */
    if (rel < min_relevance) continue;                                    /* :151 */
    double f = oracle_freshness(confidence[i], access_count[i], last_access_ms[i], now_ms,
                                0.05, 0.1);                               /* :157, freshness.ts:20-23 */
    sc[m].s = rel * 0.7 + f * 0.3;                                        /* :160 */
    sc[m].id = i; fr[i] = f; m++;
  }
  stable_sort_desc(sc, m);                                                /* :172 */
  uint32_t out = m < limit ? m : limit;                                   /* :175 */
  for (uint32_t i = 0; i < out; i++) {
    out_index[i] = (uint32_t)sc[i].id; out_score[i] = sc[i].s; out_fresh[i] = fr[sc[i].id];
  }
  free(sc); free(fr);
  return out;
}

/* ------------------------------------------------------------------------- */
/* hybridSearch — src/lib/hybrid-search.ts:275-355                            */
/* ------------------------------------------------------------------------- */
uint32_t oracle_hybrid_search(const void* X, int dtype, uint64_t n, uint32_t d, const float* q,
                              uint32_t vector_top_k, double min_vector_score,
                              const uint64_t* kw_keys, uint32_t nk, const oracle_rrf_config* cfg,
                              const uint64_t* row_keys, const uint8_t* row_ctype,
                              uint64_t* vec_ids, double* vec_scores, uint32_t* n_vec,
                              uint64_t* out_keys, double* out_scores, uint8_t* out_source,
                              uint8_t* out_ctype, int* used_rrf, int threads) {
  /* :303 vectorSearch → retriever top-k */
  uint32_t nv = (uint32_t)oracle_topk(X, dtype, n, d, q, vector_top_k, 0, vec_ids, vec_scores, 0,
                                      threads);
  nv = oracle_filter_min_score(vec_ids, vec_scores, nv, min_vector_score); /* :308-314 */
  *n_vec = nv;
  uint64_t* vkeys = (uint64_t*)malloc((nv ? nv : 1) * sizeof(uint64_t));
  uint8_t* vct = (uint8_t*)malloc(nv ? nv : 1);
  for (uint32_t i = 0; i < nv; i++) {
    vkeys[i] = row_keys ? row_keys[vec_ids[i]] : vec_ids[i];
    vct[i] = row_ctype ? row_ctype[vec_ids[i]] : ORACLE_CT_DOCUMENT;      /* :229-234 */
  }
  uint32_t out;
  if (nk > 0) {                                                           /* :333 */
    out = oracle_rrf(vkeys, vct, nv, kw_keys, nk, cfg, out_keys, out_scores, out_source, out_ctype);
    *used_rrf = 1;
  } else {                                                                /* :346-354 */
    for (uint32_t i = 0; i < nv; i++) {
      out_keys[i] = vec_ids[i]; out_scores[i] = vec_scores[i];
      out_source[i] = ORACLE_SRC_VECTOR; out_ctype[i] = vct[i];
    }
    out = nv; *used_rrf = 0;
  }
  free(vkeys); free(vct);
  return out;
}

/* ------------------------------------------------------------------------- */
/* synthetic data (definition in include/ragera_gen.h)                        */
/* ------------------------------------------------------------------------- */
typedef struct { const rag_gen_desc* g; uint64_t row0; uint32_t d; int dtype; void* out; } gen_ctx;

static void gen_rows_range(int64_t lo, int64_t hi, void* p) {
  gen_ctx* x = (gen_ctx*)p;
  for (int64_t r = lo; r < hi; r++) {
    for (uint32_t c = 0; c < x->d; c++) {
      float v = rg_corpus_elem(x->g, x->row0 + (uint64_t)r, c);
      if (x->dtype == ORACLE_BF16) ((uint16_t*)x->out)[(uint64_t)r * x->d + c] = rg_f32_to_bf16(v);
      else ((float*)x->out)[(uint64_t)r * x->d + c] = v;
    }
  }
}

void oracle_gen_rows(const rag_gen_desc* g, uint64_t row0, uint64_t nrows, uint32_t d, int dtype,
                     void* out, int threads) {
  gen_ctx c = { g, row0, d, dtype, out };
  par_for((int64_t)nrows, threads, gen_rows_range, &c);
}

void oracle_gen_queries(const rag_gen_desc* g, uint64_t b0, uint32_t nb, uint32_t d, float* out) {
  for (uint32_t b = 0; b < nb; b++)
    for (uint32_t c = 0; c < d; c++) out[(uint64_t)b * d + c] = rg_query_elem(g, b0 + b, c);
}

void oracle_gen_meta(const rag_gen_desc* g, uint64_t row0, uint64_t nrows, uint8_t* ctype,
                     double* confidence, int32_t* access_count, int64_t* last_access_ms) {
  for (uint64_t i = 0; i < nrows; i++) {
    uint64_t r = row0 + i;
    ctype[i] = r < g->memory_rows ? ORACLE_CT_MEMORY : ORACLE_CT_DOCUMENT;
    confidence[i] = rg_meta_confidence(g, r);
    access_count[i] = rg_meta_access(g, r);
    last_access_ms[i] = rg_meta_last_access_ms(g, r);
  }
}
