// common.cuh — shared declarations of libragera (B200 / sm_100a retrieval hot path).
//
// Data layout in HBM:
//   corpus    [capacity][ld]   fp32 or bf16, row-major, ld = dim rounded up to 256 elements,
//                              padding columns are zero (they add 0 to dot and norm)
//   shadow    [capacity][ld]   16-bit copy of an fp32 corpus (tensor path operand B): fp16 of the NORMALISED row
//                              (RAG_INDEX_F16_SHADOW) or bf16 of the row as it is (RAG_INDEX_BF16_SHADOW)
//   inv_norm  [capacity]       fp32 1/||x|| of the operand the tensor path reads
//   row meta  ctype u8, confidence f64, access i32, last_ms i64, key u64 (optional)
//
// Candidate keys: one u64 orders (score desc, row asc):  ordered(fp32 score) << 32 | ~row.
// 0 is the empty sentinel (every real score, even -inf, packs to a non-zero key).
#pragma once

#include <math.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>

#include "../../include/ragera.h"

#define RAG_WARP 32

// exchanged / finalised candidate record (exact fp64 score); 48 bytes
struct rag_rec {
  double score;    // exact cosine (reference order, fp64, no FMA); -inf = empty
  uint64_t id;     // global chunk id
  uint64_t key;    // fusion key
  double fresh;    // calculateFreshnessScore for memory rows, else 0
  uint32_t ctype;  // rag_content_type
  uint32_t flags;  // bit0: the producing rank could not certify its local top-k
  double conf_pad; // reserved (keeps the record 16-byte aligned for 128-bit copies)
};
static_assert(sizeof(rag_rec) == 48, "rag_rec layout");

__host__ __device__ __forceinline__ uint32_t rag_order_f32(float f) {
  uint32_t b = __builtin_bit_cast(uint32_t, f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float rag_unorder_f32(uint32_t o) {
  uint32_t b = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
  return __builtin_bit_cast(float, b);
}
__host__ __device__ __forceinline__ uint64_t rag_pack_key(float score, uint32_t row) {
  return ((uint64_t)rag_order_f32(score) << 32) | (uint64_t)(0xFFFFFFFFu - row);
}
__host__ __device__ __forceinline__ uint32_t rag_key_row(uint64_t key) {
  return 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFu);
}
__host__ __device__ __forceinline__ float rag_key_score(uint64_t key) {
  return rag_unorder_f32((uint32_t)(key >> 32));
}

#ifdef __CUDACC__
__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
  uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
  uint32_t mhi = __reduce_max_sync(0xFFFFFFFFu, hi);
  uint32_t mlo = __reduce_max_sync(0xFFFFFFFFu, hi == mhi ? lo : 0u);
  return ((uint64_t)mhi << 32) | mlo;
}
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  uint32_t hi = __shfl_sync(0xFFFFFFFFu, (uint32_t)(v >> 32), src);
  uint32_t lo = __shfl_sync(0xFFFFFFFFu, (uint32_t)v, src);
  return ((uint64_t)hi << 32) | lo;
}

// Sorted-descending per-warp candidate list in shared memory (kp <= 128 entries).
// Insertions are rare after warm-up (expected kp*(1+ln(n/kp)) per warp), so a
// cooperative shift is cheaper overall than keeping the list in registers.
__device__ __forceinline__ void warp_list_insert(uint64_t* list, int kp, uint64_t key, int lane,
                                                 uint64_t& thresh) {
  int cnt = 0;
  uint64_t tmp[RAG_MAX_CANDIDATES / RAG_WARP];
#pragma unroll
  for (int t = 0; t < RAG_MAX_CANDIDATES / RAG_WARP; t++) {
    int j = lane + t * RAG_WARP;
    tmp[t] = j < kp ? list[j] : 0ull;
    cnt += (j < kp && tmp[t] > key) ? 1 : 0;
  }
  int pos = __reduce_add_sync(0xFFFFFFFFu, cnt);
  __syncwarp();
#pragma unroll
  for (int t = 0; t < RAG_MAX_CANDIDATES / RAG_WARP; t++) {
    int j = lane + t * RAG_WARP;
    if (j >= pos && j + 1 < kp) list[j + 1] = tmp[t];
  }
  if (lane == 0) list[pos] = key;
  __syncwarp();
  thresh = list[kp - 1];
}

// Merge nlists (<= 32) sorted-descending lists of kp keys into the kp best, by one warp.
// lists[l*stride + i]; result written to out[0..kp) (global or shared).
__device__ __forceinline__ void warp_merge_lists(const uint64_t* lists, int nlists, int stride, int kp,
                                                 uint64_t* out, int lane) {
  int head = 0;
  uint64_t cur = lane < nlists ? lists[lane * stride] : 0ull;
  for (int r = 0; r < kp; r++) {
    uint64_t m = warp_max_u64(cur);
    if (lane == 0) out[r] = m;
    if (m != 0ull && cur == m) {  // keys are unique (row id is part of the key)
      head++;
      cur = head < kp ? lists[lane * stride + head] : 0ull;
    }
  }
}
#endif  // __CUDACC__

// ---------------------------------------------------------------------------------
// host-side handle
// ---------------------------------------------------------------------------------
struct rag_comm;

// device buffers of one query batch in flight. The handle owns two: `main` (the caller's
// batch) and `esc` (the compacted sub-batch of queries escalated to a stronger path).
struct rag_batch {
  float* d_q = nullptr;             size_t c_q = 0;        // [B][ld] fp32, zero padded
  __nv_bfloat16* d_qb = nullptr;    size_t c_qb = 0;       // [Bpad][ld] 16-bit tensor path operand: fp16( q / ||q|| )
  float* d_rho_q = nullptr;         size_t c_rho_q = 0;    // [B] ||q - operand(q)|| / ||q|| (rigorous certification)
  uint8_t* d_in = nullptr;          size_t c_in = 0;       // small per-batch inputs, carved per call:
  uint8_t* h_in = nullptr;          size_t c_hin = 0;      //   pinned mirror of d_in
  uint64_t* d_kw = nullptr;                                //   [B][kw_stride] keyword keys (in d_in)
  uint32_t* d_kwc = nullptr;                               //   [B] keyword counts          (in d_in)
  uint32_t* d_sel = nullptr;        size_t c_sel = 0;      // [n] queries picked for escalation
  uint64_t* d_partial = nullptr;    size_t c_partial = 0;  // [B][parts][kp]
  uint64_t* d_cand = nullptr;       size_t c_cand = 0;     // [B][RAG_MAX_CANDIDATES]
  rag_rec* d_local = nullptr;       size_t c_local = 0;    // [B][k] this rank's exact top-k
  rag_rec* d_gather = nullptr;      size_t c_gather = 0;   // [G][B][k]
  uint32_t* d_local_cnt = nullptr;  size_t c_lcnt = 0;     // [B]
  double* d_k4s = nullptr;          size_t c_k4s = 0;      // [B][2*128+2] exact sums of the small-batch K3+K4 kernel
  unsigned int* d_ticket = nullptr; size_t c_ticket = 0;   // [B] arrival tickets (zero between launches)
  // fused outputs: one device block + pinned host mirror, carved per call (out_layout in api.cu)
  uint8_t* d_out = nullptr;         size_t c_out = 0;
  uint8_t* h_out = nullptr;         size_t c_hout = 0;
  float* h_q = nullptr;             size_t c_hq = 0;       // graph replay: pinned [queries | keyword keys | counts] at a fixed address
  uint8_t* d_gin = nullptr;         size_t c_gin = 0;      //   ... and its device block (ONE H2D node)
  uint64_t* d_out_keys = nullptr;   double* d_out_scores = nullptr;
  uint8_t* d_out_src = nullptr;     uint8_t* d_out_ct = nullptr;
  uint32_t* d_out_cnt = nullptr;    uint8_t* d_out_rrf = nullptr;
  uint64_t* d_vec_ids = nullptr;    double* d_vec_scores = nullptr;
  uint32_t* d_vec_cnt = nullptr;    uint8_t* d_cert = nullptr;
  double* d_aux0 = nullptr;         double* d_aux1 = nullptr;  // memory path: relevance, freshness
  uint32_t staged_B = 0, staged_kw_stride = 0;  // staged pool
  uint32_t win_first = 0, win_count = 0;          // window of the pool the staged runs use
};

struct rag_prof_span { cudaEvent_t a, b; int cls; };

struct rag_index {
  // Every public entry point that takes the handle holds this for its whole duration (RAG_LOCK): concurrent callers —
  // libuv pool threads behind an N-API addon, the micro-batcher's worker next to direct calls — are serialised by the
  // library instead of racing on the batch buffers. Recursive: entry points call each other (load_cache → upload).
  std::recursive_mutex mu;
  rag_index_desc desc;
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  uint64_t rows = 0;
  uint32_t dim = 0, ld = 0;
  void* corpus = nullptr;
  __nv_bfloat16* shadow = nullptr;  // 16-bit tensor-path operand: == corpus for a bf16 index, else the bf16 or fp16 copy
  bool shadow_f16 = false;          // the shadow holds fp16( x / ||x|| ) (RAG_INDEX_F16_SHADOW), not bf16( x )
  float* inv_norm = nullptr;
  uint8_t* ctype = nullptr;
  double* conf = nullptr;
  int32_t* access = nullptr;
  int64_t* last_ms = nullptr;
  uint64_t* row_keys = nullptr;
  uint64_t aux_rows = 0;            // shadow / inv_norm are valid for rows [0, aux_rows)
  // rigorous certification of the tensor path: max over rows of ||x - operand(x)|| / ||x|| (float bits on the
  // device, raised by aux_build with atomicMax; 0 for a bf16 corpus) and its host mirror
  uint32_t* d_rho_x = nullptr;
  float rho_x = 0.f;
  bool rho_x_stale = false;

  rag_batch main, esc;
  rag_batch* cur = nullptr;         // the batch the launchers operate on

  uint64_t launches = 0;
  unsigned long long* d_counters = nullptr;  // [2] certified / total queries that went through K5 (rag_certified_totals)
  rag_comm* comm = nullptr;
  int nranks = 1, rank = 0;

  // per-kernel device timing (rag_profile_*)
  bool prof_on = false;
  rag_prof_span* prof_spans = nullptr;
  uint32_t prof_cap = 0, prof_used = 0;
  float prof_ms[RAG_PROF_CLASSES] = {0};
  uint32_t prof_cnt[RAG_PROF_CLASSES] = {0};

  // tensor path (K2) state of the CTA-pair kernel
  void* k2p_state = nullptr;
  // small-batch latency path: the whole call (H2D, K1, K3+K4+K5, D2H) captured once as a CUDA graph and replayed
  struct rag_graph* graph = nullptr;
};

// error plumbing (api.cu)
int rag_set_error(int code, const char* fmt, ...);
#define RAG_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return rag_set_error(RAG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                             \
  } while (0)
#define RAG_LOCK(idx) std::lock_guard<std::recursive_mutex> _rag_lock((idx)->mu)
#define RAG_CHECK(expr)             \
  do {                              \
    int _r = (expr);                \
    if (_r != RAG_OK) return _r;    \
  } while (0)

// profiling spans (api.cu): bracket one kernel class on the library stream
int rag_prof_begin(rag_index* idx, int cls);
void rag_prof_end(rag_index* idx, int token);
struct rag_prof_scope {
  rag_index* idx; int token;
  rag_prof_scope(rag_index* i, int cls) : idx(i), token(i->prof_on ? rag_prof_begin(i, cls) : -1) {}
  ~rag_prof_scope() { if (token >= 0) rag_prof_end(idx, token); }
};

// ---------------------------------------------------------------------------------
// kernel launchers (one .cu each). All operate on idx->cur and idx->stream.
// ---------------------------------------------------------------------------------
// K1 — stream path: fused cosine GEMV + per-warp top-K' (k1_stream.cu)
int k1_plan(const rag_index* idx, uint32_t B, uint32_t kp, uint32_t* parts);
// on_shadow: stream the fp16 shadow of the normalised rows instead of the fp32 corpus (RAG_PATH_SHADOW_STREAM)
int k1_launch(rag_index* idx, uint32_t B, uint32_t kp, uint32_t parts, bool on_shadow = false);
// K1x — exact path: fp64 reference-order scan of every row (k1x_exact.cu)
int k1x_plan(const rag_index* idx, uint32_t B, uint32_t kp, uint32_t* parts);
int k1x_launch(rag_index* idx, uint32_t B, uint32_t kp, uint32_t parts);
// K2 — tensor path: tcgen05 GEMM on CTA pairs (cta_group::2) + fused top-K' epilogue (k2_pair.cu).
// Operand: the bf16 corpus / bf16 shadow, or the fp32 rows read as tf32 when an fp32 index has no shadow.
int k2_available(const rag_index* idx);
int k2_plan(rag_index* idx, uint32_t B, uint32_t kp, uint32_t* parts);
int k2_launch(rag_index* idx, uint32_t B, uint32_t kp, uint32_t parts, float floor = -INFINITY);  // floor: rows scoring <= it are never candidates
void k2_destroy(rag_index* idx);
void k2_set_debug(rag_index* idx, float* d_scores);  // diagnostics: dump the scaled score matrix of the next launch
// diagnostics of the batch-1 latency path (RAGERA_SMALL_PROF=1): enable >= 0 sets / resets, dump != null prints averages
void k1_small_prof(int enable, FILE* dump);
void k34_small_prof(int enable, FILE* dump);
// K3 — merge partial lists → K' candidates per query (k3_merge.cu)
int k3_launch(rag_index* idx, uint32_t B, uint32_t kp, uint32_t parts);
// K4 — exact fp64 rescoring in reference order + local top-k + certification (k4_rescore.cu)
// eps_q (may be null): per-query addend to the selection-error bound, eps[b] = eps + eps_q[b] * eps_q_mul
// What K4's certification needs to know about the selecting pass: its error bound (eps, + eps_q[b] * eps_q_mul per query),
// `floor` = the selection score at or below which the selecting kernel dropped rows outright (-inf: it dropped nothing),
// `min_score` = the cosine below which the caller filters results anyway (hybridSearch's minVectorScore, MemoryStore's
// minRelevance; -inf: a plain top-k). All in cosine units.
struct rag_eps { double eps; const float* eps_q; double eps_q_mul; double floor; double min_score; };
int k4_launch(rag_index* idx, uint32_t B, uint32_t kp, uint32_t k, rag_eps eps, int key_has_qnorm,
              int64_t now_ms, double decay, double bonus);
// K3+K4 fused, latency variant for small batches (k4_rescore.cu)
bool k34_small_ok(const rag_index* idx, uint32_t B, uint32_t kp, uint32_t parts);
int k34_small_launch(rag_index* idx, uint32_t B, uint32_t kp, uint32_t parts, uint32_t k, rag_eps eps, int key_has_qnorm,
                     int64_t now_ms, double decay, double bonus, const struct rag_fuse_args* fuse /* non-null: run K5 in place */);
bool k34_small_fuses_exchange(const rag_index* idx);  // sharded: the in-place K5 also runs the peer-to-peer exchange
// K5 — cross-rank merge, min-cosine filter, RRF / freshness fusion, memory blend (k5_fuse.cu)
struct rag_fuse_args {
  uint32_t B, k, nranks;
  uint32_t kw_stride, out_cap;
  double min_score;
  rag_rrf_config rrf;
  uint32_t fresh_limit;
  double fresh_weight;
  int mode;  // 0 = hybrid (RRF or vector-only), 1 = memory retrieve (store.ts blend), 2 = vector only
  uint32_t mem_limit;
  double mem_min_relevance;
};
int k5_launch(rag_index* idx, const rag_fuse_args* a);
int k5_rrf_only_launch(rag_index* idx, uint32_t B, const rag_rrf_config* cfg, const uint64_t* d_vec_keys,
                       const uint8_t* d_vec_ct, const uint32_t* d_vec_cnt, uint32_t vec_stride,
                       const uint64_t* d_kw, const uint32_t* d_kwc, uint32_t kw_stride, uint32_t out_cap);
int k5_freshness_launch(rag_index* idx, uint64_t n, const double* d_conf, const int32_t* d_acc,
                        const int64_t* d_last, int64_t now_ms, double decay, double bonus, double* d_out);
// generator / maintenance kernels (gen.cu)
int gen_corpus_launch(rag_index* idx, const rag_gen_desc* g, uint64_t nrows);
int gen_queries_launch(rag_index* idx, const rag_gen_desc* g, uint64_t b0, uint32_t B, float* d_out);
int gen_meta_launch(rag_index* idx, const rag_gen_desc* g, uint64_t nrows);
int aux_build_launch(rag_index* idx, uint64_t row0, uint64_t nrows);  // shadow + inv_norm
int q_operand_launch(rag_index* idx, uint32_t B, uint32_t Bpad, bool tf32);  // bf16 cast (or none: tf32) + rho_q
int gather_batch_launch(rag_index* idx, const rag_batch* src, rag_batch* dst, uint32_t n, uint32_t kw_stride);
int iota_u64_launch(rag_index* idx, uint64_t* d, uint64_t n, uint64_t base);
// comm (comm.cu): exchange of the ranks' local top-k records.
// Default: peer-to-peer mailboxes over NVLink, fused into K5 — each rank's K5 stores its [B][k] records straight
// into every peer's mailbox (CUDA IPC mapping), raises a per-query flag, waits for the peers' flags and merges;
// no collective call on the hot path. Fallback (RAGERA_COMM=nccl, or no peer access): ncclAllGather of
// cur->d_local → cur->d_gather before K5.
struct rag_p2p_view {
  unsigned char* base[8];  // rank g's mailbox as mapped into this process (base[rank] is the local allocation)
  uint32_t nranks, rank;   // nranks <= 1: no exchange (single GPU, or the NCCL fallback already gathered)
  uint32_t step;           // exchange number: its parity selects the mailbox half
  uint32_t flag;           // the value raised in / awaited from the flag words: step << 12 | hash(B, k) — ranks
                           // that disagree on the call sequence or its shape time out instead of merging garbage
  uint64_t half_bytes;     // bytes of one parity half: [records | flags]
  uint64_t flags_off;      // offset of the flag words inside a half: [src rank][flag_stride] u32
  uint32_t flag_stride;    // flag slots per source rank (>= the batch)
  long long timeout_cycles;  // a peer that has not arrived by then: give up (status word, no trap)
  uint32_t* status;        // host-mapped word: set to 1 by a query whose exchange timed out
};
// The exchange counter: 20 bits (it travels in the flag word above the 12-bit shape hash), never 0 (the value of a flag word
// nobody has written), and — what the two parity halves of the mailbox rest on — CONSECUTIVE exchanges always differ in
// parity, across the wrap too: 0xFFFFF (odd) is followed by 2, not by 1. (Wrapping onto 1 would run two exchanges in a row on
// the same half: a fast rank could refill a mailbox its slower peer is still merging from, once every 2^20 searches.)
static inline uint32_t rag_p2p_next_step(uint32_t step) { return step >= 0xFFFFFu ? 2u : step + 1u; }
int comm_p2p_ensure(rag_index* idx, uint32_t B, uint32_t k);            // sizes / checks the mailboxes for this shape
bool comm_uses_p2p(const rag_index* idx);
int comm_p2p_next(rag_index* idx, uint32_t B, uint32_t k, rag_p2p_view* v);  // view of the next exchange
int comm_check_status(rag_index* idx);  // after a stream sync: RAG_ERR_TIMEOUT if an exchange gave up
int comm_allgather_local(rag_index* idx, uint32_t B, uint32_t k);
