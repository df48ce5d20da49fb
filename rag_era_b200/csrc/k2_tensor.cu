// k2_tensor.cu — K2, single-CTA variant (cta_group::1) + the dispatch between the K2 variants.
// The default tensor-path kernel is the CTA-pair kernel of k2_pair.cu (cta_group::2, double-buffered
// TMEM, tf32 mode); this predecessor stays selectable with RAGERA_K2_IMPL=1 and is kept under test.
//
// K2: batched cosine scoring on the 5th-gen tensor cores (tcgen05) with the top-K' selection
// fused into the epilogue, so the [B x N] score matrix never leaves the SM.
//
// Replaces the same reference loop as K1 (getTopKEmbeddings, reached from
// src/lib/hybrid-search.ts:223-224) when many queries are in flight: only then is the path a
// dense contraction  S[q][r] = <Q[q,:], X[r,:]> * inv_norm[r].
//
// Operands: Qb [Bpad][ld] bf16 (queries, rounded from fp32 per batch) and the bf16 corpus
// (or the bf16 shadow of an fp32 corpus) [rows][ld], both K-major, moved by TMA
// (cp.async.bulk.tensor, 64-byte swizzle) into a ring of shared-memory stages.
// MMA: tcgen05.mma.cta_group::1.kind::f16, M = 128 queries (TMEM lanes) x N = 256 corpus rows
// (TMEM columns), K = 16 per instruction; fp32 accumulators live in TMEM.
//
// CTA (x, y): query group y (up to 2 blocks of 128 queries = two 256-column accumulators =
// all 512 TMEM columns) x corpus tiles x, x+gridDim.x, ... Every stage carries both query
// blocks' k-slices and ONE corpus k-slice (used by two MMAs), which is what keeps L2->smem
// traffic at (128+128+256)*64 B per 8 MMAs. CTAs of different y walk the same tiles at the same
// time, so the corpus is read from HBM once and served to the other groups from L2.
//
// Warp roles (384 threads, 1 CTA/SM): w0 TMA producer · w1 MMA issuer (one elected lane) ·
// w2 TMEM allocator · w3 inverse-norm tile loader · w4..w11 epilogue (accumulator = (w-4)/4,
// TMEM lane quarter = w%4; one LANE per query).
// Epilogue: tcgen05.ld 16 columns at a time, scale by 1/||x|| (broadcast from smem), compare
// with the lane's threshold (the K'-th best at its last prune); the rare survivors are appended
// to the query's 64-slot buffer in shared memory (same packed keys as K1) and the warp prunes
// a buffer back to its K' best when it fills. The K' best per (CTA, query) go to
// partial[B][parts][K'] at the end; K3 merges, K4 rescoring in fp64 decides ids and order.
//
// Roofline: tensor pipe for B >= ~64 (algorithmic flops 2*rows*ld*B), HBM below that
// (rows*ld*2 bytes streamed once).
#include "common.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

namespace {

constexpr int BK = 32;                       // k elements per stage: 64 B of bf16 = one SWIZZLE_64B row
constexpr int TILE_N = 256;                  // corpus rows per tile (UMMA N)
constexpr int TILE_M = 128;                  // queries per block (UMMA M = TMEM lanes)
constexpr int MAX_MB = 2;                    // query blocks per CTA
constexpr int QPC = TILE_M * MAX_MB;         // queries per CTA
constexpr int A_BYTES = TILE_M * BK * 2;     // 8 KB
constexpr int B_BYTES = TILE_N * BK * 2;     // 16 KB
constexpr int STAGE_BYTES = MAX_MB * A_BYTES + B_BYTES;  // 32 KB
constexpr int K2_THREADS = 384;
constexpr int MAX_STAGES = 6;
constexpr int TMEM_COLS = 512;
constexpr int CAP = 64;                      // per-query candidate buffer in smem (K' kept + room for appends)
constexpr int K2_MAX_KP = CAP - 16;          // a 16-column chunk must always fit after a prune

using namespace tc;
constexpr uint32_t kIdesc = idesc_bf16(TILE_M, TILE_N);

struct k2_params {
  uint32_t n_rows, ld, B, kp, parts, stages, n_tiles;
  const float* inv_norm;
  uint64_t* partial;
  float* dbg_scores;  // optional [B][n_rows] raw scaled scores (diagnostics; small problems only)
  uint32_t mode;      // diagnostics (RAGERA_K2_MODE): 0 normal · 1 epilogue reads TMEM but selects nothing · 2 epilogue skips TMEM
  unsigned long long* cyc;  // diagnostics (RAGERA_K2_PROF): [grid][12 warps][8] cycle counters
};
__device__ __forceinline__ long long clk() { return clock64(); }

// Keep the K' largest of the `cnt` (<= 64) keys in one query's buffer, sorted descending, by one warp.
// Logical entry i lives in physical slot (i + rot) & 63 (rot = owning lane: de-conflicts the
// banks when all lanes of a warp append at once). Rank by counting: keys are unique (the row is
// part of the key). Returns the K'-th key (the new threshold) or 0 if fewer than K' remain.
__device__ __forceinline__ uint64_t warp_prune(uint64_t* buf, int cnt, int kp, int rot, int lane) {
  const uint64_t k0 = lane < cnt ? buf[(lane + rot) & (CAP - 1)] : 0ull;
  const uint64_t k1 = lane + 32 < cnt ? buf[(lane + 32 + rot) & (CAP - 1)] : 0ull;
  int r0 = 0, r1 = 0;
  // all-pairs compare through shuffles (no shared-memory round trips); empty slots (0) rank last
#pragma unroll
  for (int j = 0; j < 32; j++) {
    const uint64_t a = shfl_u64(k0, j), b = shfl_u64(k1, j);
    r0 += (a > k0 ? 1 : 0) + (b > k0 ? 1 : 0);
    r1 += (a > k1 ? 1 : 0) + (b > k1 ? 1 : 0);
  }
  __syncwarp();
  buf[lane] = 0ull;
  buf[lane + 32] = 0ull;
  __syncwarp();
  if (k0 != 0ull && r0 < kp) buf[(r0 + rot) & (CAP - 1)] = k0;
  if (k1 != 0ull && r1 < kp) buf[(r1 + rot) & (CAP - 1)] = k1;
  __syncwarp();
  return cnt >= kp ? buf[(kp - 1 + rot) & (CAP - 1)] : 0ull;
}

__global__ void __launch_bounds__(K2_THREADS, 1)
k2_tensor_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, const k2_params P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // layout: [stages][A0 | A1 | B] · lists [QPC][CAP] u64 · inv [2][256] f32 · barriers · tmem ptr
  unsigned char* stage_base = smem;
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem + (size_t)P.stages * STAGE_BYTES);
  float* s_inv = reinterpret_cast<float*>(lists + (size_t)QPC * CAP);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_inv + 2 * TILE_N);
  uint64_t* full = bars;                    // [stages]
  uint64_t* empty = bars + MAX_STAGES;      // [stages]
  uint64_t* tmem_full = bars + 2 * MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint64_t* inv_full = tmem_empty + 1;      // [2]
  uint64_t* inv_empty = inv_full + 2;       // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(inv_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t q0 = blockIdx.y * QPC;
  const uint32_t nmb = min((uint32_t)MAX_MB, (P.B - q0 + TILE_M - 1) / TILE_M);
  const uint32_t nkb = P.ld / BK;
  const uint32_t n_epi_warps = nmb * 4;

  for (uint32_t i = threadIdx.x; i < QPC * CAP; i += K2_THREADS) lists[i] = 0ull;
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < P.stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, n_epi_warps);
    for (int s = 0; s < 2; s++) { mbar_init(&inv_full[s], 1); mbar_init(&inv_empty[s], n_epi_warps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      long long w_empty = 0, t_begin = clk();
      for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        for (uint32_t kb = 0; kb < nkb; kb++) {
          const long long t0 = clk();
          mbar_wait(&empty[stage], phase ^ 1);
          w_empty += clk() - t0;
          unsigned char* sa = stage_base + (size_t)stage * STAGE_BYTES;
          mbar_expect_tx(&full[stage], nmb * A_BYTES + B_BYTES);
          for (uint32_t mb = 0; mb < nmb; mb++)
            tma_load_2d(sa + mb * A_BYTES, &map_q, &full[stage], (int)(kb * BK), (int)(q0 + mb * TILE_M));
          tma_load_2d(sa + MAX_MB * A_BYTES, &map_x, &full[stage], (int)(kb * BK), (int)(tile * TILE_N));
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
      }
      if (P.cyc) {
        unsigned long long* c = P.cyc + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 12 + warp) * 8;
        c[0] = (unsigned long long)(clk() - t_begin); c[1] = (unsigned long long)w_empty;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, it = 0;
      long long w_tmem = 0, w_full = 0, t_begin = clk();
      for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, it++) {
        long long t0 = clk();
        mbar_wait(tmem_empty, (it & 1) ^ 1);  // the epilogue has drained both accumulators
        w_tmem += clk() - t0;
        tcgen05_fence_after();
        for (uint32_t kb = 0; kb < nkb; kb++) {
          t0 = clk();
          mbar_wait(&full[stage], phase);
          w_full += clk() - t0;
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(stage_base + (size_t)stage * STAGE_BYTES);
          const uint32_t sb = sa + MAX_MB * A_BYTES;
#pragma unroll
          for (uint32_t k16 = 0; k16 < BK / 16; k16++) {
            const uint64_t bdesc = umma_desc_sw64(sb + k16 * 32);
            for (uint32_t mb = 0; mb < nmb; mb++)
              tcgen05_mma_f16(tmem_base + mb * TILE_N, umma_desc_sw64(sa + mb * A_BYTES + k16 * 32), bdesc, kIdesc,
                              (kb | k16) != 0 ? 1u : 0u);
          }
          tcgen05_commit(&empty[stage]);  // frees the smem stage once these MMAs have read it
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit(tmem_full);  // accumulators of this tile are complete
      }
      if (P.cyc) {
        unsigned long long* c = P.cyc + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 12 + warp) * 8;
        c[0] = (unsigned long long)(clk() - t_begin); c[1] = (unsigned long long)w_tmem; c[2] = (unsigned long long)w_full;
      }
    }
  } else if (warp == 3) {
    // ===== inverse-norm tile loader (NaN marks rows that must never be selected) =====
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, it++) {
      const uint32_t buf = it & 1;
      mbar_wait(&inv_empty[buf], ((it >> 1) & 1) ^ 1);
      const uint32_t r0 = tile * TILE_N + lane * 8;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const uint32_t r = r0 + i;
        const float x = r < P.n_rows ? __ldg(P.inv_norm + r) : 0.f;
        v[i] = x > 0.f ? x : __int_as_float(0x7FC00000);
      }
      float4* dst = reinterpret_cast<float4*>(s_inv + buf * TILE_N + lane * 8);
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
      __syncwarp();
      if (lane == 0) mbar_arrive(&inv_full[buf]);
    }
  } else if (warp >= 4 && (uint32_t)(warp - 4) < n_epi_warps) {
    // ===== epilogue: one lane per query =====
    // Selection = threshold filter + append + occasional prune: a score above the lane's threshold
    // (the K'-th best at its last prune) is appended to the query's 64-slot buffer; when fewer
    // than 16 slots are free the warp prunes that buffer to its K' best and raises the threshold.
    // Each prune roughly doubles the rows seen, so a query is pruned O(log(rows/K')) times.
    const uint32_t e = warp - 4, mb = e >> 2, quarter = warp & 3;
    const uint32_t ql = mb * TILE_M + quarter * 32 + lane;  // query within the CTA
    const uint32_t qg = q0 + ql;                             // query in the batch
    const bool live = qg < P.B;
    uint64_t* warp_bufs = lists + (size_t)(mb * TILE_M + quarter * 32) * CAP;
    uint64_t* mybuf = warp_bufs + (size_t)lane * CAP;
    float thr = live ? -INFINITY : INFINITY;
    int cnt = 0;
    const int kp = (int)P.kp;
    long long c_sel = 0, c_prune = 0, c_wait = 0, c_ld = 0, n_app = 0, n_prune = 0;
    const long long t_begin = clk();
    const uint32_t taddr0 = tmem_base + ((quarter * 32u) << 16) + mb * TILE_N;

    auto process16 = [&](const uint32_t (&v)[16], const float* inv, uint32_t row) {
      float s[16];
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 w = *reinterpret_cast<const float4*>(inv + i);
        s[i] = __uint_as_float(v[i]) * w.x;
        s[i + 1] = __uint_as_float(v[i + 1]) * w.y;
        s[i + 2] = __uint_as_float(v[i + 2]) * w.z;
        s[i + 3] = __uint_as_float(v[i + 3]) * w.w;
      }
      if (P.dbg_scores && live) {
#pragma unroll
        for (int i = 0; i < 16; i++)
          if (row + i < P.n_rows) P.dbg_scores[(size_t)qg * P.n_rows + row + i] = s[i];
      }
      if (P.mode != 0) return;
      const long long ts = clk();
      // rows arrive in increasing order, so a later equal score can never displace an earlier one:
      // the strict float compare against the K'-th best is exact. NaN (padding / zero-norm rows)
      // never passes. Survivors are rare (O(K' log rows) per query), so they are handled four
      // columns at a time: one vote per group, then each lane with a hit appends it.
#pragma unroll
      for (int g = 0; g < 4; g++) {
        unsigned pm = (s[4 * g] > thr ? 1u : 0u) | (s[4 * g + 1] > thr ? 2u : 0u) | (s[4 * g + 2] > thr ? 4u : 0u) |
                      (s[4 * g + 3] > thr ? 8u : 0u);
        while (__any_sync(0xFFFFFFFFu, pm != 0u)) {
          if (pm != 0u) {
            const int i = __ffs(pm) - 1;
            pm &= pm - 1;
            const float val = i == 0 ? s[4 * g] : (i == 1 ? s[4 * g + 1] : (i == 2 ? s[4 * g + 2] : s[4 * g + 3]));
            mybuf[(cnt + lane) & (CAP - 1)] = rag_pack_key(val, row + 4 * g + i);
            cnt++;
            n_app++;
          }
        }
      }
      {
        const long long tp = clk();
        unsigned need = __ballot_sync(0xFFFFFFFFu, cnt > CAP - 16);
        while (need) {
          n_prune++;
          const int src = __ffs(need) - 1;
          need &= need - 1;
          const int c = __shfl_sync(0xFFFFFFFFu, cnt, src);
          const uint64_t t = warp_prune(warp_bufs + (size_t)src * CAP, c, kp, src, lane);
          if (lane == src) {
            cnt = min(c, kp);
            if (t != 0ull) thr = rag_key_score(t);
          }
        }
        c_prune += clk() - tp;
      }
      c_sel += clk() - ts;
    };

    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, it++) {
      const uint32_t buf = it & 1;
      const long long tw = clk();
      mbar_wait(&inv_full[buf], (it >> 1) & 1);
      mbar_wait(tmem_full, it & 1);
      c_wait += clk() - tw;
      tcgen05_fence_after();
      const float* inv = s_inv + buf * TILE_N;
      const uint32_t row0 = tile * TILE_N;
      if (P.mode != 2) {
#pragma unroll 1
        for (uint32_t c0 = 0; c0 < TILE_N; c0 += 32) {
          uint32_t va[16], vb[16];
          const long long tl = clk();
          tmem_ld16(taddr0 + c0, va);
          tmem_ld16(taddr0 + c0 + 16, vb);
          tmem_ld_wait();
          c_ld += clk() - tl;
          process16(va, inv + c0, row0 + c0);
          process16(vb, inv + c0 + 16, row0 + c0 + 16);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(tmem_empty);
        mbar_arrive(&inv_empty[buf]);
      }
    }
    if (P.cyc && lane == 0) {
      unsigned long long* c = P.cyc + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 12 + warp) * 8;
      c[0] = (unsigned long long)(clk() - t_begin); c[1] = (unsigned long long)c_wait; c[2] = (unsigned long long)c_ld;
      c[3] = (unsigned long long)c_sel; c[4] = (unsigned long long)c_prune; c[5] = (unsigned long long)n_app;
      c[6] = (unsigned long long)n_prune;
    }
    // final prune of every query of this warp, then publish the K' best (sorted) of this CTA
    for (int src = 0; src < 32; src++) {
      const int c = __shfl_sync(0xFFFFFFFFu, cnt, src);
      warp_prune(warp_bufs + (size_t)src * CAP, c, kp, src, lane);
    }
    if (live) {
      uint64_t* out = P.partial + ((size_t)qg * P.parts + blockIdx.x) * P.kp;
      for (int j = 0; j < kp; j++) out[j] = mybuf[(j + lane) & (CAP - 1)];
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// ---- host side ------------------------------------------------------------------------------
struct k2_state {
  PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  int max_smem = 0;
  bool attr_set = false;
  float* dbg = nullptr;  // set by rag_debug_tensor_scores for one launch
  uint32_t mode = 0;     // RAGERA_K2_MODE (diagnostics)
  bool prof = false;     // RAGERA_K2_PROF (diagnostics): print per-role cycle counters after each launch
  unsigned long long* d_cyc = nullptr;
};

size_t k2_smem_bytes(uint32_t stages, uint32_t kp) {
  (void)kp;
  return (size_t)stages * STAGE_BYTES + (size_t)QPC * CAP * 8 + 2 * TILE_N * 4 + (2 * MAX_STAGES + 6) * 8 + 16;
}

int k2_init(rag_index* idx) {
  if (idx->k2_state) return RAG_OK;
  k2_state* st = new k2_state();
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
    delete st;
    cudaGetLastError();
    return rag_set_error(RAG_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  }
  st->encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  if (const char* m = getenv("RAGERA_K2_MODE")) st->mode = (uint32_t)atoi(m);
  if (const char* m = getenv("RAGERA_K2_PROF")) st->prof = atoi(m) != 0;
  cudaDeviceGetAttribute(&st->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, idx->device);
  idx->k2_state = st;
  return RAG_OK;
}

int make_map(k2_state* st, CUtensorMap* m, const void* base, uint64_t rows, uint32_t ld, uint32_t box_rows) {
  cuuint64_t dims[2] = {ld, rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = st->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return rag_set_error(RAG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return RAG_OK;
}

uint32_t pick_stages(const k2_state* st, uint32_t kp) {
  uint32_t s = MAX_STAGES;
  while (s > 0 && k2_smem_bytes(s, kp) > (size_t)st->max_smem) s--;
  return s;
}

}  // namespace

namespace {
// RAGERA_K2_IMPL=1 selects this file's single-CTA kernel (cta_group::1); the default is the CTA-pair
// kernel of k2_pair.cu (cta_group::2, epilogue overlapped with the next tile's MMAs)
int k2_impl() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RAGERA_K2_IMPL");
    v = e ? atoi(e) : 2;
  }
  return v;
}
}  // namespace

// bf16 operand (bf16 corpus / bf16 shadow) on either kernel, or the fp32 corpus as tf32 on the pair kernel
int k2_available(const rag_index* idx) {
  if (!idx->inv_norm) return 0;
  if (idx->shadow) return 1;
  return idx->desc.dtype == RAG_F32 && k2_impl() != 1;
}

static int k2s_plan(rag_index* idx, uint32_t B, uint32_t kp, uint32_t* parts) {
  RAG_CHECK(k2_init(idx));
  k2_state* st = (k2_state*)idx->k2_state;
  if (kp > K2_MAX_KP) return rag_set_error(RAG_ERR_UNSUPPORTED, "tensor path keeps at most %d candidates per query (K'=%u)", K2_MAX_KP, kp);
  if (pick_stages(st, kp) < 2) return rag_set_error(RAG_ERR_UNSUPPORTED, "tensor path: not enough shared memory for K'=%u", kp);
  const uint32_t groups = (B + QPC - 1) / QPC;
  if (groups > (uint32_t)idx->sm_count)
    return rag_set_error(RAG_ERR_UNSUPPORTED, "tensor path: batch %u exceeds %d queries per launch", B, idx->sm_count * QPC);
  const uint64_t n_tiles = (idx->rows + TILE_N - 1) / TILE_N;
  uint64_t slabs = (uint64_t)idx->sm_count / groups;
  if (slabs > n_tiles) slabs = n_tiles;
  if (slabs < 1) slabs = 1;
  *parts = (uint32_t)slabs;
  return RAG_OK;
}

static int k2s_launch(rag_index* idx, uint32_t B, uint32_t kp, uint32_t parts) {
  if (idx->rows == 0) return rag_set_error(RAG_ERR_STATE, "search on an empty index");
  if (idx->rows >= 0xFFFFFF00ull) return rag_set_error(RAG_ERR_UNSUPPORTED, "more than 2^32-257 rows per shard");
  RAG_CHECK(k2_init(idx));
  k2_state* st = (k2_state*)idx->k2_state;
  rag_batch* bt = idx->cur;
  const uint32_t Bpad = (B + TILE_M - 1) / TILE_M * TILE_M;
  {
    // queries → bf16 [Bpad][ld], zero padded (grow through a temporary so the size is tracked)
    size_t need = (size_t)Bpad * idx->ld * 2;
    if (need > bt->c_qb || !bt->d_qb) {
      if (bt->d_qb) RAG_CUDA(cudaFree(bt->d_qb));
      bt->d_qb = nullptr; bt->c_qb = 0;
      RAG_CUDA(cudaMalloc((void**)&bt->d_qb, need));
      bt->c_qb = need;
    }
  }
  RAG_CHECK(q_to_bf16_launch(idx, B, Bpad));

  rag_prof_scope ps(idx, RAG_PROF_TENSOR);
  CUtensorMap map_q, map_x;
  RAG_CHECK(make_map(st, &map_q, bt->d_qb, Bpad, idx->ld, TILE_M));
  RAG_CHECK(make_map(st, &map_x, idx->shadow, idx->rows, idx->ld, TILE_N));
  k2_params P;
  P.n_rows = (uint32_t)idx->rows;
  P.ld = idx->ld;
  P.B = B;
  P.kp = kp;
  P.parts = parts;
  P.stages = pick_stages(st, kp);
  P.n_tiles = (uint32_t)((idx->rows + TILE_N - 1) / TILE_N);
  P.inv_norm = idx->inv_norm;
  P.partial = bt->d_partial;
  P.dbg_scores = st->dbg;
  P.mode = st->mode;
  P.cyc = nullptr;
  const uint32_t groups_ = (B + QPC - 1) / QPC;
  const size_t n_cyc = (size_t)parts * groups_ * 12 * 8;
  if (st->prof) {
    if (st->d_cyc) cudaFree(st->d_cyc);
    RAG_CUDA(cudaMalloc((void**)&st->d_cyc, n_cyc * 8));
    RAG_CUDA(cudaMemsetAsync(st->d_cyc, 0, n_cyc * 8, idx->stream));
    P.cyc = st->d_cyc;
  }
  const size_t smem = k2_smem_bytes(P.stages, kp);
  if (!st->attr_set) {
    RAG_CUDA(cudaFuncSetAttribute(k2_tensor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, st->max_smem));
    st->attr_set = true;
  }
  const uint32_t groups = (B + QPC - 1) / QPC;
  k2_tensor_kernel<<<dim3(parts, groups), K2_THREADS, smem, idx->stream>>>(map_q, map_x, P);
  RAG_CUDA(cudaGetLastError());
  idx->launches++;
  if (st->prof) {
    static int printed = 0;
    std::vector<unsigned long long> h(n_cyc);
    RAG_CUDA(cudaMemcpyAsync(h.data(), st->d_cyc, n_cyc * 8, cudaMemcpyDeviceToHost, idx->stream));
    RAG_CUDA(cudaStreamSynchronize(idx->stream));
    if (printed++ % 16 == 8) {
      const size_t nct = (size_t)parts * groups_;
      double acc[12][8] = {{0}};
      for (size_t c = 0; c < nct; c++)
        for (int w = 0; w < 12; w++)
          for (int j = 0; j < 8; j++) acc[w][j] += (double)h[(c * 12 + w) * 8 + j] / nct;
      fprintf(stderr, "[k2 prof] B=%u rows=%u kp=%u stages=%u ctas=%zu (avg cycles per CTA)\n", B, P.n_rows, kp, P.stages, nct);
      fprintf(stderr, "  producer: total %.0f wait_empty %.0f\n", acc[0][0], acc[0][1]);
      fprintf(stderr, "  mma     : total %.0f wait_tmem_empty %.0f wait_full %.0f\n", acc[1][0], acc[1][1], acc[1][2]);
      for (int w = 4; w < 12; w++)
        fprintf(stderr, "  epi w%-2d : total %.0f wait %.0f ld %.0f select %.0f (prune %.0f) appends(lane0) %.0f prunes %.0f\n", w,
                acc[w][0], acc[w][1], acc[w][2], acc[w][3], acc[w][4], acc[w][5], acc[w][6]);
    }
  }
  return RAG_OK;
}

int k2_plan(rag_index* idx, uint32_t B, uint32_t kp, uint32_t* parts) {
  return k2_impl() == 1 ? k2s_plan(idx, B, kp, parts) : k2p_plan(idx, B, kp, parts);
}
int k2_launch(rag_index* idx, uint32_t B, uint32_t kp, uint32_t parts) {
  return k2_impl() == 1 ? k2s_launch(idx, B, kp, parts) : k2p_launch(idx, B, kp, parts);
}

void k2_set_debug(rag_index* idx, float* d_scores) {
  if (k2_impl() != 1) { k2p_set_debug(idx, d_scores); return; }
  if (k2_init(idx) == RAG_OK) ((k2_state*)idx->k2_state)->dbg = d_scores;
}

void k2_destroy(rag_index* idx) {
  k2p_destroy(idx);
  if (idx->k2_state && ((k2_state*)idx->k2_state)->d_cyc) cudaFree(((k2_state*)idx->k2_state)->d_cyc);
  delete (k2_state*)idx->k2_state;
  idx->k2_state = nullptr;
}
