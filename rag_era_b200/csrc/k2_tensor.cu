// k2_tensor.cu — K2: batched cosine scoring on the 5th-gen tensor cores (tcgen05) with the
// top-K' selection fused into the epilogue, so the [B x N] score matrix never leaves the SM.
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
// with the lane's running threshold; the rare survivors are inserted warp-cooperatively into
// the query's sorted K' list in shared memory (same packed keys and insert as K1). Lists go
// to partial[B][parts][K'] at the end; K3 merges, K4 rescoring in fp64 decides ids and order.
//
// Roofline: tensor pipe for B >= ~64 (algorithmic flops 2*rows*ld*B), HBM below that
// (rows*ld*2 bytes streamed once).
#include "common.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

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

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major operand, 64-byte swizzle: rows of 64 B, 8-row atoms
// of 512 B (SBO), LBO unused (1). cute/arch/mma_sm100_desc.hpp: start>>4 [0,14) · LBO>>4
// [16,30) · SBO>>4 [32,46) · version=1 [46,48) · layout SWIZZLE_64B=4 [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
// instruction descriptor kind::f16: D=f32 (1<<4) · A=bf16 (1<<7) · B=bf16 (1<<10) · both K-major ·
// N>>3 at [17,23) · M>>4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);

struct k2_params {
  uint32_t n_rows, ld, B, kp, parts, stages, n_tiles;
  const float* inv_norm;
  uint64_t* partial;
  float* dbg_scores;  // optional [B][n_rows] raw scaled scores (diagnostics; small problems only)
};

__global__ void __launch_bounds__(K2_THREADS, 1)
k2_tensor_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, const k2_params P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // layout: [stages][A0 | A1 | B] · lists [QPC][kp] u64 · inv [2][256] f32 · barriers · tmem ptr
  unsigned char* stage_base = smem;
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem + (size_t)P.stages * STAGE_BYTES);
  float* s_inv = reinterpret_cast<float*>(lists + (size_t)QPC * P.kp);
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

  for (uint32_t i = threadIdx.x; i < QPC * P.kp; i += K2_THREADS) lists[i] = 0ull;
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
      for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        for (uint32_t kb = 0; kb < nkb; kb++) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* sa = stage_base + (size_t)stage * STAGE_BYTES;
          mbar_expect_tx(&full[stage], nmb * A_BYTES + B_BYTES);
          for (uint32_t mb = 0; mb < nmb; mb++)
            tma_load_2d(sa + mb * A_BYTES, &map_q, &full[stage], (int)(kb * BK), (int)(q0 + mb * TILE_M));
          tma_load_2d(sa + MAX_MB * A_BYTES, &map_x, &full[stage], (int)(kb * BK), (int)(tile * TILE_N));
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, it = 0;
      for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, it++) {
        mbar_wait(tmem_empty, (it & 1) ^ 1);  // the epilogue has drained both accumulators
        tcgen05_fence_after();
        for (uint32_t kb = 0; kb < nkb; kb++) {
          mbar_wait(&full[stage], phase);
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
    const uint32_t e = warp - 4, mb = e >> 2, quarter = warp & 3;
    const uint32_t ql = mb * TILE_M + quarter * 32 + lane;  // query within the CTA
    const uint32_t qg = q0 + ql;                             // query in the batch
    const bool live = qg < P.B;
    uint64_t* warp_lists = lists + (size_t)(mb * TILE_M + quarter * 32) * P.kp;
    float thr = live ? -INFINITY : INFINITY;
    const uint32_t taddr0 = tmem_base + ((quarter * 32u) << 16) + mb * TILE_N;
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, it++) {
      const uint32_t buf = it & 1;
      mbar_wait(&inv_full[buf], (it >> 1) & 1);
      mbar_wait(tmem_full, it & 1);
      tcgen05_fence_after();
      const float* inv = s_inv + buf * TILE_N;
      const uint32_t row0 = tile * TILE_N;
#pragma unroll 1
      for (uint32_t c0 = 0; c0 < TILE_N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr0 + c0, v);
        tmem_ld_wait();
        float s[16];
        uint32_t pm = 0;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 w = *reinterpret_cast<const float4*>(inv + c0 + i);
          s[i] = __uint_as_float(v[i]) * w.x;
          s[i + 1] = __uint_as_float(v[i + 1]) * w.y;
          s[i + 2] = __uint_as_float(v[i + 2]) * w.z;
          s[i + 3] = __uint_as_float(v[i + 3]) * w.w;
        }
#pragma unroll
        for (int i = 0; i < 16; i++) pm |= (s[i] > thr) ? (1u << i) : 0u;
        if (P.dbg_scores && live) {
#pragma unroll
          for (int i = 0; i < 16; i++)
            if (row0 + c0 + i < P.n_rows) P.dbg_scores[(size_t)qg * P.n_rows + row0 + c0 + i] = s[i];
        }
        if (__any_sync(0xFFFFFFFFu, pm != 0)) {
#pragma unroll
          for (int i = 0; i < 16; i++) {
            // rows arrive in increasing order, so a later equal score can never displace an earlier one:
            // the strict float compare is exact and the packed-key insert keeps (score desc, row asc)
            unsigned pending = __ballot_sync(0xFFFFFFFFu, s[i] > thr);
            const uint64_t key = rag_pack_key(s[i], row0 + c0 + i);
            while (pending) {
              const int src = __ffs(pending) - 1;
              pending &= pending - 1;
              const uint64_t kk = shfl_u64(key, src);
              uint64_t t;
              warp_list_insert(warp_lists + (size_t)src * P.kp, (int)P.kp, kk, lane, t);
              if (lane == src) thr = t != 0ull ? rag_key_score(t) : -INFINITY;
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(tmem_empty);
        mbar_arrive(&inv_empty[buf]);
      }
    }
    // publish this CTA's lists
    if (live) {
      uint64_t* out = P.partial + ((size_t)qg * P.parts + blockIdx.x) * P.kp;
      const uint64_t* mine = lists + (size_t)ql * P.kp;
      for (uint32_t j = 0; j < P.kp; j++) out[j] = mine[j];
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
  float* dbg = nullptr;  // set by rag_debug_k2_scores for one launch
};

size_t k2_smem_bytes(uint32_t stages, uint32_t kp) {
  return (size_t)stages * STAGE_BYTES + (size_t)QPC * kp * 8 + 2 * TILE_N * 4 + (2 * MAX_STAGES + 6) * 8 + 16;
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

int k2_available(const rag_index* idx) { return idx->shadow != nullptr && idx->inv_norm != nullptr; }

int k2_plan(rag_index* idx, uint32_t B, uint32_t kp, uint32_t* parts) {
  RAG_CHECK(k2_init(idx));
  k2_state* st = (k2_state*)idx->k2_state;
  if (kp > 64) return rag_set_error(RAG_ERR_UNSUPPORTED, "tensor path keeps at most 64 candidates per query (K'=%u)", kp);
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

int k2_launch(rag_index* idx, uint32_t B, uint32_t kp, uint32_t parts) {
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
  const size_t smem = k2_smem_bytes(P.stages, kp);
  if (!st->attr_set) {
    RAG_CUDA(cudaFuncSetAttribute(k2_tensor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, st->max_smem));
    st->attr_set = true;
  }
  const uint32_t groups = (B + QPC - 1) / QPC;
  k2_tensor_kernel<<<dim3(parts, groups), K2_THREADS, smem, idx->stream>>>(map_q, map_x, P);
  RAG_CUDA(cudaGetLastError());
  idx->launches++;
  return RAG_OK;
}

void k2_set_debug(rag_index* idx, float* d_scores) {
  if (k2_init(idx) == RAG_OK) ((k2_state*)idx->k2_state)->dbg = d_scores;
}

void k2_destroy(rag_index* idx) {
  delete (k2_state*)idx->k2_state;
  idx->k2_state = nullptr;
}
