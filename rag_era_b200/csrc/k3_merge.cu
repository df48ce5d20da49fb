// k3_merge.cu — K3: merge the per-CTA candidate lists of K1/K2 into K' candidates per query.
//
// Input  partial[B][parts][kp]  packed keys (sorted per part, 0 = empty)
// Output cand[B][RAG_MAX_CANDIDATES] packed keys sorted descending (score desc, row asc)
//
// Tournament kernel (default): one WARP per query, several queries per CTA. The query's parts*kp keys are
// copied to shared memory (coalesced); every lane owns the lists l, l+32, ... and keeps their heads in
// registers; K' rounds of {warp-wide 64-bit max, the winning lane advances its list} emit the merged order.
// ~100 cycles per emitted key: the batch-1024 merge takes a few microseconds instead of 55 (the scan
// kernel below inserts into per-warp sorted lists, which is serial per surviving key).
// Scan kernel (fallback when a query's keys exceed the shared-memory budget): one CTA per query, every warp
// scans a strided share of the keys through the threshold-guarded sorted list K1 uses, warp 0 merges.
// Bytes: parts*kp*8 per query (4.6 KB at 18 parts, K'=32) — latency-, not bandwidth-bound.
#include "common.cuh"

namespace {
constexpr int K3_THREADS = 512;
constexpr int K3_WARPS = K3_THREADS / 32;

__global__ void __launch_bounds__(K3_THREADS)
k3_merge_kernel(const uint64_t* __restrict__ partial, uint32_t m /* parts*kp */, uint32_t kp,
                uint64_t* __restrict__ cand) {
  __shared__ uint64_t lists[K3_WARPS * RAG_MAX_CANDIDATES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t b = blockIdx.x;
  const uint64_t* in = partial + (size_t)b * m;
  uint64_t* mylist = lists + (size_t)warp * kp;
  for (uint32_t i = lane; i < kp; i += 32) mylist[i] = 0ull;
  __syncwarp();

  uint64_t thresh = 0ull;
  for (uint32_t base = warp * 32; base < m; base += K3_THREADS) {
    const uint32_t i = base + lane;
    const uint64_t key = i < m ? in[i] : 0ull;
    unsigned pending = __ballot_sync(0xFFFFFFFFu, key > thresh);
    while (pending) {
      const int src = __ffs(pending) - 1;
      pending &= pending - 1;
      const uint64_t kk = shfl_u64(key, src);
      if (kk > thresh) warp_list_insert(mylist, kp, kk, lane, thresh);
    }
  }
  __syncthreads();
  if (warp == 0) {
    uint64_t* out = cand + (size_t)b * RAG_MAX_CANDIDATES;
    warp_merge_lists(lists, K3_WARPS, kp, kp, out, lane);
    for (uint32_t i = kp + lane; i < RAG_MAX_CANDIDATES; i += 32) out[i] = 0ull;
  }
}
constexpr int K3T_MAX_WARPS = 8;
constexpr int K3T_LISTS_PER_LANE = 8;             // parts <= 256
constexpr size_t K3T_SMEM_BUDGET = 96 * 1024;

__global__ void __launch_bounds__(K3T_MAX_WARPS * 32)
k3_tournament_kernel(const uint64_t* __restrict__ partial, uint32_t B, uint32_t parts, uint32_t kp, uint64_t* __restrict__ cand) {
  extern __shared__ __align__(16) uint64_t k3t_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const uint32_t b = blockIdx.x * warps + warp;
  if (b >= B) return;
  const uint32_t m = parts * kp;
  uint64_t* sl = k3t_smem + (size_t)warp * m;
  const uint64_t* in = partial + (size_t)b * m;
  for (uint32_t i = lane; i < m; i += 32) sl[i] = in[i];
  __syncwarp();

  uint32_t pos[K3T_LISTS_PER_LANE];
  uint64_t head[K3T_LISTS_PER_LANE];
  uint64_t best = 0ull;
#pragma unroll
  for (int j = 0; j < K3T_LISTS_PER_LANE; j++) {
    const uint32_t list = lane + 32 * j;
    pos[j] = list * kp;
    head[j] = list < parts ? sl[pos[j]] : 0ull;
    best = max(best, head[j]);
  }
  uint64_t* out = cand + (size_t)b * RAG_MAX_CANDIDATES;
  uint32_t r = 0;
  for (; r < kp; r++) {
    const uint64_t w = warp_max_u64(best);
    if (w == 0ull) break;  // every list is exhausted: the remaining ranks stay empty
    if (lane == 0) out[r] = w;
    if (best == w) {  // keys are unique (the row is part of the key): exactly one lane, one list
      best = 0ull;
#pragma unroll
      for (int j = 0; j < K3T_LISTS_PER_LANE; j++) {
        if (head[j] == w) {
          pos[j]++;
          head[j] = pos[j] < (lane + 32 * j + 1) * kp ? sl[pos[j]] : 0ull;
        }
        best = max(best, head[j]);
      }
    }
  }
  for (uint32_t i = r + lane; i < RAG_MAX_CANDIDATES; i += 32) out[i] = 0ull;
}
}  // namespace

int k3_launch(rag_index* idx, uint32_t B, uint32_t kp, uint32_t parts) {
  rag_prof_scope ps(idx, RAG_PROF_MERGE);
  const size_t per_query = (size_t)parts * kp * 8;
  if (parts <= 32 * K3T_LISTS_PER_LANE && per_query <= K3T_SMEM_BUDGET) {
    uint32_t warps = (uint32_t)(K3T_SMEM_BUDGET / per_query);
    if (warps > K3T_MAX_WARPS) warps = K3T_MAX_WARPS;
    // spread small batches over the SMs instead of packing 8 queries into few CTAs
    while (warps > 1 && (B + warps - 1) / warps < (uint32_t)idx->sm_count) warps >>= 1;
    // the attribute is per device: set it whenever the launch needs more than the default 48 KB
    if (warps * per_query > 48 * 1024)
      RAG_CUDA(cudaFuncSetAttribute(k3_tournament_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K3T_SMEM_BUDGET));
    k3_tournament_kernel<<<(B + warps - 1) / warps, warps * 32, warps * per_query, idx->stream>>>(idx->cur->d_partial, B, parts, kp,
                                                                                                 idx->cur->d_cand);
  } else {
    k3_merge_kernel<<<B, K3_THREADS, 0, idx->stream>>>(idx->cur->d_partial, parts * kp, kp, idx->cur->d_cand);
  }
  RAG_CUDA(cudaGetLastError());
  idx->launches++;
  return RAG_OK;
}
