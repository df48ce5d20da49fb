// k3_merge.cu — K3: merge the per-CTA candidate lists of K1/K2 into K' candidates per query.
//
// Input  partial[B][parts][kp]  packed keys (sorted per part, 0 = empty)
// Output cand[B][RAG_MAX_CANDIDATES] packed keys sorted descending (score desc, row asc)
//
// One CTA per query. Every warp scans a strided share of the parts*kp keys through the
// same threshold-guarded sorted list K1 uses, then warp 0 merges the 16 warp lists.
// Bytes: parts*kp*8 per query (≈19 KB at 148 parts, K'=16) — latency-, not bandwidth-bound.
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
}  // namespace

int k3_launch(rag_index* idx, uint32_t B, uint32_t kp, uint32_t parts) {
  rag_prof_scope ps(idx, RAG_PROF_MERGE);
  k3_merge_kernel<<<B, K3_THREADS, 0, idx->stream>>>(idx->cur->d_partial, parts * kp, kp, idx->cur->d_cand);
  RAG_CUDA(cudaGetLastError());
  idx->launches++;
  return RAG_OK;
}
