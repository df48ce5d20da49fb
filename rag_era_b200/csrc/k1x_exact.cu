// k1x_exact.cu — K1x: the exact path. Every row is scored with the reference's own
// arithmetic (binary64, left to right, no FMA — exact_chain.cuh), so candidate selection
// has no approximation error beyond one fp32 rounding of the exact cosine, which is
// monotone: exact(a) > exact(b)  ⇒  fp32(exact(a)) >= fp32(exact(b)). K4 then orders the
// K' survivors on the unrounded values.
//
// This is the escalation target for queries the stream / tensor paths cannot certify
// (near-ties across the K' boundary), and RAG_PATH_EXACT for callers who want it.
// Roofline: FP64 pipe — 2 dependent chains of D steps per row (4 fp64 ops + 1 conversion per element;
// the ||q||^2 chain runs once per warp).
// Each warp takes 32 consecutive rows at a time; rows past the end are clamped to the last
// row and discarded.
#include "exact_chain.cuh"

namespace {
using namespace rag_exact;

constexpr int K1X_WARPS = 4;

template <bool BF16>
__global__ void __launch_bounds__(K1X_WARPS * 32)
k1x_exact_kernel(const void* __restrict__ X, uint32_t n_rows, uint32_t ld, const float* __restrict__ Q,
                 uint32_t kp, uint32_t parts, uint64_t* __restrict__ partial) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem + (size_t)K1X_WARPS * WARP_BYTES);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t b = blockIdx.y;
  const float* q = Q + (size_t)b * ld;
  unsigned char* wsm = smem + (size_t)warp * WARP_BYTES;
  uint64_t* mylist = lists + (size_t)warp * kp;
  for (uint32_t i = lane; i < kp; i += 32) mylist[i] = 0ull;
  __syncwarp();

  // ||q||^2 is the same chain for every row: once per warp
  const double nq = query_norm_sq(q, ld, lane);

  uint64_t thresh = 0ull;
  const uint32_t n_blocks = (n_rows + 31) / 32;
  for (uint32_t blk = blockIdx.x * K1X_WARPS + warp; blk < n_blocks; blk += gridDim.x * K1X_WARPS) {
    const uint32_t row = blk * 32 + lane;
    const bool in_range = row < n_rows;
    chains c = warp_exact_sums<BF16>(X, ld, q, in_range ? row : n_rows - 1, wsm, lane);
    c.nq = nq;
    const double s = finish(c);
    // zero-norm rows: NaN in the reference; defined as never selected (SURVEY N-nan)
    const float sf = (in_range && s == s) ? __double2float_rn(s) : -INFINITY;
    const uint64_t key = in_range ? rag_pack_key(sf, row) : 0ull;
    unsigned pending = __ballot_sync(0xFFFFFFFFu, key > thresh);
    while (pending) {
      const int src = __ffs(pending) - 1;
      pending &= pending - 1;
      const uint64_t kk = shfl_u64(key, src);
      if (kk > thresh) warp_list_insert(mylist, kp, kk, lane, thresh);
    }
  }
  __syncthreads();
  if (warp == 0)
    warp_merge_lists(lists, K1X_WARPS, kp, kp, partial + ((size_t)b * parts + blockIdx.x) * kp, lane);
}

}  // namespace

int k1x_plan(const rag_index* idx, uint32_t B, uint32_t kp, uint32_t* parts) {
  (void)kp;
  const uint64_t n_blocks = (idx->rows + 31) / 32;
  uint64_t p = (n_blocks + K1X_WARPS - 1) / K1X_WARPS;
  // a few waves per SM when there is one query; fewer parts per query as the batch grows
  uint64_t cap = (uint64_t)idx->sm_count * (B >= 8 ? 1 : 4);
  if (p > cap) p = cap;
  if (p < 1) p = 1;
  *parts = (uint32_t)p;
  return RAG_OK;
}

int k1x_launch(rag_index* idx, uint32_t B, uint32_t kp, uint32_t parts) {
  rag_prof_scope ps(idx, RAG_PROF_STREAM);
  if (idx->rows == 0) return rag_set_error(RAG_ERR_STATE, "search on an empty index");
  if (idx->rows >= 0xFFFFFFFFull) return rag_set_error(RAG_ERR_UNSUPPORTED, "more than 2^32-2 rows per shard");
  const size_t smem = (size_t)K1X_WARPS * WARP_BYTES + (size_t)K1X_WARPS * kp * sizeof(uint64_t);
  const bool bf16 = idx->desc.dtype == RAG_BF16;
  auto kern = bf16 ? k1x_exact_kernel<true> : k1x_exact_kernel<false>;
  RAG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3(parts, B), K1X_WARPS * 32, smem, idx->stream>>>(idx->corpus, (uint32_t)idx->rows, idx->ld,
                                                               idx->cur->d_q, kp, parts, idx->cur->d_partial);
  RAG_CUDA(cudaGetLastError());
  idx->launches++;
  return RAG_OK;
}
