// gen.cu — maintenance kernels: on-device synthetic corpus / queries / metadata
// (definition in include/ragera_gen.h; bit-identical to the host generator), the bf16
// shadow + inverse-norm builder for the tensor path, and the query fp32→bf16 cast.
#include "common.cuh"

namespace {

// one thread produces 8 consecutive columns of one row (two float4 / one uint4 store)
template <bool BF16>
__global__ void gen_corpus_kernel(void* __restrict__ X, uint64_t nrows, uint32_t dim, uint32_t ld,
                                  uint64_t id_base, rag_gen_desc g) {
  const uint32_t groups = ld / 8;
  const uint64_t total = nrows * groups;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t row = t / groups;
    const uint32_t c0 = (uint32_t)(t % groups) * 8;
    const uint64_t grow = id_base + row;
    const uint64_t src = rg_src_row(&g, grow);
    const uint32_t cl = rg_cluster(&g, src);
    const float sc = rg_row_scale(&g, src);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const uint32_t col = c0 + i;
      if (col < dim) {
        const float centre = rg_gauss(rg_hash3(g.seed ^ 0xCE27E5u, cl, col));
        const float nz = rg_gauss(rg_hash3(g.seed, src, col));
        v[i] = __fmul_rn(sc, __fadd_rn(centre, __fmul_rn(g.noise, nz)));
      } else {
        v[i] = 0.0f;
      }
    }
    if (BF16) {
      uint4 o;
      o.x = (uint32_t)rg_f32_to_bf16(v[0]) | ((uint32_t)rg_f32_to_bf16(v[1]) << 16);
      o.y = (uint32_t)rg_f32_to_bf16(v[2]) | ((uint32_t)rg_f32_to_bf16(v[3]) << 16);
      o.z = (uint32_t)rg_f32_to_bf16(v[4]) | ((uint32_t)rg_f32_to_bf16(v[5]) << 16);
      o.w = (uint32_t)rg_f32_to_bf16(v[6]) | ((uint32_t)rg_f32_to_bf16(v[7]) << 16);
      reinterpret_cast<uint4*>((uint16_t*)X + row * ld + c0)[0] = o;
    } else {
      float4* p = reinterpret_cast<float4*>((float*)X + row * ld + c0);
      p[0] = make_float4(v[0], v[1], v[2], v[3]);
      p[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

__global__ void gen_queries_kernel(float* __restrict__ out, uint64_t b0, uint32_t B, uint32_t dim, rag_gen_desc g) {
  const uint64_t total = (uint64_t)B * dim;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t b = t / dim;
    const uint32_t col = (uint32_t)(t % dim);
    out[t] = rg_query_elem(&g, b0 + b, col);
  }
}

__global__ void gen_meta_kernel(uint64_t nrows, uint64_t id_base, rag_gen_desc g, uint8_t* ctype, double* conf,
                                int32_t* access, int64_t* last_ms) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  const uint64_t r = id_base + i;
  ctype[i] = r < g.memory_rows ? RAG_CT_MEMORY : RAG_CT_DOCUMENT;
  conf[i] = rg_meta_confidence(&g, r);
  access[i] = rg_meta_access(&g, r);
  last_ms[i] = rg_meta_last_access_ms(&g, r);
}

// Rounding residual of one fp32 value against the operand the tensor path actually multiplies, squared.
//   bf16 shadow: (v - bf16_rne(v))^2, exact in fp32 (the difference of two floats this close is representable).
//   tf32 (fp32 rows read by TMA as TFLOAT32): the hardware keeps 10 mantissa bits, by truncation or by rounding
//   — either way the operand is one of the two tf32 neighbours of v, so the bound is the distance to the
//   FARTHER neighbour (0 when v is itself a tf32 value).
__device__ __forceinline__ float resid2_bf16(float v, uint16_t b) {
  const float e = __fsub_rn(v, rg_bf16_to_f32(b));
  return e * e;
}
__device__ __forceinline__ float resid2_tf32(float v) {
  const uint32_t u = __float_as_uint(v);
  const uint32_t low = u & 0x1FFFu;
  if (low == 0u) return 0.f;
  const float lo = __uint_as_float(u & ~0x1FFFu);          // truncation towards zero
  const float hi = __uint_as_float((u & ~0x1FFFu) + 0x2000u);  // next tf32 value away from zero
  const float e = fmaxf(fabsf(__fsub_rn(v, lo)), fabsf(__fsub_rn(hi, v)));
  return e * e;
}

// one warp per row: bf16 shadow (fp32 corpus only), 1/||x|| of the row itself (NOT of its rounded copy: the
// selected score then differs from the exact cosine only by the operand rounding, which rho bounds), and
// rho_x = max over rows of ||x - operand(x)|| / ||x|| (atomicMax on the float's bits; 0 for a bf16 corpus,
// whose rows are the operand)
template <bool SRC_BF16>
__global__ void aux_build_kernel(const void* __restrict__ X, __nv_bfloat16* __restrict__ shadow,
                                 float* __restrict__ inv_norm, uint32_t* __restrict__ rho_bits, uint64_t row0,
                                 uint64_t nrows, uint32_t ld) {
  const int lane = threadIdx.x & 31;
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  float rho_max = 0.f;
  for (uint64_t r = w; r < nrows; r += nw) {
    const uint64_t row = row0 + r;
    float ss = 0.f, ee = 0.f;
    if (SRC_BF16) {
      const uint4* p = reinterpret_cast<const uint4*>((const uint16_t*)X + row * ld);
      for (uint32_t i = lane; i < ld / 8; i += 32) {
        const uint4 v = p[i];
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const float a = __uint_as_float(u[j] << 16), b = __uint_as_float(u[j] & 0xFFFF0000u);
          ss = fmaf(a, a, ss); ss = fmaf(b, b, ss);
        }
      }
    } else {
      const float4* p = reinterpret_cast<const float4*>((const float*)X + row * ld);
      uint2* sp = shadow ? reinterpret_cast<uint2*>(shadow + row * ld) : nullptr;
      for (uint32_t i = lane; i < ld / 4; i += 32) {
        const float4 v = p[i];
        ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
        if (sp) {
          const uint16_t b0 = rg_f32_to_bf16(v.x), b1 = rg_f32_to_bf16(v.y), b2 = rg_f32_to_bf16(v.z),
                         b3 = rg_f32_to_bf16(v.w);
          sp[i] = make_uint2((uint32_t)b0 | ((uint32_t)b1 << 16), (uint32_t)b2 | ((uint32_t)b3 << 16));
          ee += resid2_bf16(v.x, b0) + resid2_bf16(v.y, b1) + resid2_bf16(v.z, b2) + resid2_bf16(v.w, b3);
        } else {
          ee += resid2_tf32(v.x) + resid2_tf32(v.y) + resid2_tf32(v.z) + resid2_tf32(v.w);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
      ee += __shfl_xor_sync(0xFFFFFFFFu, ee, o);
    }
    if (lane == 0) {
      inv_norm[row] = ss > 0.f ? rsqrtf(ss) : 0.f;
      if (!SRC_BF16 && ss > 0.f) rho_max = fmaxf(rho_max, sqrtf(ee / ss));
    }
  }
  if (!SRC_BF16 && lane == 0 && rho_max > 0.f) atomicMax(rho_bits, __float_as_uint(rho_max));
}

// one warp per query: the bf16 operand of the tensor path (rows >= B are zero padding) and the query's
// rounding residual rho_q[b] = ||q - bf16(q)|| / ||q|| for the rigorous certification bound
__global__ void q_to_bf16_kernel(const float* __restrict__ q, __nv_bfloat16* __restrict__ qb, float* __restrict__ rho_q,
                                 uint32_t B, uint32_t Bpad, uint32_t ld) {
  const int lane = threadIdx.x & 31;
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t b = w; b < Bpad; b += nw) {
    float ss = 0.f, ee = 0.f;
    const float4* src = reinterpret_cast<const float4*>(q + (size_t)b * ld);
    uint2* dst = reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(qb) + (size_t)b * ld);
    for (uint32_t i = lane; i < ld / 4; i += 32) {
      const float4 v = b < B ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      const uint16_t b0 = rg_f32_to_bf16(v.x), b1 = rg_f32_to_bf16(v.y), b2 = rg_f32_to_bf16(v.z), b3 = rg_f32_to_bf16(v.w);
      dst[i] = make_uint2((uint32_t)b0 | ((uint32_t)b1 << 16), (uint32_t)b2 | ((uint32_t)b3 << 16));
      ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
      ee += resid2_bf16(v.x, b0) + resid2_bf16(v.y, b1) + resid2_bf16(v.z, b2) + resid2_bf16(v.w, b3);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
      ee += __shfl_xor_sync(0xFFFFFFFFu, ee, o);
    }
    if (lane == 0 && b < B) rho_q[b] = ss > 0.f ? sqrtf(ee / ss) : 0.f;
  }
}

// the same residual for queries the tensor path reads as tf32 (no cast: TMA converts the fp32 queries)
__global__ void q_rho_tf32_kernel(const float* __restrict__ q, float* __restrict__ rho_q, uint32_t B, uint32_t ld) {
  const int lane = threadIdx.x & 31;
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t b = w; b < B; b += nw) {
    float ss = 0.f, ee = 0.f;
    for (uint32_t i = lane; i < ld; i += 32) {
      const float v = q[(size_t)b * ld + i];
      ss = fmaf(v, v, ss);
      ee += resid2_tf32(v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
      ee += __shfl_xor_sync(0xFFFFFFFFu, ee, o);
    }
    if (lane == 0) rho_q[b] = ss > 0.f ? sqrtf(ee / ss) : 0.f;
  }
}

// escalation: copy the selected queries (and their keyword lists) of one batch into another
__global__ void gather_batch_kernel(const float* __restrict__ q_src, float* __restrict__ q_dst, uint32_t ld,
                                    const uint64_t* __restrict__ kw_src, uint64_t* __restrict__ kw_dst,
                                    const uint32_t* __restrict__ kwc_src, uint32_t* __restrict__ kwc_dst,
                                    uint32_t kw_stride, const uint32_t* __restrict__ sel) {
  const uint32_t i = blockIdx.x, b = sel[i];
  const float4* s = reinterpret_cast<const float4*>(q_src + (size_t)b * ld);
  float4* d = reinterpret_cast<float4*>(q_dst + (size_t)i * ld);
  for (uint32_t c = threadIdx.x; c < ld / 4; c += blockDim.x) d[c] = s[c];
  if (kw_src) {
    for (uint32_t c = threadIdx.x; c < kw_stride; c += blockDim.x)
      kw_dst[(size_t)i * kw_stride + c] = kw_src[(size_t)b * kw_stride + c];
    if (threadIdx.x == 0) kwc_dst[i] = kwc_src[b];
  }
}

__global__ void iota_u64_kernel(uint64_t* d, uint64_t n, uint64_t base) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = base + i;
}

uint32_t grid_for(uint64_t work, uint32_t threads, int sm_count) {
  uint64_t blocks = (work + threads - 1) / threads;
  uint64_t cap = (uint64_t)sm_count * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (uint32_t)blocks;
}

}  // namespace

int gen_corpus_launch(rag_index* idx, const rag_gen_desc* g, uint64_t nrows) {
  if (nrows == 0) return RAG_OK;
  const uint64_t work = nrows * (idx->ld / 8);
  const uint32_t grid = grid_for(work, 256, idx->sm_count);
  if (idx->desc.dtype == RAG_BF16)
    gen_corpus_kernel<true><<<grid, 256, 0, idx->stream>>>(idx->corpus, nrows, idx->dim, idx->ld, idx->desc.id_base, *g);
  else
    gen_corpus_kernel<false><<<grid, 256, 0, idx->stream>>>(idx->corpus, nrows, idx->dim, idx->ld, idx->desc.id_base, *g);
  RAG_CUDA(cudaGetLastError());
  idx->launches++;
  return RAG_OK;
}

int gen_queries_launch(rag_index* idx, const rag_gen_desc* g, uint64_t b0, uint32_t B, float* d_out) {
  const uint32_t grid = grid_for((uint64_t)B * idx->dim, 256, idx->sm_count);
  gen_queries_kernel<<<grid, 256, 0, idx->stream>>>(d_out, b0, B, idx->dim, *g);
  RAG_CUDA(cudaGetLastError());
  idx->launches++;
  return RAG_OK;
}

int gen_meta_launch(rag_index* idx, const rag_gen_desc* g, uint64_t nrows) {
  if (nrows == 0) return RAG_OK;
  gen_meta_kernel<<<(uint32_t)((nrows + 255) / 256), 256, 0, idx->stream>>>(nrows, idx->desc.id_base, *g, idx->ctype,
                                                                           idx->conf, idx->access, idx->last_ms);
  RAG_CUDA(cudaGetLastError());
  idx->launches++;
  return RAG_OK;
}

int aux_build_launch(rag_index* idx, uint64_t row0, uint64_t nrows) {
  if (nrows == 0 || !idx->inv_norm) return RAG_OK;
  const uint32_t grid = grid_for(nrows * 32, 256, idx->sm_count);
  if (idx->desc.dtype == RAG_BF16)
    aux_build_kernel<true><<<grid, 256, 0, idx->stream>>>(idx->corpus, nullptr, idx->inv_norm, idx->d_rho_x, row0, nrows, idx->ld);
  else
    aux_build_kernel<false><<<grid, 256, 0, idx->stream>>>(idx->corpus, idx->shadow, idx->inv_norm, idx->d_rho_x, row0, nrows, idx->ld);
  RAG_CUDA(cudaGetLastError());
  idx->rho_x_stale = true;
  idx->launches++;
  return RAG_OK;
}

// the queries' tensor-path operand (bf16 cast, or nothing for tf32) and their rounding residuals rho_q
int q_operand_launch(rag_index* idx, uint32_t B, uint32_t Bpad, bool tf32) {
  rag_batch* bt = idx->cur;
  if ((size_t)B * 4 > bt->c_rho_q || !bt->d_rho_q) {
    if (bt->d_rho_q) RAG_CUDA(cudaFree(bt->d_rho_q));
    bt->d_rho_q = nullptr;
    bt->c_rho_q = 0;
    const size_t need = (size_t)(B < 1024 ? 1024 : B) * 4;
    RAG_CUDA(cudaMalloc((void**)&bt->d_rho_q, need));
    bt->c_rho_q = need;
  }
  const uint32_t rows = tf32 ? B : Bpad;
  const uint32_t grid = grid_for((uint64_t)rows * 32, 256, idx->sm_count);
  if (tf32) q_rho_tf32_kernel<<<grid, 256, 0, idx->stream>>>(bt->d_q, bt->d_rho_q, B, idx->ld);
  else q_to_bf16_kernel<<<grid, 256, 0, idx->stream>>>(bt->d_q, bt->d_qb, bt->d_rho_q, B, Bpad, idx->ld);
  RAG_CUDA(cudaGetLastError());
  idx->launches++;
  return RAG_OK;
}

int gather_batch_launch(rag_index* idx, const rag_batch* src, rag_batch* dst, uint32_t n, uint32_t kw_stride) {
  if (n == 0) return RAG_OK;
  const bool kw = kw_stride > 0 && src->d_kw && dst->d_kw;
  gather_batch_kernel<<<n, 128, 0, idx->stream>>>(src->d_q, dst->d_q, idx->ld, kw ? src->d_kw : nullptr, dst->d_kw,
                                                  src->d_kwc, dst->d_kwc, kw_stride, dst->d_sel);
  RAG_CUDA(cudaGetLastError());
  idx->launches++;
  return RAG_OK;
}

int iota_u64_launch(rag_index* idx, uint64_t* d, uint64_t n, uint64_t base) {
  if (n == 0) return RAG_OK;
  iota_u64_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, idx->stream>>>(d, n, base);
  RAG_CUDA(cudaGetLastError());
  idx->launches++;
  return RAG_OK;
}
