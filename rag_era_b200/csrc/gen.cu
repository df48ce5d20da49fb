// gen.cu — maintenance kernels: on-device synthetic corpus / queries / metadata
// (definition in include/ragera_gen.h; bit-identical to the host generator), the bf16
// shadow + inverse-norm builder for the tensor path, and the query fp32→bf16 cast.
#include "common.cuh"

#include <cuda_fp16.h>
#include <stdlib.h>

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

// ---- tensor-path operands and their rounding residuals (the rigorous certification bound, api.cu::make_plan) ------
// What K2 multiplies:            queries                       rows
//   fp32 index + fp16 shadow     fp16( q / ||q|| )             fp16( x / ||x|| )      (pre-normalised: no scaling in the epilogue)
//   fp32 index + bf16 shadow     bf16( q / ||q|| )             bf16( x ),   scaled by 1/||x|| in the epilogue
//   bf16 index                   bf16( q / ||q|| )             x itself,    scaled by 1/||x||
//   fp32 index, no shadow        q (tf32, converted by TMA)    x (tf32),    scaled by 1/||x||
// fp16 keeps 11 significant bits against bf16's 8, at the same tensor-core rate (kind::f16; both operands must have
// the same format); the normalisation puts every element in [-1, 1], far from fp16's range limits. The residuals
//   rho_q[b] = || operand(q) - q/||q|| ||           rho_x = max over rows || operand(x) - x || / ||x||
// are MEASURED here (fp32 arithmetic, differences of nearby floats are exact), so outliers, subnormals or a value
// that sits on a rounding midpoint are accounted for as they are — there is no statistical assumption.
__device__ __forceinline__ float sq(float e) { return e * e; }
__device__ __forceinline__ float resid2_bf16(float v, uint16_t b) { return sq(__fsub_rn(v, rg_bf16_to_f32(b))); }
__device__ __forceinline__ float resid2_f16(float v, __half h) { return sq(__fsub_rn(v, __half2float(h))); }
// tf32 (fp32 rows read by TMA as TFLOAT32): the hardware keeps 10 mantissa bits, by truncation or by rounding —
// either way the operand is one of the two tf32 neighbours of v, so the bound is the distance to the FARTHER
// neighbour (0 when v is itself a tf32 value).
__device__ __forceinline__ float resid2_tf32(float v) {
  const uint32_t u = __float_as_uint(v);
  if ((u & 0x1FFFu) == 0u) return 0.f;
  const float lo = __uint_as_float(u & ~0x1FFFu);              // truncation towards zero
  const float hi = __uint_as_float((u & ~0x1FFFu) + 0x2000u);  // next tf32 value away from zero
  return sq(fmaxf(fabsf(__fsub_rn(v, lo)), fabsf(__fsub_rn(hi, v))));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b, float& ee) {
  const __half ha = __float2half_rn(a), hb = __float2half_rn(b);
  ee += resid2_f16(a, ha) + resid2_f16(b, hb);
  return (uint32_t)__half_as_ushort(ha) | ((uint32_t)__half_as_ushort(hb) << 16);
}
__device__ __forceinline__ uint32_t pack_b2(float a, float b, float& ee) {
  const uint16_t ba = rg_f32_to_bf16(a), bb = rg_f32_to_bf16(b);
  ee += resid2_bf16(a, ba) + resid2_bf16(b, bb);
  return (uint32_t)ba | ((uint32_t)bb << 16);
}

// one warp per row. SHADOW: 0 = none (bf16 corpus: rho 0; fp32 corpus: the tf32 residual), 1 = bf16 copy,
// 2 = fp16 copy of the NORMALISED row (a zero row becomes NaNs: it can never be selected).
// inv_norm = 1/||x|| of the row itself (not of its rounded copy).
template <bool SRC_BF16, int SHADOW>
__global__ void aux_build_kernel(const void* __restrict__ X, uint16_t* __restrict__ shadow,
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
      uint2* sp = reinterpret_cast<uint2*>(shadow + row * ld);
      for (uint32_t i = lane; i < ld / 4; i += 32) {
        const float4 v = p[i];
        ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
        if (SHADOW == 1) {
          const uint16_t b0 = rg_f32_to_bf16(v.x), b1 = rg_f32_to_bf16(v.y), b2 = rg_f32_to_bf16(v.z),
                         b3 = rg_f32_to_bf16(v.w);
          sp[i] = make_uint2((uint32_t)b0 | ((uint32_t)b1 << 16), (uint32_t)b2 | ((uint32_t)b3 << 16));
          ee += resid2_bf16(v.x, b0) + resid2_bf16(v.y, b1) + resid2_bf16(v.z, b2) + resid2_bf16(v.w, b3);
        } else if (SHADOW == 0) {
          ee += resid2_tf32(v.x) + resid2_tf32(v.y) + resid2_tf32(v.z) + resid2_tf32(v.w);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
    const float inv = ss > 0.f ? rsqrtf(ss) : 0.f;
    if (!SRC_BF16 && SHADOW == 2) {
      // second walk over the row (L1/L2 hits): the fp16 copy of x * (1/||x||) and its residual, already relative
      const float4* p = reinterpret_cast<const float4*>((const float*)X + row * ld);
      uint2* sp = reinterpret_cast<uint2*>(shadow + row * ld);
      for (uint32_t i = lane; i < ld / 4; i += 32) {
        const float4 v = p[i];
        if (inv > 0.f) sp[i] = make_uint2(pack_h2(v.x * inv, v.y * inv, ee), pack_h2(v.z * inv, v.w * inv, ee));
        else sp[i] = make_uint2(0x7E007E00u, 0x7E007E00u);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ee += __shfl_xor_sync(0xFFFFFFFFu, ee, o);
    if (lane == 0) {
      inv_norm[row] = inv;
      if (!SRC_BF16 && ss > 0.f) rho_max = fmaxf(rho_max, SHADOW == 2 ? sqrtf(ee) : sqrtf(ee / ss));
    }
  }
  if (!SRC_BF16 && lane == 0 && rho_max > 0.f) atomicMax(rho_bits, __float_as_uint(rho_max));
}

// one warp per query: the fp16 operand q / ||q|| of the 16-bit tensor path (rows >= B are zero padding) and the
// query's residual rho_q[b] = || fp16(q/||q||) - q/||q|| ||. QBF16: bf16 instead — kind::f16 wants both operands in ONE
// 16-bit format (an fp16 x bf16 instruction descriptor is an illegal instruction on sm_100a, measured), so bf16 rows
// (a bf16 corpus, a bf16 shadow) take bf16 queries.
template <bool QBF16>
__global__ void q_to_f16_kernel(const float* __restrict__ q, uint16_t* __restrict__ qh, float* __restrict__ rho_q,
                                uint32_t B, uint32_t Bpad, uint32_t ld) {
  const int lane = threadIdx.x & 31;
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t b = w; b < Bpad; b += nw) {
    const float4* src = reinterpret_cast<const float4*>(q + (size_t)b * ld);
    uint2* dst = reinterpret_cast<uint2*>(qh + (size_t)b * ld);
    float ss = 0.f, ee = 0.f;
    if (b < B)
      for (uint32_t i = lane; i < ld / 4; i += 32) {
        const float4 v = src[i];
        ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
    const float inv = ss > 0.f ? rsqrtf(ss) : 0.f;
    for (uint32_t i = lane; i < ld / 4; i += 32) {
      const float4 v = b < B ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      dst[i] = QBF16 ? make_uint2(pack_b2(v.x * inv, v.y * inv, ee), pack_b2(v.z * inv, v.w * inv, ee))
                     : make_uint2(pack_h2(v.x * inv, v.y * inv, ee), pack_h2(v.z * inv, v.w * inv, ee));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ee += __shfl_xor_sync(0xFFFFFFFFu, ee, o);
    if (lane == 0 && b < B) rho_q[b] = sqrtf(ee);
  }
}

// the residual of queries the tensor path reads as tf32 (no cast: TMA converts the fp32 queries)
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
  uint16_t* sh = reinterpret_cast<uint16_t*>(idx->shadow);
#define AUX_GO(SRC, MODE) aux_build_kernel<SRC, MODE><<<grid, 256, 0, idx->stream>>>(idx->corpus, sh, idx->inv_norm, idx->d_rho_x, row0, nrows, idx->ld)
  if (idx->desc.dtype == RAG_BF16) AUX_GO(true, 0);
  else if (!idx->shadow) AUX_GO(false, 0);
  else if (idx->shadow_f16) AUX_GO(false, 2);
  else AUX_GO(false, 1);
#undef AUX_GO
  RAG_CUDA(cudaGetLastError());
  idx->rho_x_stale = true;
  idx->launches++;
  return RAG_OK;
}

// the queries' tensor-path operand (normalised fp16, or nothing for tf32) and their rounding residuals rho_q
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
  else if (!idx->shadow_f16) q_to_f16_kernel<true><<<grid, 256, 0, idx->stream>>>(bt->d_q, reinterpret_cast<uint16_t*>(bt->d_qb), bt->d_rho_q, B, Bpad, idx->ld);
  else q_to_f16_kernel<false><<<grid, 256, 0, idx->stream>>>(bt->d_q, reinterpret_cast<uint16_t*>(bt->d_qb), bt->d_rho_q, B, Bpad, idx->ld);
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
