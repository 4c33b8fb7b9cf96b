// Row-wise pieces of the attention backward (autograd of xformers.ops.memory_efficient_attention as called at
// croco/blocks.py:126-130,192-196).  The five contractions of the backward run as batched tcgen05 GEMMs
// (s3r_gemm_bf16_batched: S = Q K^T, dP = dO V^T, dV = P^T dO, dK = dS^T Q, dQ = dS K); these two kernels are the
// softmax recomputation and the softmax Jacobian in between, one warp per (image, head, query) row:
//   P  = softmax(scale * S)                                   fp32 S -> bf16 P   (columns [n, ldp) zero-filled)
//   dS = scale * P o (dP - rowsum(dO o O))                     bf16 dS           (columns [n, ldp) zero-filled)
// HBM-bound: 4 + 2 (resp. 2 + 4 + 2) bytes per score.
#include <cuda_bf16.h>

#include "s3r_common.cuh"

#define AB_WARPS 8

__global__ void __launch_bounds__(32 * AB_WARPS) s3r_softmax_rows_kernel(const float* __restrict__ S,
                                                                        __nv_bfloat16* __restrict__ P, long long rows,
                                                                        int n, int ld, int ldp, float scale) {
  const long long row = (long long)blockIdx.x * AB_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* s = S + row * ld;
  __nv_bfloat16* p = P + row * ldp;
  float mx = -INFINITY;
  for (int c = lane; c < n; c += 32) mx = fmaxf(mx, s[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float k2 = scale * 1.4426950408889634f;  // exp(scale * (s - mx)) = exp2(k2 * (s - mx))
  float sum = 0.f;
  for (int c = lane; c < n; c += 32) sum += exp2f(k2 * (s[c] - mx));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
  for (int c = lane; c < ldp; c += 32) p[c] = __float2bfloat16_rn(c < n ? exp2f(k2 * (s[c] - mx)) * inv : 0.0f);
}

__global__ void __launch_bounds__(32 * AB_WARPS) s3r_attn_ds_kernel(const __nv_bfloat16* __restrict__ P,
                                                                   const float* __restrict__ dP,
                                                                   const __nv_bfloat16* __restrict__ O,
                                                                   const __nv_bfloat16* __restrict__ dO,
                                                                   __nv_bfloat16* __restrict__ dS, long long rows, int n,
                                                                   int ld, int ldp, float scale) {
  const long long row = (long long)blockIdx.x * AB_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  // D = sum_d dO[row, d] * O[row, d], head_dim 64: two elements per lane
  const __nv_bfloat162 o2 = reinterpret_cast<const __nv_bfloat162*>(O + row * 64)[lane];
  const __nv_bfloat162 g2 = reinterpret_cast<const __nv_bfloat162*>(dO + row * 64)[lane];
  const float2 of = __bfloat1622float2(o2), gf = __bfloat1622float2(g2);
  float D = of.x * gf.x + of.y * gf.y;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) D += __shfl_xor_sync(0xffffffffu, D, o);
  const __nv_bfloat16* p = P + row * ldp;
  const float* dp = dP + row * ld;
  __nv_bfloat16* ds = dS + row * ldp;
  for (int c = lane; c < ldp; c += 32)
    ds[c] = __float2bfloat16_rn(c < n ? scale * __bfloat162float(p[c]) * (dp[c] - D) : 0.0f);
}

extern "C" int s3r_softmax_rows_bf16(const float* S, void* P, int64_t rows, int32_t n, int32_t ld, int32_t ldp, float scale,
                                     void* stream) {
  if (rows < 0 || n <= 0 || ld < n || ldp < n) return S3R_ERR_INVALID_ARG;
  if (rows == 0) return S3R_OK;
  if (!S || !P) return S3R_ERR_INVALID_ARG;
  const long long blocks = (rows + AB_WARPS - 1) / AB_WARPS;
  if (blocks > 0x7fffffffLL) return S3R_ERR_UNSUPPORTED;
  s3r_softmax_rows_kernel<<<(unsigned)blocks, 32 * AB_WARPS, 0, (cudaStream_t)stream>>>(S, (__nv_bfloat16*)P, rows, n, ld,
                                                                                        ldp, scale);
  S3R_CUDA_CHECK(cudaGetLastError());
  return S3R_OK;
}

extern "C" int s3r_attention_ds_bf16(const void* P, const float* dP, const void* O, const void* dO, void* dS, int64_t rows,
                                     int32_t n, int32_t ld, int32_t ldp, float scale, void* stream) {
  if (rows < 0 || n <= 0 || ld < n || ldp < n) return S3R_ERR_INVALID_ARG;
  if (rows == 0) return S3R_OK;
  if (!P || !dP || !O || !dO || !dS) return S3R_ERR_INVALID_ARG;
  const long long blocks = (rows + AB_WARPS - 1) / AB_WARPS;
  if (blocks > 0x7fffffffLL) return S3R_ERR_UNSUPPORTED;
  s3r_attn_ds_kernel<<<(unsigned)blocks, 32 * AB_WARPS, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)P, dP, (const __nv_bfloat16*)O, (const __nv_bfloat16*)dO, (__nv_bfloat16*)dS, rows, n, ld, ldp,
      scale);
  S3R_CUDA_CHECK(cudaGetLastError());
  return S3R_OK;
}
