// LayerNorm over the last dimension for the bf16 ViT trunks (nn.LayerNorm(dim, eps=1e-6) of croco/blocks.py:140-147,
// 206-217 and enc_norm / dec_norm, croco.py:34): y = (x - mean) * rsqrt(var + eps) * w + b, statistics and arithmetic in
// fp32, bf16 in / out.  One warp per row, the whole row in registers (C <= 2048): 16-byte loads / stores, two-pass
// variance (no cancellation), no shared memory.  HBM-bound: 4*C bytes per row.  Launched with programmatic dependent
// launch so that its (tiny) prologue overlaps the tail of the GEMM that produced x.
#include <cuda_bf16.h>

#include "s3r_common.cuh"

#define LN_WARPS 4
#define LN_MAX_VEC 8  // 8 x (32 lanes x 8 bf16) = 2048 channels

__global__ void __launch_bounds__(32 * LN_WARPS) s3r_layernorm_kernel(const __nv_bfloat16* __restrict__ x,
                                                                      const __nv_bfloat16* __restrict__ w,
                                                                      const __nv_bfloat16* __restrict__ b,
                                                                      __nv_bfloat16* __restrict__ y, int M, int C,
                                                                      long long ldx, float eps, int pdl) {
  if (pdl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const int nvec = C / 256;  // full 32-lane x 8-element vectors per row (C % 256 == 0)
  const uint4* xr = reinterpret_cast<const uint4*>(x + (long long)row * ldx);
  float v[LN_MAX_VEC][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; i++) {
    if (i < nvec) {
      const uint4 u = xr[i * 32 + lane];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const float2 f = __bfloat1622float2(h[t]);
        v[i][2 * t] = f.x, v[i][2 * t + 1] = f.y;
        sum += f.x + f.y;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; i++) {
    if (i < nvec) {
#pragma unroll
      for (int t = 0; t < 8; t++) {
        const float d = v[i][t] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)C + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + (long long)row * C);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  const uint4* br = reinterpret_cast<const uint4*>(b);
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; i++) {
    if (i < nvec) {
      const uint4 uw = wr[i * 32 + lane], ub = br[i * 32 + lane];
      const __nv_bfloat162* hw = reinterpret_cast<const __nv_bfloat162*>(&uw);
      const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&ub);
      uint4 o;
      __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const float2 fw = __bfloat1622float2(hw[t]), fb = __bfloat1622float2(hb[t]);
        ho[t] = __floats2bfloat162_rn((v[i][2 * t] - mean) * rstd * fw.x + fb.x, (v[i][2 * t + 1] - mean) * rstd * fw.y + fb.y);
      }
      yr[i * 32 + lane] = o;
    }
  }
}

extern "C" int s3r_layernorm_bf16(const void* x, const void* weight, const void* bias, void* y, int32_t M, int32_t C,
                                  int64_t ldx, float eps, void* stream) {
  if (M < 0 || C <= 0 || ldx < C) return S3R_ERR_INVALID_ARG;
  if (M == 0) return S3R_OK;
  if (!x || !weight || !bias || !y) return S3R_ERR_INVALID_ARG;
  if (C % 256 || C > 256 * LN_MAX_VEC || ldx % 8 || (((uintptr_t)x | (uintptr_t)y | (uintptr_t)weight | (uintptr_t)bias) & 15))
    return S3R_ERR_UNSUPPORTED;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((M + LN_WARPS - 1) / LN_WARPS, 1, 1);
  cfg.blockDim = dim3(32 * LN_WARPS, 1, 1);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  const int pdl = s3r_pdl_enabled();
  if (pdl) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  S3R_CUDA_CHECK(cudaLaunchKernelEx(&cfg, s3r_layernorm_kernel, (const __nv_bfloat16*)x, (const __nv_bfloat16*)weight,
                                    (const __nv_bfloat16*)bias, (__nv_bfloat16*)y, (int)M, (int)C, (long long)ldx, eps, pdl));
  return S3R_OK;
}
