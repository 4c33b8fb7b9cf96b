// LayerNorm over the last dimension for the bf16 ViT trunks (nn.LayerNorm(dim, eps=1e-6) of croco/blocks.py:140-147,
// 206-217 and enc_norm / dec_norm, croco.py:34): y = (x - mean) * rsqrt(var + eps) * w + b, statistics and arithmetic in
// fp32, bf16 in / out.  One warp per row, the whole row in registers (C <= 2048): 16-byte loads / stores, two-pass
// variance (no cancellation), no shared memory.  HBM-bound: 4*C bytes per row.  Launched with programmatic dependent
// launch so that its (tiny) prologue overlaps the tail of the GEMM that produced x.
#include <cuda_bf16.h>

#include "s3r_common.cuh"

#define LN_WARPS 4
#define LN_MAX_VEC 8  // 8 x (32 lanes x 8 bf16) = 2048 channels

// NV: compile-time number of 256-channel vectors per row (3 = 768, 4 = 1024 channels: the two trunk widths) so that only
// the registers the row needs are allocated (95 -> ~60: 8 instead of 5 CTAs per SM), 0 = generic (C <= 2048).  The affine
// parameters are requested together with the row, not after the two reductions (one dependent memory round trip less).
template <typename Tin, int NV>
__global__ void __launch_bounds__(32 * LN_WARPS) s3r_layernorm_kernel(const Tin* __restrict__ x,
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
  constexpr int NMAX = NV ? NV : LN_MAX_VEC;
  const int nvec = NV ? NV : C / 256;  // full 32-lane x 8-element vectors per row (C % 256 == 0)
  float v[NMAX][8];
  float sum = 0.f;
  uint4 uwv[NMAX], ubv[NMAX];
  if (NV) {
#pragma unroll
    for (int i = 0; i < NMAX; i++) {
      uwv[i] = __ldg(reinterpret_cast<const uint4*>(w) + i * 32 + lane);
      ubv[i] = __ldg(reinterpret_cast<const uint4*>(b) + i * 32 + lane);
    }
  }
  if (sizeof(Tin) == 4) {  // fp32 residual stream: two 16-byte loads per 8 elements
    const float4* xr = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + (long long)row * ldx);
#pragma unroll
    for (int i = 0; i < NMAX; i++) {
      if (i < nvec) {
        const float4 a = xr[(i * 32 + lane) * 2], b2 = xr[(i * 32 + lane) * 2 + 1];
        v[i][0] = a.x, v[i][1] = a.y, v[i][2] = a.z, v[i][3] = a.w;
        v[i][4] = b2.x, v[i][5] = b2.y, v[i][6] = b2.z, v[i][7] = b2.w;
        sum += (a.x + a.y) + (a.z + a.w) + (b2.x + b2.y) + (b2.z + b2.w);
      }
    }
  } else {
    const uint4* xr = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(x) + (long long)row * ldx);
#pragma unroll
    for (int i = 0; i < NMAX; i++) {
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
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NMAX; i++) {
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
  for (int i = 0; i < NMAX; i++) {
    if (i < nvec) {
      const uint4 uw = NV ? uwv[i] : wr[i * 32 + lane], ub = NV ? ubv[i] : br[i * 32 + lane];
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

template <typename Tin>
static int launch_layernorm(const Tin* x, const void* weight, const void* bias, void* y, int32_t M, int32_t C, int64_t ldx,
                            float eps, void* stream) {
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
  auto kern = C == 1024 ? s3r_layernorm_kernel<Tin, 4> : (C == 768 ? s3r_layernorm_kernel<Tin, 3> : s3r_layernorm_kernel<Tin, 0>);
  S3R_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, x, (const __nv_bfloat16*)weight,
                                    (const __nv_bfloat16*)bias, (__nv_bfloat16*)y, (int)M, (int)C, (long long)ldx, eps, pdl));
  return S3R_OK;
}

extern "C" int s3r_layernorm_bf16(const void* x, const void* weight, const void* bias, void* y, int32_t M, int32_t C,
                                  int64_t ldx, float eps, void* stream) {
  return launch_layernorm((const __nv_bfloat16*)x, weight, bias, y, M, C, ldx, eps, stream);
}

extern "C" int s3r_layernorm_f32_bf16(const float* x, const void* weight, const void* bias, void* y, int32_t M, int32_t C,
                                      int64_t ldx, float eps, void* stream) {
  return launch_layernorm(x, weight, bias, y, M, C, ldx, eps, stream);
}

// ------------------------------------------------------------------------------------------------ backward
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * w,  xhat = (x - mean) * rstd   (statistics recomputed
// from x: the forward keeps nothing), dweight += sum_rows dy * xhat, dbias += sum_rows dy (fp32, atomically accumulated:
// the caller zero-fills or lets them add into .grad).  One warp per row, rows strided over a persistent grid so that the
// per-lane column partials of dweight / dbias stay in registers for the warp's whole share; they are reduced across the
// CTA's warps in shared memory and leave with one atomicAdd per column per CTA.  C % 256 == 0, C <= 1024.
#define LNB_WARPS 4
#define LNB_MAX_VEC 4

__global__ void __launch_bounds__(32 * LNB_WARPS) s3r_layernorm_bwd_kernel(
    const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ dy,
    __nv_bfloat16* __restrict__ dx, float* __restrict__ dweight, float* __restrict__ dbias, int M, int C, long long ldx,
    float eps) {
  __shared__ float s_red[LNB_WARPS][2][32 * 8];  // one 256-column vector group at a time
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = C / 256;
  float accw[LNB_MAX_VEC][8], accb[LNB_MAX_VEC][8];
#pragma unroll
  for (int i = 0; i < LNB_MAX_VEC; i++)
#pragma unroll
    for (int t = 0; t < 8; t++) accw[i][t] = accb[i][t] = 0.f;
  float wv[LNB_MAX_VEC][8];
#pragma unroll
  for (int i = 0; i < LNB_MAX_VEC; i++) {
    if (i < nvec) {
      const uint4 u = reinterpret_cast<const uint4*>(w)[i * 32 + lane];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const float2 f = __bfloat1622float2(h[t]);
        wv[i][2 * t] = f.x, wv[i][2 * t + 1] = f.y;
      }
    }
  }
  for (int row = blockIdx.x * LNB_WARPS + warp; row < M; row += gridDim.x * LNB_WARPS) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + (long long)row * ldx);
    const uint4* gr = reinterpret_cast<const uint4*>(dy + (long long)row * C);
    float v[LNB_MAX_VEC][8], g[LNB_MAX_VEC][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < LNB_MAX_VEC; i++) {
      if (i < nvec) {
        const uint4 u = xr[i * 32 + lane], ug = gr[i * 32 + lane];
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
        const __nv_bfloat162* hg = reinterpret_cast<const __nv_bfloat162*>(&ug);
#pragma unroll
        for (int t = 0; t < 4; t++) {
          const float2 f = __bfloat1622float2(h[t]), fg = __bfloat1622float2(hg[t]);
          v[i][2 * t] = f.x, v[i][2 * t + 1] = f.y;
          g[i][2 * t] = fg.x, g[i][2 * t + 1] = fg.y;
          sum += f.x + f.y;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < LNB_MAX_VEC; i++)
      if (i < nvec)
#pragma unroll
        for (int t = 0; t < 8; t++) {
          const float d = v[i][t] - mean;
          sq += d * d;
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)C + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LNB_MAX_VEC; i++)
      if (i < nvec)
#pragma unroll
        for (int t = 0; t < 8; t++) {
          const float xh = (v[i][t] - mean) * rstd;
          accw[i][t] += g[i][t] * xh;
          accb[i][t] += g[i][t];
          const float gw = g[i][t] * wv[i][t];
          v[i][t] = xh;
          g[i][t] = gw;
          s1 += gw;
          s2 += gw * xh;
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 /= (float)C;
    s2 /= (float)C;
    uint4* dr = reinterpret_cast<uint4*>(dx + (long long)row * C);
#pragma unroll
    for (int i = 0; i < LNB_MAX_VEC; i++) {
      if (i < nvec) {
        uint4 o;
        __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int t = 0; t < 4; t++)
          ho[t] = __floats2bfloat162_rn(rstd * (g[i][2 * t] - s1 - v[i][2 * t] * s2),
                                        rstd * (g[i][2 * t + 1] - s1 - v[i][2 * t + 1] * s2));
        dr[i * 32 + lane] = o;
      }
    }
  }
  if (!dweight && !dbias) return;
  // column partials: warps -> shared memory -> one atomicAdd per column per CTA
  for (int i = 0; i < nvec; i++) {
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 8; t++) {
      // the value of accw[i][t] for a compile-time i: select without dynamic register indexing
      float aw = 0.f, ab = 0.f;
#pragma unroll
      for (int ii = 0; ii < LNB_MAX_VEC; ii++)
        if (ii == i) aw = accw[ii][t], ab = accb[ii][t];
      s_red[warp][0][lane * 8 + t] = aw;
      s_red[warp][1][lane * 8 + t] = ab;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 256; c += 32 * LNB_WARPS) {
      float aw = 0.f, ab = 0.f;
#pragma unroll
      for (int k = 0; k < LNB_WARPS; k++) aw += s_red[k][0][c], ab += s_red[k][1][c];
      // column of (vector group i, lane l, element t) = i * 256 + l * 8 + t = i * 256 + c
      if (dweight) atomicAdd(dweight + i * 256 + c, aw);
      if (dbias) atomicAdd(dbias + i * 256 + c, ab);
    }
  }
}

extern "C" int s3r_layernorm_bwd_bf16(const void* x, const void* weight, const void* dy, void* dx, float* dweight,
                                      float* dbias, int32_t M, int32_t C, int64_t ldx, float eps, void* stream) {
  if (M < 0 || C <= 0 || ldx < C) return S3R_ERR_INVALID_ARG;
  if (M == 0) return S3R_OK;
  if (!x || !weight || !dy || !dx) return S3R_ERR_INVALID_ARG;
  if (C % 256 || C > 256 * LNB_MAX_VEC || ldx % 8 || (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)weight) & 15))
    return S3R_ERR_UNSUPPORTED;
  int grid = (M + LNB_WARPS - 1) / LNB_WARPS;
  if (grid > 148 * 4) grid = 148 * 4;
  s3r_layernorm_bwd_kernel<<<grid, 32 * LNB_WARPS, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)weight, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, dweight, dbias,
      (int)M, (int)C, (long long)ldx, eps);
  S3R_CUDA_CHECK(cudaGetLastError());
  return S3R_OK;
}
