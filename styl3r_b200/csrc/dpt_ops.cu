// Memory-bound helpers of the DPT heads in the bf16 NHWC inference layout (sm_100a).
//
// s3r_upsample2x_nhwc_bf16: bilinear x2 upsampling with align_corners=True - F.interpolate(scale_factor=2,
// mode="bilinear", align_corners=True) of heads/dpt_block.py:209-211 (FeatureFusionBlock), dpt_head.py:57 and
// dpt_gs_head.py:138 - with the optional residual add of dpt_gs_head.py:140 (`out + input_merger(img)`) fused.
// Same interpolation arithmetic as ATen's upsample_bilinear2d (fp32 weights, source index = dst * (in-1)/(out-1)).
// One thread per (output pixel, 8 channels): 16-byte loads/stores, fully coalesced along C.
#include <cuda_bf16.h>

#include "s3r_common.cuh"

__device__ __forceinline__ void bf8_to_f(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__global__ void __launch_bounds__(256) s3r_upsample2x_kernel(const uint4* __restrict__ x, const uint4* __restrict__ add,
                                                             uint4* __restrict__ y, int n, int h, int w, int c8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int oh = 2 * h, ow = 2 * w;
  const long long total = (long long)n * oh * ow * c8;
  if (i >= total) return;
  const int cb = (int)(i % c8);
  long long t = i / c8;
  const int ox = (int)(t % ow);
  t /= ow;
  const int oy = (int)(t % oh);
  const int b = (int)(t / oh);
  const float rh = oh > 1 ? (float)(h - 1) / (float)(oh - 1) : 0.f;
  const float rw = ow > 1 ? (float)(w - 1) / (float)(ow - 1) : 0.f;
  const float sy = rh * (float)oy, sx = rw * (float)ox;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly = sy - (float)y0, lx = sx - (float)x0;
  const float hy = 1.0f - ly, hx = 1.0f - lx;
  const size_t base = (size_t)b * h * w;
  float v00[8], v01[8], v10[8], v11[8];
  bf8_to_f(x[(base + (size_t)y0 * w + x0) * c8 + cb], v00);
  bf8_to_f(x[(base + (size_t)y0 * w + x1) * c8 + cb], v01);
  bf8_to_f(x[(base + (size_t)y1 * w + x0) * c8 + cb], v10);
  bf8_to_f(x[(base + (size_t)y1 * w + x1) * c8 + cb], v11);
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; j++) r[j] = hy * (hx * v00[j] + lx * v01[j]) + ly * (hx * v10[j] + lx * v11[j]);
  if (add) {
    float a[8];
    bf8_to_f(add[i], a);
    // the reference rounds the interpolated map to the tensor dtype before the add; keep that order
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = __bfloat162float(__float2bfloat16_rn(r[j])) + a[j];
  }
  uint4 o;
  __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int j = 0; j < 4; j++) po[j] = __floats2bfloat162_rn(r[2 * j], r[2 * j + 1]);
  y[i] = o;
}

extern "C" int s3r_upsample2x_nhwc_bf16(const void* x, const void* add, void* y, int32_t n, int32_t h, int32_t w,
                                        int32_t c, void* stream) {
  if (n < 0 || h <= 0 || w <= 0 || c <= 0) return S3R_ERR_INVALID_ARG;
  if (n == 0) return S3R_OK;
  if (!x || !y) return S3R_ERR_INVALID_ARG;
  if (c % 8 || (((uintptr_t)x | (uintptr_t)y | (uintptr_t)add) & 15)) return S3R_ERR_UNSUPPORTED;
  const long long total = (long long)n * 4 * h * w * (c / 8);
  s3r_upsample2x_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)x, (const uint4*)add, (uint4*)y, n, h, w, c / 8);
  S3R_CUDA_CHECK(cudaGetLastError());
  return S3R_OK;
}
