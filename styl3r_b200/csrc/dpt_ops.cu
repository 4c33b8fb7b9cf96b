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

// s3r_im2col7x7_bf16: patch matrix of the 7x7 / pad 3 image-skip convolution of the Gaussian-parameter head
// (dpt_gs_head.py:113-118, `input_merger`: Conv2d(3, 256, 7, 1, 3) + ReLU) in ONE pass: img [B,3,H,W] bf16 (planar) ->
// cols [B*H*W, 152] bf16, column k < 147 = (ci, kh, kw) like F.unfold, columns 147..151 zero (the GEMM needs K % 8 == 0).
// Replaces F.unfold + transpose copy + F.pad (three passes over a 147-column matrix: 1.4 ms for 12 images) with one
// write-only pass; one thread per (pixel, 8 columns): one 16-byte store, 8 cached 2-byte loads.
__global__ void __launch_bounds__(256) s3r_im2col7x7_kernel(const __nv_bfloat16* __restrict__ img, uint4* __restrict__ cols,
                                                            int B, int H, int W) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * H * W * 19;
  if (i >= total) return;
  const int chunk = (int)(i % 19);
  const long long pix = i / 19;
  const int x = (int)(pix % W);
  const int y = (int)((pix / W) % H);
  const int b = (int)(pix / ((long long)W * H));
  const __nv_bfloat16* im = img + (size_t)b * 3 * H * W;
  __nv_bfloat16 v[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int k = chunk * 8 + j;
    __nv_bfloat16 val = __float2bfloat16_rn(0.0f);
    if (k < 147) {
      const int ci = k / 49, r = k - ci * 49;
      const int kh = r / 7, kw = r - kh * 7;
      const int yy = y + kh - 3, xx = x + kw - 3;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) val = im[((size_t)ci * H + yy) * W + xx];
    }
    v[j] = val;
  }
  cols[i] = *reinterpret_cast<const uint4*>(v);
}

extern "C" int s3r_im2col7x7_bf16(const void* img, void* cols, int32_t B, int32_t H, int32_t W, void* stream) {
  if (B < 0 || H <= 0 || W <= 0) return S3R_ERR_INVALID_ARG;
  if (B == 0) return S3R_OK;
  if (!img || !cols) return S3R_ERR_INVALID_ARG;
  if ((uintptr_t)cols & 15) return S3R_ERR_UNSUPPORTED;
  const long long total = (long long)B * H * W * 19;
  if ((total + 255) / 256 > 0x7fffffffLL) return S3R_ERR_UNSUPPORTED;
  s3r_im2col7x7_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)img,
                                                                                            (uint4*)cols, B, H, W);
  S3R_CUDA_CHECK(cudaGetLastError());
  return S3R_OK;
}

// s3r_gather_other_views_bf16: the cross-view context of the second decoder (AsymmetricCroCoMulti._decoder,
// backbone_croco_multiview.py:170-178): for every view i >= 1 the tokens of all OTHER views in view order, built straight
// from the two tensors the decoder branches keep - x0 [b, l, c] (view 0) and x1 [b, v-1, l, c] (views 1..v-1) ->
// ctx [b, v-1, (v-1)*l, c].  One 16-byte vector per thread (c % 8 == 0); replaces cat + index + two reshape copies per layer.
__global__ void __launch_bounds__(256) s3r_gather_other_views_kernel(const uint4* __restrict__ x0, const uint4* __restrict__ x1,
                                                                     uint4* __restrict__ ctx, int b, int v, int l, int c8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)b * (v - 1) * (v - 1) * l * c8;
  if (i >= total) return;
  const int cc = (int)(i % c8);
  long long t = i / c8;
  const int tok = (int)(t % l);
  t /= l;
  const int slot = (int)(t % (v - 1));   // position inside the context of this query view
  t /= (v - 1);
  const int qi = (int)(t % (v - 1));     // query view - 1
  const int bi = (int)(t / (v - 1));
  const int src = slot < qi + 1 ? slot : slot + 1;  // source view: the views != qi + 1 in ascending order
  const uint4 val = src == 0 ? x0[((size_t)bi * l + tok) * c8 + cc]
                             : x1[(((size_t)bi * (v - 1) + (src - 1)) * l + tok) * c8 + cc];
  ctx[i] = val;
}

extern "C" int s3r_gather_other_views_bf16(const void* x0, const void* x1, void* ctx, int32_t b, int32_t v, int32_t l,
                                           int32_t c, void* stream) {
  if (b < 0 || v < 2 || l <= 0 || c <= 0) return S3R_ERR_INVALID_ARG;
  if (b == 0) return S3R_OK;
  if (!x0 || !x1 || !ctx) return S3R_ERR_INVALID_ARG;
  if (c % 8 || (((uintptr_t)x0 | (uintptr_t)x1 | (uintptr_t)ctx) & 15)) return S3R_ERR_UNSUPPORTED;
  const long long total = (long long)b * (v - 1) * (v - 1) * l * (c / 8);
  s3r_gather_other_views_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)x0, (const uint4*)x1, (uint4*)ctx, b, v, l, c / 8);
  S3R_CUDA_CHECK(cudaGetLastError());
  return S3R_OK;
}
