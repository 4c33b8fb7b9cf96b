// Input staging: the reference's `rescale_and_crop` (src/dataset/shims/crop_shim.py:11-79) +
// `normalize_image` (src/dataset/shims/normalize_shim.py:15-18) as two kernels, bit-exact with Pillow's 8-bit LANCZOS
// resize (third-party libImaging/Resample.c: separable passes, int32 fixed point with PRECISION_BITS = 22, uint8
// intermediate image; restated in oracle/resize_oracle.py).  The reference does this per image on the CPU through
// PIL (float -> uint8 -> PIL -> numpy -> float, one D2H + H2D round trip each).
//
//   pass 1 (horizontal): float planes [p, h_in, w_in] --quantise (x*255, clip, truncate)--> taps --> uint8 [p, h_in, w_s]
//   pass 2 (vertical)  : uint8 [p, h_in, w_s] --> taps --> crop window --> u8/255 (-> (x - mean)/std) --> float [p, h_out, w_out]
//
// Tap tables (bounds = (first, count), coefficients int32 [out, ksize]) are computed once per (in, out) size on the
// host in float64 exactly as Pillow's precompute_coeffs / normalize_coeffs_8bpc do.  HBM-bound: 4*h_in*w_in read +
// 2*h_in*w_s + 4*h_out*w_out per plane.
#include "s3r_common.cuh"

#define RS_PRECISION_BITS 22

__device__ __forceinline__ int rs_quant(float x) {  // torch: (image * 255).clip(0, 255).type(torch.uint8)
  const float v = fminf(fmaxf(x * 255.0f, 0.0f), 255.0f);
  return (int)v;
}
__device__ __forceinline__ int rs_clip8(int acc) {
  const int v = acc >> RS_PRECISION_BITS;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__global__ void __launch_bounds__(256) s3r_resize_h_kernel(const float* __restrict__ img, int planes, int h_in, int w_in,
                                                           int w_s, const int* __restrict__ bounds,
                                                           const int* __restrict__ kk, int ksize,
                                                           uint8_t* __restrict__ mid) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)planes * h_in * w_s) return;
  const int xx = (int)(i % w_s);
  const long long py = i / w_s;  // plane * h_in + y
  const int x0 = bounds[2 * xx], n = bounds[2 * xx + 1];
  const float* src = img + py * w_in + x0;
  const int* k = kk + (size_t)xx * ksize;
  int acc = 1 << (RS_PRECISION_BITS - 1);
  for (int x = 0; x < n; x++) acc += rs_quant(src[x]) * k[x];
  mid[i] = (uint8_t)rs_clip8(acc);
}

__global__ void __launch_bounds__(256) s3r_resize_v_kernel(const uint8_t* __restrict__ mid, int planes, int h_in, int w_s,
                                                           const int* __restrict__ bounds, const int* __restrict__ kk,
                                                           int ksize, int crop_row, int crop_col, int h_out, int w_out,
                                                           const float* __restrict__ mean, const float* __restrict__ stdv,
                                                           int channels, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)planes * h_out * w_out) return;
  const int xo = (int)(i % w_out);
  const int yo = (int)((i / w_out) % h_out);
  const int p = (int)(i / ((long long)w_out * h_out));
  const int yy = yo + crop_row, xx = xo + crop_col;
  const int y0 = bounds[2 * yy], n = bounds[2 * yy + 1];
  const uint8_t* src = mid + ((size_t)p * h_in + y0) * w_s + xx;
  const int* k = kk + (size_t)yy * ksize;
  int acc = 1 << (RS_PRECISION_BITS - 1);
  for (int y = 0; y < n; y++) acc += (int)src[(size_t)y * w_s] * k[y];
  float v = (float)((double)rs_clip8(acc) / 255.0);  // np.array(img) / 255 in float64, then the tensor dtype
  if (mean) v = (v - mean[p % channels]) / stdv[p % channels];
  out[i] = v;
}

extern "C" int s3r_rescale_crop(const float* img, int32_t planes, int32_t channels, int32_t h_in, int32_t w_in, int32_t h_s,
                                int32_t w_s, const int32_t* h_bounds, const int32_t* h_coeffs, int32_t h_ksize,
                                const int32_t* v_bounds, const int32_t* v_coeffs, int32_t v_ksize, int32_t crop_row,
                                int32_t crop_col, int32_t h_out, int32_t w_out, const float* mean, const float* stdv,
                                uint8_t* scratch, float* out, void* stream) {
  if (planes < 0 || channels <= 0 || h_in <= 0 || w_in <= 0 || h_s <= 0 || w_s <= 0 || h_out <= 0 || w_out <= 0 ||
      h_ksize <= 0 || v_ksize <= 0)
    return S3R_ERR_INVALID_ARG;
  if (crop_row < 0 || crop_col < 0 || crop_row + h_out > h_s || crop_col + w_out > w_s) return S3R_ERR_INVALID_ARG;
  if ((mean == nullptr) != (stdv == nullptr)) return S3R_ERR_INVALID_ARG;
  if (planes == 0) return S3R_OK;
  if (!img || !h_bounds || !h_coeffs || !v_bounds || !v_coeffs || !scratch || !out) return S3R_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const long long n1 = (long long)planes * h_in * w_s, n2 = (long long)planes * h_out * w_out;
  s3r_resize_h_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(img, planes, h_in, w_in, w_s, h_bounds, h_coeffs,
                                                                    h_ksize, scratch);
  S3R_CUDA_CHECK(cudaGetLastError());
  s3r_resize_v_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(scratch, planes, h_in, w_s, v_bounds, v_coeffs, v_ksize,
                                                                    crop_row, crop_col, h_out, w_out, mean, stdv,
                                                                    channels, out);
  S3R_CUDA_CHECK(cudaGetLastError());
  return S3R_OK;
}
