// RoPE-2D, in place on tokens[B,N,H,D] — replaces the reference's curope extension
// (src/model/encoder/backbone/croco/curope/kernels.cu:17-108, curope.cpp:49-69) and mirrors the numerics of
// the PyTorch RoPE2D it falls back to (croco/pos_embed.py:112-159): features [0, D/2) rotate with the y
// position, [D/2, D) with the x position; inside a half the pairs are (d, d + D/4) and the angle is
// pos * fwd * base^(-d/(D/4)).  fwd = -1 applies the inverse rotation (= the backward pass).
//
// One thread owns VEC consecutive pair-lanes of one token-half and a slice of `hpt` heads (blockIdx.y), so
// sin/cos are shared by the heads of the slice and every access is a 16-byte vector; the head slices keep
// the grid large enough (>= ~64k threads) for the 257..4112-token shapes of the ViT.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "s3r_common.cuh"

template <typename T> struct Cvt;
template <> struct Cvt<float> {
  static __device__ __forceinline__ float to(float v) { return v; }
  static __device__ __forceinline__ float from(float v) { return v; }
};
template <> struct Cvt<__half> {
  static __device__ __forceinline__ float to(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from(float v) { return __float2half_rn(v); }
};
template <> struct Cvt<__nv_bfloat16> {
  static __device__ __forceinline__ float to(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from(float v) { return __float2bfloat16_rn(v); }
};

template <typename T, int VEC>
struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };

// threads: one per (token, half, group of VEC pairs); groups per half = Q / VEC
template <typename T, int VEC>
__global__ void __launch_bounds__(256) s3r_rope2d_kernel(T* __restrict__ tok, const long long* __restrict__ pos,
                                                        long long n_tokens, int N, int H, int Q, long long stride_b,
                                                        long long stride_n, long long stride_h, float base, float fwd,
                                                        int hpt) {
  const int gph = Q / VEC;  // groups per half
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long token = gid / (2 * gph);
  if (token >= n_tokens) return;
  const int r = (int)(gid - token * 2 * gph);
  const int half = r / gph, grp = r - half * gph;
  const long long b = token / N, n = token - b * N;
  const float p = (float)pos[token * 2 + half];
  float cs[VEC], sn[VEC];
#pragma unroll
  for (int k = 0; k < VEC; k++) {
    const int d = grp * VEC + k;
    const float inv_freq = fwd / powf(base, (float)d / (float)Q);
    sincosf(p * inv_freq, &sn[k], &cs[k]);
  }
  T* t0 = tok + b * stride_b + n * stride_n + (long long)half * 2 * Q + (long long)grp * VEC;
  using P = Pack<T, VEC>;
  const int h0 = blockIdx.y * hpt, h1 = min(H, h0 + hpt);
#pragma unroll 2
  for (int h = h0; h < h1; h++) {
    T* th = t0 + (long long)h * stride_h;
    P u = *reinterpret_cast<const P*>(th);
    P v = *reinterpret_cast<const P*>(th + Q);
    P uo, vo;
#pragma unroll
    for (int k = 0; k < VEC; k++) {
      const float uf = Cvt<T>::to(u.v[k]), vf = Cvt<T>::to(v.v[k]);
      uo.v[k] = Cvt<T>::from(uf * cs[k] - vf * sn[k]);
      vo.v[k] = Cvt<T>::from(vf * cs[k] + uf * sn[k]);
    }
    *reinterpret_cast<P*>(th) = uo;
    *reinterpret_cast<P*>(th + Q) = vo;
  }
}

template <typename T, int VEC>
static int launch(void* tokens, const int64_t* pos, int B, int N, int H, int Q, int64_t sb, int64_t sn, int64_t sh,
                  float base, float fwd, cudaStream_t st) {
  const long long n_tokens = (long long)B * N;
  const long long threads = n_tokens * 2 * (Q / VEC);
  const int block = 256;
  const long long grid = (threads + block - 1) / block;
  int hpt = H;  // heads per thread: shrink until the grid has >= 64k threads (or one head per thread)
  while (hpt > 1 && threads * ((H + hpt - 1) / hpt) < 65536) hpt = (hpt + 1) / 2;
  dim3 g((unsigned)grid, (unsigned)((H + hpt - 1) / hpt));
  s3r_rope2d_kernel<T, VEC><<<g, block, 0, st>>>((T*)tokens, (const long long*)pos, n_tokens, N, H, Q, sb, sn, sh, base,
                                                 fwd, hpt);
  S3R_CUDA_CHECK(cudaGetLastError());
  return S3R_OK;
}

template <typename T, int VMAX>
static int dispatch_vec(void* tokens, const int64_t* pos, int B, int N, int H, int D, int64_t sb, int64_t sn,
                        int64_t sh, float base, float fwd, cudaStream_t st) {
  const int Q = D / 4;
  const bool aligned = ((uintptr_t)tokens % (sizeof(T) * VMAX) == 0) && (Q % VMAX == 0) && (sb % VMAX == 0) &&
                       (sn % VMAX == 0) && (sh % VMAX == 0);
  if (aligned) return launch<T, VMAX>(tokens, pos, B, N, H, Q, sb, sn, sh, base, fwd, st);
  return launch<T, 1>(tokens, pos, B, N, H, Q, sb, sn, sh, base, fwd, st);
}

extern "C" int s3r_rope2d(void* tokens, const int64_t* pos, int32_t B, int32_t N, int32_t H, int32_t D,
                          int64_t stride_b, int64_t stride_n, int64_t stride_h, float base, float fwd, int32_t dtype,
                          void* stream) {
  if (B < 0 || N < 0 || H < 0 || D <= 0) return S3R_ERR_INVALID_ARG;
  if (D % 4 != 0) return S3R_ERR_INVALID_ARG;  // same contract as curope.cpp:58
  if (B == 0 || N == 0 || H == 0) return S3R_OK;
  if (!tokens || !pos) return S3R_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case 0: return dispatch_vec<float, 4>(tokens, pos, B, N, H, D, stride_b, stride_n, stride_h, base, fwd, st);
    case 1: return dispatch_vec<__half, 8>(tokens, pos, B, N, H, D, stride_b, stride_n, stride_h, base, fwd, st);
    case 2: return dispatch_vec<__nv_bfloat16, 8>(tokens, pos, B, N, H, D, stride_b, stride_n, stride_h, base, fwd, st);
    default: return S3R_ERR_INVALID_ARG;
  }
}
