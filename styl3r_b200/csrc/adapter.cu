// Fused head epilogue -> Gaussians: replaces ~15 elementwise ATen kernels and the materialised rearranges of
// src/model/encoder/encoder_noposplat_multi_token_style.py:178-251, heads/postprocess.py:45-61 (exp-depth),
// common/gaussian_adapter.py:122-153 (UnifiedGaussianAdapter) and common/gaussians.py:8-44.
//
// Inputs are the planar NCHW head outputs of one view batch: pts_raw [B,3,HW], params [B,8,HW] (density logit,
// 3 scale logits, 4 quaternion xyzw), app [B,3*d_sh,HW].  Outputs are written at Gaussian index
// (b, view*HW + pixel) of the scene-level buffers [B, G, ...] (G = n_views*HW), i.e. directly in the layout the
// rasterizer reads:
//   means = xyz/|xyz| * expm1(|xyz|);  opacity = 0.5*(1 - (1-p)^e + p^(1/e)), p = sigmoid(logit)
//   scales = min(0.001*softplus(s), 0.3);  q = q/(|q| + 1e-8);  R from xyzw with 2/(q.q + 1e-8);  cov = R S S^T R^T
//   harmonics[c][k] = app[c*d_sh + k] * sh_mask[k]
#include "s3r_common.cuh"

// Element (batch b, channel c, pixel p) of a head output sits at  b*sb + c*sc + p*sp  (planar NCHW: sb = C*HW, sc = HW,
// sp = 1; pixel-major NHWC rows of pitch ld, as the tcgen05 1x1-conv GEMM writes them: sb = HW*ld, sc = 1, sp = ld).
struct HeadStrides {
  long long sb, sc, sp;
};

__global__ void __launch_bounds__(256) s3r_gaussian_adapter_kernel(
    const float* __restrict__ pts_raw, const float* __restrict__ params, const float* __restrict__ app,
    HeadStrides st_pts, HeadStrides st_prm, HeadStrides st_app,
    const float* __restrict__ sh_mask, int B, int HW, int d_sh, int view, int G, float exponent,
    float* __restrict__ means, float* __restrict__ cov, float* __restrict__ harm, float* __restrict__ opac,
    float* __restrict__ scales_out, float* __restrict__ rot_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * HW) return;
  const int b = (int)(i / HW), p = (int)(i - (long long)b * HW);
  const float* pr = pts_raw + b * st_pts.sb + p * st_pts.sp;
  const float* ga = params + b * st_prm.sb + p * st_prm.sp;
  const long long pc = st_pts.sc, gc = st_prm.sc, ac = st_app.sc;
  const size_t g = (size_t)b * G + (size_t)view * HW + p;
  // exp-depth point map
  const float x = pr[0], y = pr[pc], z = pr[2 * pc];
  const float d = sqrtf(x * x + y * y + z * z);
  const float k = expm1f(d) / fmaxf(d, 1e-8f);
  means[g * 3 + 0] = x * k;
  means[g * 3 + 1] = y * k;
  means[g * 3 + 2] = z * k;
  // opacity
  const float pdf = 1.0f / (1.0f + expf(-ga[0]));
  float o = pdf;
  if (exponent != 1.0f) o = 0.5f * (1.0f - powf(1.0f - pdf, exponent) + powf(pdf, 1.0f / exponent));
  opac[g] = o;
  // scales
  float s[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const float v = ga[(1 + c) * gc];
    const float sp = v > 20.f ? v : log1pf(expf(v));  // torch softplus (beta 1, threshold 20)
    s[c] = fminf(0.001f * sp, 0.3f);
  }
  // rotation
  float q[4];
#pragma unroll
  for (int c = 0; c < 4; c++) q[c] = ga[(4 + c) * gc];
  const float qn = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]) + 1e-8f;
#pragma unroll
  for (int c = 0; c < 4; c++) q[c] = q[c] / qn;
  const float qi = q[0], qj = q[1], qk = q[2], qr = q[3];
  const float two_s = 2.0f / ((qi * qi + qj * qj + qk * qk + qr * qr) + 1e-8f);
  const float R[9] = {1 - two_s * (qj * qj + qk * qk), two_s * (qi * qj - qk * qr), two_s * (qi * qk + qj * qr),
                      two_s * (qi * qj + qk * qr), 1 - two_s * (qi * qi + qk * qk), two_s * (qj * qk - qi * qr),
                      two_s * (qi * qk - qj * qr), two_s * (qj * qk + qi * qr), 1 - two_s * (qi * qi + qj * qj)};
  // cov = (R S)(R S)^T
  float M[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) M[3 * r + c] = R[3 * r + c] * s[c];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++)
      cov[g * 9 + 3 * r + c] = (M[3 * r] * s[0]) * R[3 * c] + (M[3 * r + 1] * s[1]) * R[3 * c + 1] +
                               (M[3 * r + 2] * s[2]) * R[3 * c + 2];
  // harmonics [3, d_sh]
  const float* ap = app + b * st_app.sb + p * st_app.sp;
  for (int c = 0; c < 3; c++)
    for (int kk = 0; kk < d_sh; kk++) harm[(g * 3 + c) * d_sh + kk] = ap[(c * d_sh + kk) * ac] * sh_mask[kk];
  if (scales_out) {
#pragma unroll
    for (int c = 0; c < 3; c++) scales_out[g * 3 + c] = s[c];
  }
  if (rot_out) {
#pragma unroll
    for (int c = 0; c < 4; c++) rot_out[g * 4 + c] = q[c];
  }
}

extern "C" int s3r_gaussian_adapter(const float* pts_raw, const float* params, const float* app, const float* sh_mask,
                                    int32_t B, int32_t HW, int32_t d_sh, int32_t view, int32_t G, float exponent,
                                    float* means, float* cov, float* harmonics, float* opacities, float* scales,
                                    float* rotations, void* stream) {
  if (B < 0 || HW < 0 || d_sh <= 0 || view < 0 || (long long)(view + 1) * HW > G) return S3R_ERR_INVALID_ARG;
  if (B == 0 || HW == 0) return S3R_OK;
  if (!pts_raw || !params || !app || !sh_mask || !means || !cov || !harmonics || !opacities) return S3R_ERR_INVALID_ARG;
  const long long n = (long long)B * HW;
  const HeadStrides sp{3LL * HW, HW, 1}, sg{8LL * HW, HW, 1}, sa{3LL * d_sh * HW, HW, 1};
  s3r_gaussian_adapter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      pts_raw, params, app, sp, sg, sa, sh_mask, B, HW, d_sh, view, G, exponent, means, cov, harmonics, opacities,
      scales, rotations);
  S3R_CUDA_CHECK(cudaGetLastError());
  return S3R_OK;
}

extern "C" int s3r_gaussian_adapter_nhwc(const float* pts_raw, const float* params, const float* app, int32_t ld_pts,
                                         int32_t ld_params, int32_t ld_app, const float* sh_mask, int32_t B, int32_t HW,
                                         int32_t d_sh, int32_t view, int32_t G, float exponent, float* means, float* cov,
                                         float* harmonics, float* opacities, float* scales, float* rotations,
                                         void* stream) {
  if (B < 0 || HW < 0 || d_sh <= 0 || view < 0 || (long long)(view + 1) * HW > G) return S3R_ERR_INVALID_ARG;
  if (ld_pts < 3 || ld_params < 8 || ld_app < 3 * d_sh) return S3R_ERR_INVALID_ARG;
  if (B == 0 || HW == 0) return S3R_OK;
  if (!pts_raw || !params || !app || !sh_mask || !means || !cov || !harmonics || !opacities) return S3R_ERR_INVALID_ARG;
  const long long n = (long long)B * HW;
  const HeadStrides sp{(long long)HW * ld_pts, 1, ld_pts}, sg{(long long)HW * ld_params, 1, ld_params},
      sa{(long long)HW * ld_app, 1, ld_app};
  s3r_gaussian_adapter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      pts_raw, params, app, sp, sg, sa, sh_mask, B, HW, d_sh, view, G, exponent, means, cov, harmonics, opacities,
      scales, rotations);
  S3R_CUDA_CHECK(cudaGetLastError());
  return S3R_OK;
}
