// Pose update on device: w2c_out[i] = SE3_exp([rho_i, theta_i]) @ w2c_in[i].
// Mirrors src/misc/cam_utils.py:67-137 (SO3_exp, V, SE3_exp, update_pose) — including the small-angle
// branches at |theta| < 1e-5 — without the per-view Python loop and its host synchronisations.
#include "s3r_common.cuh"

__global__ void s3r_se3_update_kernel(const float* __restrict__ w2c_in, const float* __restrict__ rho,
                                      const float* __restrict__ theta, float* __restrict__ w2c_out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float tx = theta[3 * i], ty = theta[3 * i + 1], tz = theta[3 * i + 2];
  const float Wm[9] = {0.f, -tz, ty, tz, 0.f, -tx, -ty, tx, 0.f};
  float W2[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) W2[3 * r + c] = Wm[3 * r] * Wm[c] + Wm[3 * r + 1] * Wm[3 + c] + Wm[3 * r + 2] * Wm[6 + c];
  const float angle = sqrtf(tx * tx + ty * ty + tz * tz);
  float a, b, c1, c2;  // R = I + a W + b W2 ; V = I + c1 W + c2 W2
  if (angle < 1e-5f) {
    a = 1.f; b = 0.5f; c1 = 0.5f; c2 = 1.0f / 6.0f;
  } else {
    const float s = sinf(angle), c = cosf(angle);
    a = s / angle;
    b = (1.f - c) / (angle * angle);
    c1 = b;
    c2 = (angle - s) / (angle * angle * angle);
  }
  float R[9], V[9];
#pragma unroll
  for (int k = 0; k < 9; k++) {
    const float id = (k == 0 || k == 4 || k == 8) ? 1.f : 0.f;
    R[k] = id + a * Wm[k] + b * W2[k];
    V[k] = id + c1 * Wm[k] + c2 * W2[k];
  }
  const float r0 = rho[3 * i], r1 = rho[3 * i + 1], r2 = rho[3 * i + 2];
  const float t[3] = {V[0] * r0 + V[1] * r1 + V[2] * r2, V[3] * r0 + V[4] * r1 + V[5] * r2,
                      V[6] * r0 + V[7] * r1 + V[8] * r2};
  const float* M = w2c_in + 16 * i;
  float* O = w2c_out + 16 * i;
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 4; c++)
      O[4 * r + c] = R[3 * r] * M[c] + R[3 * r + 1] * M[4 + c] + R[3 * r + 2] * M[8 + c] + t[r] * M[12 + c];
#pragma unroll
  for (int c = 0; c < 4; c++) O[12 + c] = M[12 + c];
}

extern "C" int s3r_se3_update_w2c(const float* w2c_in, const float* rho, const float* theta, float* w2c_out, int32_t n,
                                  void* stream) {
  if (!w2c_in || !rho || !theta || !w2c_out || n < 0) return S3R_ERR_INVALID_ARG;
  if (n == 0) return S3R_OK;
  s3r_se3_update_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(w2c_in, rho, theta, w2c_out, n);
  S3R_CUDA_CHECK(cudaGetLastError());
  return S3R_OK;
}
