// Camera set-up of render_cuda on device, one thread per view — replaces ~25 small torch kernels (batched LU
// inverse, einsum, acos, tan, matmul, …) and the two `.item()` host syncs per view of the reference
// (src/model/decoder/cuda_splatting.py:65-72,81-88,104-105; src/geometry/projection.py:247-261).
//
//   scale      = 1/near (scale-invariant) or 1
//   c2w'       = c2w with translation * scale;  near' = near*scale, far' = far*scale
//   fov        = angle between the rays K^-1 [0,.5,1] / [1,.5,1]  and  [.5,0,1] / [.5,1,1]
//   proj       = [[1/tanx,0,0,0],[0,1/tany,0,0],[0,0,f/(f-n),-fn/(f-n)],[0,0,1,0]]   (symmetric frustum)
//   outputs    = transpose(inverse(c2w')), transpose(proj), their product, campos, (tanx,tany), scale
// Internals run in fp64 and are rounded once to fp32 (the torch reference rounds after every op; parity is
// tolerance-level, ~1e-6 relative — the rasterizer's bit-exactness is defined for given camera matrices).
#include "s3r_common.cuh"

__device__ static bool inv4(const double* m, double* out) {
  double a[4][8];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      a[r][c] = m[4 * r + c];
      a[r][4 + c] = (r == c) ? 1.0 : 0.0;
    }
  for (int col = 0; col < 4; col++) {
    int piv = col;
    double best = fabs(a[col][col]);
    for (int r = col + 1; r < 4; r++)
      if (fabs(a[r][col]) > best) { best = fabs(a[r][col]); piv = r; }
    if (best == 0.0) return false;
    if (piv != col)
      for (int c = 0; c < 8; c++) { double t = a[col][c]; a[col][c] = a[piv][c]; a[piv][c] = t; }
    const double d = 1.0 / a[col][col];
    for (int c = 0; c < 8; c++) a[col][c] *= d;
    for (int r = 0; r < 4; r++) {
      if (r == col) continue;
      const double f = a[r][col];
      if (f != 0.0)
        for (int c = 0; c < 8; c++) a[r][c] -= f * a[col][c];
    }
  }
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) out[4 * r + c] = a[r][4 + c];
  return true;
}

__global__ void s3r_camera_setup_kernel(const float* __restrict__ extr, const float* __restrict__ intr,
                                        const float* __restrict__ near_, const float* __restrict__ far_,
                                        int scale_invariant, int input_is_w2c, int n, float* __restrict__ viewmatrix,
                                        float* __restrict__ projmatrix, float* __restrict__ projraw,
                                        float* __restrict__ campos, float* __restrict__ tanfov,
                                        float* __restrict__ scales) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = scale_invariant ? 1.0f / near_[i] : 1.0f;
  const double nr = (double)(near_[i] * s), fr = (double)(far_[i] * s);
  double e[16], w2c[16];
  if (input_is_w2c) {
    // pose-align loop: the world->camera matrix is carried directly.  Scaling the camera-to-world translation by s
    // equals scaling the world->camera translation by s; the camera centre is -R^T t.
    for (int k = 0; k < 16; k++) w2c[k] = extr[16 * i + k];
    for (int r = 0; r < 3; r++) w2c[4 * r + 3] = (double)(extr[16 * i + 4 * r + 3] * s);
    for (int r = 0; r < 3; r++)
      e[4 * r + 3] = -(w2c[0 + r] * w2c[3] + w2c[4 + r] * w2c[7] + w2c[8 + r] * w2c[11]);
  } else {
    for (int k = 0; k < 16; k++) e[k] = extr[16 * i + k];
    for (int r = 0; r < 3; r++) e[4 * r + 3] = (double)(extr[16 * i + 4 * r + 3] * s);
    if (!inv4(e, w2c))
      for (int k = 0; k < 16; k++) w2c[k] = nan("");
  }
  // K^-1 (3x3 adjugate)
  const float* K = intr + 9 * i;
  const double a = K[0], b = K[1], c = K[2], d = K[3], ee = K[4], f = K[5], g = K[6], h = K[7], k9 = K[8];
  const double det = a * (ee * k9 - f * h) - b * (d * k9 - f * g) + c * (d * h - ee * g);
  const double id = 1.0 / det;
  const double Ki[9] = {(ee * k9 - f * h) * id, (c * h - b * k9) * id, (b * f - c * ee) * id,
                        (f * g - d * k9) * id,  (a * k9 - c * g) * id, (c * d - a * f) * id,
                        (d * h - ee * g) * id,  (b * g - a * h) * id,  (a * ee - b * d) * id};
  const double pts[4][3] = {{0, 0.5, 1}, {1, 0.5, 1}, {0.5, 0, 1}, {0.5, 1, 1}};
  double ray[4][3];
  for (int q = 0; q < 4; q++) {
    double v[3];
    for (int r = 0; r < 3; r++) v[r] = Ki[3 * r] * pts[q][0] + Ki[3 * r + 1] * pts[q][1] + Ki[3 * r + 2] * pts[q][2];
    const double nn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int r = 0; r < 3; r++) ray[q][r] = v[r] / nn;
  }
  const double fovx = acos(fmin(1.0, fmax(-1.0, ray[0][0] * ray[1][0] + ray[0][1] * ray[1][1] + ray[0][2] * ray[1][2])));
  const double fovy = acos(fmin(1.0, fmax(-1.0, ray[2][0] * ray[3][0] + ray[2][1] * ray[3][1] + ray[2][2] * ray[3][2])));
  const double tx = tan(0.5 * fovx), ty = tan(0.5 * fovy);
  // projection (row-major mathematical matrix)
  double P[16];
  for (int k2 = 0; k2 < 16; k2++) P[k2] = 0.0;
  P[0] = 1.0 / tx;
  P[5] = 1.0 / ty;
  P[10] = fr / (fr - nr);
  P[11] = -(fr * nr) / (fr - nr);
  P[14] = 1.0;
  // outputs in the transposed layout: out[4*col + row] = M[row][col]
  float vt[16], pt[16];
  for (int r = 0; r < 4; r++)
    for (int cc = 0; cc < 4; cc++) {
      vt[4 * cc + r] = (float)w2c[4 * r + cc];
      pt[4 * cc + r] = (float)P[4 * r + cc];
    }
  for (int k2 = 0; k2 < 16; k2++) {
    viewmatrix[16 * i + k2] = vt[k2];
    projraw[16 * i + k2] = pt[k2];
  }
  // full = view_T @ proj_T (fp32 product of the rounded factors, like the torch reference)
  for (int r = 0; r < 4; r++)
    for (int cc = 0; cc < 4; cc++) {
      float acc = 0.f;
      for (int k2 = 0; k2 < 4; k2++) acc = fmaf(vt[4 * r + k2], pt[4 * k2 + cc], acc);
      projmatrix[16 * i + 4 * r + cc] = acc;
    }
  for (int r = 0; r < 3; r++) campos[3 * i + r] = (float)e[4 * r + 3];
  tanfov[2 * i] = (float)tx;
  tanfov[2 * i + 1] = (float)ty;
  scales[i] = s;
}

extern "C" int s3r_camera_setup(const float* extrinsics, const float* intrinsics, const float* near_, const float* far_,
                                int32_t scale_invariant, int32_t input_is_w2c, int32_t n, float* viewmatrix, float* projmatrix,
                                float* projmatrix_raw, float* campos, float* tanfov, float* scales, void* stream) {
  if (n < 0) return S3R_ERR_INVALID_ARG;
  if (n == 0) return S3R_OK;
  if (!extrinsics || !intrinsics || !near_ || !far_ || !viewmatrix || !projmatrix || !projmatrix_raw || !campos ||
      !tanfov || !scales)
    return S3R_ERR_INVALID_ARG;
  s3r_camera_setup_kernel<<<(n + 31) / 32, 32, 0, (cudaStream_t)stream>>>(
      extrinsics, intrinsics, near_, far_, scale_invariant, input_is_w2c, n, viewmatrix, projmatrix, projmatrix_raw, campos, tanfov,
      scales);
  S3R_CUDA_CHECK(cudaGetLastError());
  return S3R_OK;
}
