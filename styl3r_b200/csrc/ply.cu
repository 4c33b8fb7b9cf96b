// .ply export packing: the per-Gaussian vertex record of src/model/ply_export.py:26-74,
//   x y z | nx ny nz (= 0) | f_dc_0..2 | [f_rest_*] | opacity | scale_0..2 (log) | rot_0..3 (w x y z)
// written by one kernel straight into the binary little-endian row layout of the file, so the export is one D2H copy
// + one write instead of ~10 ATen ops, 7 host copies, a scipy quaternion round trip and a Python tuple per Gaussian.
//
// Rotation: the reference passes the xyzw quaternions through scipy `R.from_quat(q).as_matrix()` ->
// `R.from_matrix(m).as_quat()` (ply_export.py:46-49), i.e. normalisation plus a sign canonicalisation decided by the
// matrix -> quaternion branch (largest of m00, m11, m22, trace).  The same float64 arithmetic is restated here.
#include "s3r_common.cuh"

__global__ void __launch_bounds__(256) s3r_ply_pack_kernel(const float* __restrict__ means, const float* __restrict__ scales,
                                                           const float* __restrict__ rot, const float* __restrict__ harm,
                                                           const float* __restrict__ opac, const float* __restrict__ xform,
                                                           int n, int d_sh, int n_rest, float* __restrict__ out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const int row = 17 + n_rest;
  float* o = out + (size_t)g * row;
  float sx = 0.f, sy = 0.f, sz = 0.f, sf = 1.f;
  if (xform) sx = xform[0], sy = xform[1], sz = xform[2], sf = xform[3];
  // ply_export.py:36-43 (shift_and_scale): (mean - median) / factor, scale / factor - fp32 like the reference
  float m[3] = {means[g * 3 + 0] - sx, means[g * 3 + 1] - sy, means[g * 3 + 2] - sz};
  float s[3] = {scales[g * 3 + 0], scales[g * 3 + 1], scales[g * 3 + 2]};
  if (xform) {
#pragma unroll
    for (int c = 0; c < 3; c++) m[c] = m[c] / sf, s[c] = s[c] / sf;
  }
  o[0] = m[0], o[1] = m[1], o[2] = m[2];
  o[3] = o[4] = o[5] = 0.f;
  // harmonics [g][3][d_sh]: f_dc = [..., 0]; f_rest = [..., 1:] flattened (channel-major), ply_export.py:53-54
  for (int c = 0; c < 3; c++) o[6 + c] = harm[((size_t)g * 3 + c) * d_sh];
  if (n_rest) {
    for (int c = 0; c < 3; c++)
      for (int k = 1; k < d_sh; k++) o[9 + c * (d_sh - 1) + (k - 1)] = harm[((size_t)g * 3 + c) * d_sh + k];
  }
  float* t = o + 9 + n_rest;
  t[0] = opac[g];
  t[1] = logf(s[0]), t[2] = logf(s[1]), t[3] = logf(s[2]);
  // scipy from_quat (normalise) -> as_matrix -> from_matrix -> as_quat, float64
  double x = rot[g * 4 + 0], y = rot[g * 4 + 1], z = rot[g * 4 + 2], w = rot[g * 4 + 3];
  const double nrm = sqrt(x * x + y * y + z * z + w * w);
  x /= nrm, y /= nrm, z /= nrm, w /= nrm;
  const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w, xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z,
               xw = x * w;
  double M[3][3];
  M[0][0] = x2 - y2 - z2 + w2, M[1][0] = 2 * (xy + zw), M[2][0] = 2 * (xz - yw);
  M[0][1] = 2 * (xy - zw), M[1][1] = -x2 + y2 - z2 + w2, M[2][1] = 2 * (yz + xw);
  M[0][2] = 2 * (xz + yw), M[1][2] = 2 * (yz - xw), M[2][2] = -x2 - y2 + z2 + w2;
  const double dec[4] = {M[0][0], M[1][1], M[2][2], M[0][0] + M[1][1] + M[2][2]};
  int choice = 0;
#pragma unroll
  for (int i = 1; i < 4; i++)
    if (dec[i] > dec[choice]) choice = i;  // first maximum, like numpy argmax
  double q[4];
  if (choice != 3) {
    const int i = choice, j = (i + 1) % 3, k = (j + 1) % 3;
    q[i] = 1 - dec[3] + 2 * M[i][i];
    q[j] = M[j][i] + M[i][j];
    q[k] = M[k][i] + M[i][k];
    q[3] = M[k][j] - M[j][k];
  } else {
    q[0] = M[2][1] - M[1][2];
    q[1] = M[0][2] - M[2][0];
    q[2] = M[1][0] - M[0][1];
    q[3] = 1 + dec[3];
  }
  const double qn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  t[4] = (float)(q[3] / qn), t[5] = (float)(q[0] / qn), t[6] = (float)(q[1] / qn), t[7] = (float)(q[2] / qn);  // w x y z
}

extern "C" int s3r_ply_pack(const float* means, const float* scales, const float* rotations, const float* harmonics,
                            const float* opacities, const float* xform, int32_t n, int32_t d_sh, int32_t save_rest,
                            float* out, void* stream) {
  if (n < 0 || d_sh <= 0) return S3R_ERR_INVALID_ARG;
  if (n == 0) return S3R_OK;
  if (!means || !scales || !rotations || !harmonics || !opacities || !out) return S3R_ERR_INVALID_ARG;
  const int n_rest = save_rest ? 3 * (d_sh - 1) : 0;
  s3r_ply_pack_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(means, scales, rotations, harmonics, opacities,
                                                                        xform, n, d_sh, n_rest, out);
  S3R_CUDA_CHECK(cudaGetLastError());
  return S3R_OK;
}
