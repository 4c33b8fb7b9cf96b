// Rasterizer backward (placeholder until the kernels land).
#include "s3r_common.cuh"

extern "C" size_t s3r_raster_backward_scratch_bytes(int32_t n_views, int32_t P) {
  return (size_t)n_views * P * 12 * sizeof(float);
}

extern "C" int s3r_raster_backward(const s3r_raster_params*, const void*, size_t, int64_t, const s3r_raster_grads*,
                                   void*) {
  return S3R_ERR_UNSUPPORTED;
}
