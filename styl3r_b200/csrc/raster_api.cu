// C-ABI entry points of the rasterizer (see include/styl3r_b200.h).  Replaces the torch-extension surface
// of the third-party rasterizer driven by src/model/decoder/cuda_splatting.py:101-129: one call renders all
// views of a batch; no host synchronisation, no allocation, no torch types.
#include <string.h>

#include "s3r_common.cuh"

extern "C" int s3r_abi_version(void) { return S3R_ABI_VERSION; }

extern "C" const char* s3r_error_string(int code) {
  switch (code) {
    case S3R_OK: return "ok";
    case S3R_ERR_INVALID_ARG: return "invalid argument";
    case S3R_ERR_UNSUPPORTED: return "unsupported shape (compiled limits: <=255 tiles per axis, <=4096 tiles, SH degree <=3)";
    case S3R_ERR_STATE_TOO_SMALL: return "state buffer too small";
    case S3R_ERR_CUDA: return "CUDA runtime error";
    case S3R_ERR_NO_DEVICE: return "no CUDA device";
    default: return "unknown error";
  }
}

extern "C" int s3r_raster_layout_query(int32_t n_views, int32_t P, int32_t width, int32_t height, int64_t capacity,
                                       s3r_raster_layout* out) {
  if (!out || n_views <= 0 || P <= 0 || width <= 0 || height <= 0 || capacity < 0) return S3R_ERR_INVALID_ARG;
  memset(out, 0, sizeof(*out));
  const int tx = (width + S3R_TILE - 1) / S3R_TILE, ty = (height + S3R_TILE - 1) / S3R_TILE;
  if (tx > 255 || ty > 255 || (int64_t)tx * ty > S3R_MAX_TILES) return S3R_ERR_UNSUPPORTED;
  if (capacity >= (int64_t)1 << 32) return S3R_ERR_UNSUPPORTED;
  const int64_t tiles = (int64_t)tx * ty, chunks = ((int64_t)P + S3R_CHUNK - 1) / S3R_CHUNK;
  const int64_t nvP = (int64_t)n_views * P, nvT = (int64_t)n_views * tiles, HW = (int64_t)width * height;
  const int64_t cap = capacity > 0 ? capacity : 1;
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    int64_t o = off;
    off = s3r_align_up(off + bytes, 256);
    return o;
  };
  out->status = take(4 * 8);
  out->counters = take(8 * 4);
  out->depths = take(nvP * 4);
  out->xy = take(nvP * 8);
  out->conic_opacity = take(nvP * 16);
  out->rgb = take(nvP * 16);
  out->rect = take(nvP * 4);
  out->chunk_hist = take((int64_t)n_views * chunks * tiles * 2);
  out->chunk_base = take((int64_t)n_views * chunks * tiles * 4);
  out->tile_count = take(nvT * 4);
  out->ranges = take(nvT * 8);
  out->keys_unsorted = take(cap * 8);
  out->keys_tmp = take(cap * 8);
  out->point_list = take(cap * 4);
  out->point_keys = take(cap * 8);
  out->records = take(cap * S3R_REC_BYTES);
  out->final_T = take((int64_t)n_views * HW * 4);
  out->n_contrib = take((int64_t)n_views * HW * 4);
  out->grecords = take(nvP * S3R_REC_BYTES);
  out->work_order = take(nvT * 4);
  out->blists = take(cap * 8 * 4);
  out->bcounts = take(nvT * 8 * 4);
  out->n_contrib_blk = take((int64_t)n_views * HW * 4);
  out->total_bytes = off;
  out->tiles_x = tx;
  out->tiles_y = ty;
  out->tiles = (int32_t)tiles;
  out->chunks = (int32_t)chunks;
  return S3R_OK;
}

static int validate(const s3r_raster_params* p) {
  if (!p) return S3R_ERR_INVALID_ARG;
  if (p->n_views <= 0 || p->n_sets <= 0 || p->P <= 0 || p->width <= 0 || p->height <= 0) return S3R_ERR_INVALID_ARG;
  if (!p->means3D || !p->cov3D || !p->opacities || !p->viewmatrix || !p->projmatrix || !p->tanfov || !p->background)
    return S3R_ERR_INVALID_ARG;
  if ((p->shs == nullptr) == (p->colors_precomp == nullptr)) return S3R_ERR_INVALID_ARG;
  if (p->cov_stride != 6 && p->cov_stride != 9) return S3R_ERR_INVALID_ARG;
  if (p->shs) {
    if (p->sh_degree < 0 || p->sh_degree > 3) return S3R_ERR_UNSUPPORTED;
    if (p->sh_coeffs < (p->sh_degree + 1) * (p->sh_degree + 1)) return S3R_ERR_INVALID_ARG;
    if (p->sh_degree > 0 && !p->campos) return S3R_ERR_INVALID_ARG;
  }
  if (!p->view_set && p->n_sets != p->n_views) return S3R_ERR_INVALID_ARG;
  return S3R_OK;
}

extern "C" int s3r_raster_forward_stages(const s3r_raster_params* params, const s3r_raster_outputs* out, void* state,
                                         size_t state_bytes, int64_t capacity, uint32_t stage_mask, void* stream) {
  int rc = validate(params);
  if (rc != S3R_OK) return rc;
  if (!out || !out->color || !out->depth || !out->opacity || !out->radii || !state) return S3R_ERR_INVALID_ARG;
  s3r_raster_layout L;
  rc = s3r_raster_layout_query(params->n_views, params->P, params->width, params->height, capacity, &L);
  if (rc != S3R_OK) return rc;
  if ((int64_t)state_bytes < L.total_bytes) return S3R_ERR_STATE_TOO_SMALL;
  cudaStream_t st = (cudaStream_t)stream;
  char* s = (char*)state;
  if ((stage_mask & S3R_STAGE_PREPROCESS) && (rc = s3r_launch_preprocess(*params, L, s, out->radii, st)) != S3R_OK)
    return rc;
  if ((stage_mask & S3R_STAGE_BIN) && (rc = s3r_launch_bin(*params, L, s, capacity, st)) != S3R_OK) return rc;
  if ((stage_mask & S3R_STAGE_SORT) && (rc = s3r_launch_sort(*params, L, s, st)) != S3R_OK) return rc;
  if ((stage_mask & S3R_STAGE_BLEND) && (rc = s3r_launch_blend(*params, *out, L, s, st)) != S3R_OK) return rc;
  return S3R_OK;
}

extern "C" int s3r_raster_forward(const s3r_raster_params* params, const s3r_raster_outputs* out, void* state,
                                  size_t state_bytes, int64_t capacity, void* stream) {
  return s3r_raster_forward_stages(params, out, state, state_bytes, capacity, S3R_STAGE_ALL, stream);
}

extern "C" int s3r_raster_read_status(const void* state, int64_t host_out[4], void* stream) {
  if (!state || !host_out) return S3R_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  S3R_CUDA_CHECK(cudaMemcpyAsync(host_out, state, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  S3R_CUDA_CHECK(cudaStreamSynchronize(st));
  return S3R_OK;
}
