// Shared declarations of the styl3r_b200 CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/styl3r_b200.h"

#define S3R_CHUNK 256          // Gaussians per preprocess / emit CTA (= one bit-mask row of 8 words)
#define S3R_REC_FLOATS 12      // blend record: x y A B | C o r g | b depth ex ey
#define S3R_REC_BYTES 48
#ifndef S3R_SORT_SMEM_CAP
#define S3R_SORT_SMEM_CAP 4096 // per-tile instances sorted entirely in shared memory
#endif

#define S3R_CUDA_CHECK(x)                    \
  do {                                       \
    cudaError_t e__ = (x);                   \
    if (e__ != cudaSuccess) return S3R_ERR_CUDA; \
  } while (0)

struct S3rViewConst {  // per-view constants staged in shared memory by the per-Gaussian kernels
  float vm[16];
  float pm[16];
  float campos[3];
  float tanx, tany, scale, scale2;
  int set;
};

static inline int64_t s3r_align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// kernels (defined in the .cu files, launched from raster_api.cu)
int s3r_launch_preprocess(const s3r_raster_params& p, const s3r_raster_layout& L, char* state, int32_t* radii,
                          cudaStream_t st);
int s3r_launch_bin(const s3r_raster_params& p, const s3r_raster_layout& L, char* state, int64_t capacity,
                   cudaStream_t st);
int s3r_launch_sort(const s3r_raster_params& p, const s3r_raster_layout& L, char* state, cudaStream_t st);
int s3r_launch_blend(const s3r_raster_params& p, const s3r_raster_outputs& o, const s3r_raster_layout& L,
                     char* state, cudaStream_t st);
