// Shared declarations of the styl3r_b200 CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/styl3r_b200.h"

#define S3R_CHUNK 256          // Gaussians per preprocess / emit CTA (= one bit-mask row of 8 words)
#define S3R_REC_FLOATS 12      // blend record: x y B' C' | A' o r g | b depth cellmask -  ((B', C'), (r, g), (b, depth) are
                               // aligned register pairs for the packed f32x2 instructions; per Gaussian the last two
                               // floats are the cull half-extents (ex, ey), which the sort epilogue turns into the
                               // instance's 16-bit mask over the 4x4-pixel cells of its tile)
// The record carries the conic pre-scaled into the log2 domain, A' = -0.5*log2(e)*A, B' = -log2(e)*B,
// C' = -0.5*log2(e)*C, so that the blend kernels evaluate  log2(G) = dx*(A'*dx + B'*dy) + C'*dy*dy  with two FMAs and
// feed MUFU.EX2 directly (5 FP32 ops instead of 10).  Decisions that must equal the oracle's unfused fp32 arithmetic
// (power > 0, alpha >= 1/255) are re-taken from the exact per-Gaussian conic (conic_opacity[]) whenever the fast value
// lands inside a guard band around a threshold.
#define S3R_KA (-0.72134752044448170368f)   // -0.5 * log2(e)
#define S3R_KB (-1.44269504088896340736f)   // -log2(e)
#define S3R_INV_KA (-1.38629436111989061883f)  // A = A' * (-2 ln 2)
#define S3R_INV_KB (-0.69314718055994530942f)  // B = B' * (-ln 2)
#define S3R_ALPHA_BAND 1e-4f   // relative half-width of the guard band around alpha = 1/255
#define S3R_PZERO_BAND 1e-5f   // |log2 G| below this: the sign of the exponent is decided exactly
#define S3R_REC_BYTES 48
#ifndef S3R_SORT_SMEM_CAP
// per-tile instances sorted entirely in shared memory (2 x 8 B each).  3584 -> 56 KB + 8 KB of histograms per CTA: three
// CTAs per SM like 4096, but they leave ~30 KB of the SM's 228 KB to L1, which the epilogue's random gathers of
// (xy, conic, rgb) need: tile sort 30.1 -> 27.9 us per launch, 19.8k -> 21.2k views/s (scripts/sweep_sortcap.sh, A/B/A/B)
#define S3R_SORT_SMEM_CAP 3584
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
  float fx, fy, limx, limy;  // focal lengths in pixels and frustum clamp limits (computeCov2D)
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

// persistent grid of the blend kernel on the current device: SM count and resident CTAs (raster_blend.cu); the tile sort
// lays the blend work queue out for it
int s3r_blend_grid(int* sms, int* slots);

// Programmatic dependent launch switch shared by all kernels (S3R_TUNE_PDL; defined in gemm_tcgen05.cu)
int s3r_pdl_enabled();
int& s3r_blend_only_tile();  // S3R_TUNE_BLEND_ONLY_TILE (defined in raster_blend.cu)
int& s3r_blend_kernel_choice();  // S3R_TUNE_BLEND_KERNEL: 0 = warp-granular kernel over block lists, 1 = tile-granular kernel
int s3r_launch_blend_blocks(const s3r_raster_params& p, const s3r_raster_outputs& o, const s3r_raster_layout& L,
                            char* state, cudaStream_t st, int only_tile);
int& s3r_bwd_mode_override();  // S3R_TUNE_BWD_ALL (defined in raster_backward.cu)

// Launch with the programmatic-stream-serialization attribute: the grid may become resident while the previous kernel
// of the stream drains; the kernel runs its global-memory-free prologue (shared-memory zeroing, mbarrier init), then
// s3r_grid_dependency_sync() blocks until the previous kernel has completed and flushed.  Captured as a programmatic
// edge by CUDA graphs.  Without the attribute the device-side instructions are no-ops.
int& s3r_raster_pdl_mask();  // S3R_TUNE_RASTER_PDL: bit k = stage k (preprocess, scan, emit, sort, blend) launches with PDL

template <typename... KArgs, typename... Args>
static inline cudaError_t s3r_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                         bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (pdl && s3r_pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    na++;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

#ifdef __CUDACC__
__device__ __forceinline__ void s3r_grid_dependency_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: remember what was configured per device
// (a process-wide flag would leave a second GPU at the 48 KB default).
template <typename F>
static inline int s3r_ensure_dynamic_smem(F* func, size_t bytes, size_t (&configured)[64]) {
  if (bytes <= 48 * 1024) return S3R_OK;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return S3R_ERR_CUDA;
  dev &= 63;
  if (bytes > configured[dev]) {
    if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
      return S3R_ERR_CUDA;
    configured[dev] = bytes;
  }
  return S3R_OK;
}
